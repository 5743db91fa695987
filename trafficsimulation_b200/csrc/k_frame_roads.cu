// k_frame_roads.cu -- frame + band rasterisation + lane directions + sidewalks + highway entrances
// as ONE closed-form write kernel.
//
// Replaces (Simulation/city_model.py): _place_thick_wall :315-319, _place_sidewalk_inner_ring
// :329-360, _clear_interior :366-369, and _build_roads_and_sidewalks from the band lists on
// (:396-495) with _make_intersection :211-306, _compute_lane_dirs :1275-1368,
// _override_corner_lane_dirs :498-558, _replace_boundary_highways_with_entrances :1370-1420.
//
// Every one of those passes is a function of (x, y), the two band descriptors covering the cell and
// the descriptors of its 4 neighbours, so the 13 in-place sweeps of the reference collapse into a
// single pass that only WRITES the planes: 1 (type) + 2 (dirs) + 1 (aux) B/cell, no reads except
// the O(W+H) line tables (L1/L2 resident).  block_id is written as a whole by the zoning pass.
#include "cells_frame.cuh"

namespace tsim {

template <int VEC>
__global__ void __launch_bounds__(256) frame_roads_kernel(tsim_cfg c, uint8_t *__restrict__ T, uint16_t *__restrict__ D,
                                                          uint8_t *__restrict__ A, int32_t *__restrict__ B,
                                                          const uint32_t *__restrict__ rowt, const uint32_t *__restrict__ colt) {
    const Geo g(c);
    const int W = g.W, H = g.H;
    const int xblocks = (W + VEC * 256 - 1) / (VEC * 256);
    const int xv = ((blockIdx.x % xblocks) * blockDim.x + threadIdx.x) * VEC;
    const int ly = blockIdx.x / xblocks;            // local row inside the shard allocation
    const int y = c.win_y0 + ly;
    if (xv >= W || y < 0 || y >= H) return;
    const uint32_t r0 = y > 0 ? __ldg(rowt + y - 1) : 0u, r1 = __ldg(rowt + y), r2 = y + 1 < H ? __ldg(rowt + y + 1) : 0u;
    uint32_t ce[VEC + 2];
#pragma unroll
    for (int i = 0; i < VEC + 2; i++) {
        const int x = xv - 1 + i;
        ce[i] = (x >= 0 && x < W) ? __ldg(colt + x) : 0u;
    }
    uint8_t tt[VEC], aa[VEC];
    uint16_t dd[VEC];
#pragma unroll
    for (int i = 0; i < VEC; i++) {
        const int x = xv + i;
        int t = T_WALL;
        uint32_t d = 0, a = 0;
        if (x < W) frame_roads_cell(c, g, r0, r1, r2, ce[i], ce[i + 1], ce[i + 2], x, y, t, d, a);
        tt[i] = (uint8_t)t; dd[i] = (uint16_t)d; aa[i] = (uint8_t)a;
    }
    const size_t base = (size_t)ly * W + xv;
    if (VEC == 4) {
        *reinterpret_cast<uchar4 *>(T + base) = make_uchar4(tt[0], tt[1], tt[2], tt[3]);
        *reinterpret_cast<ushort4 *>(D + base) = make_ushort4(dd[0], dd[1], dd[2], dd[3]);
        *reinterpret_cast<uchar4 *>(A + base) = make_uchar4(aa[0], aa[1], aa[2], aa[3]);
    } else {
#pragma unroll
        for (int i = 0; i < VEC; i++)
            if (xv + i < W) { T[base + i] = tt[i]; D[base + i] = dd[i]; A[base + i] = aa[i]; }
    }
}

}  // namespace tsim

using namespace tsim;

extern "C" tsim_status tsim_layout_frame_roads(const tsim_cfg *cfg, const tsim_planes *p, const tsim_lines *lines, void *stream) {
    tsim_status s = check_cfg(cfg);
    if (s != TSIM_OK) return s;
    if (!p || !p->cell_type || !p->dirs || !p->aux || !p->block_id || !lines || !lines->row || !lines->col) {
        set_error("tsim_layout_frame_roads: NULL plane or line table");
        return TSIM_ERR_CONFIG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int W = cfg->width, rows = cfg->win_rows;
    if (W % 4 == 0) {
        dim3 grid((unsigned)div_up(W, 4 * 256) * rows);
        frame_roads_kernel<4><<<grid, 256, 0, st>>>(*cfg, p->cell_type, p->dirs, p->aux, p->block_id, lines->row, lines->col);
    } else {
        dim3 grid((unsigned)div_up(W, 256) * rows);
        frame_roads_kernel<1><<<grid, 256, 0, st>>>(*cfg, p->cell_type, p->dirs, p->aux, p->block_id, lines->row, lines->col);
    }
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}
