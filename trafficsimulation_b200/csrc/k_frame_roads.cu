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
#include <algorithm>
#include <cstdlib>
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

// Bulk region: cells whose 3x3 neighbourhood lies inside the interior and outside the forced ring-corner
// squares (thickness <= 4) and the frame.  There the cell is a function of the descriptor triples only.
struct Bulk {
    int x0, x1, y0, y1;
    __host__ __device__ explicit Bulk(const Geo &g) : x0(g.ixmin + 5), x1(g.ixmax - 5), y0(g.iymin + 5), y1(g.iymax - 5) {}
};

// 16 cells per thread.  Bulk strips: one class look-up per cell from the CTA's shared copy of the row's
// table line, three 128-bit stores.  Everything else (the frame, ~0.5 % of a large city): closed form.
__global__ void __launch_bounds__(256) frame_roads_lut_kernel(tsim_cfg c, uint8_t *__restrict__ T, uint16_t *__restrict__ D, uint8_t *__restrict__ A,
                                                              const uint32_t *__restrict__ rowt, const uint32_t *__restrict__ colt,
                                                              const uint8_t *__restrict__ rowc, const uint8_t *__restrict__ colc,
                                                              const uint32_t *__restrict__ lut, int ncc) {
    __shared__ uint32_t s_lut[256];
    const Geo g(c);
    const Bulk bk(g);
    const int W = g.W, H = g.H;
    const int xblocks = (W + 16 * 256 - 1) / (16 * 256);
    const int ly = blockIdx.x / xblocks;
    const int y = c.win_y0 + ly;
    const bool row_bulk = y >= bk.y0 && y <= bk.y1;
    if (row_bulk) {
        const uint32_t *line = lut + (size_t)rowc[y] * ncc;
        for (int i = threadIdx.x; i < ncc; i += blockDim.x) s_lut[i] = line[i];
    }
    __syncthreads();
    const int xv = ((blockIdx.x % xblocks) * blockDim.x + threadIdx.x) * 16;
    if (xv >= W) return;
    uint32_t tw[4] = {0, 0, 0, 0}, aw[4] = {0, 0, 0, 0}, dw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (row_bulk && xv >= bk.x0 && xv + 15 <= bk.x1) {
        const uint4 cq = __ldg(reinterpret_cast<const uint4 *>(colc + xv));
        const uint32_t cw[4] = {cq.x, cq.y, cq.z, cq.w};
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t e = s_lut[(cw[k >> 2] >> (8 * (k & 3))) & 0xffu];
            tw[k >> 2] |= (e & 0xffu) << (8 * (k & 3));
            aw[k >> 2] |= ((e >> 8) & 0xffu) << (8 * (k & 3));
            dw[k >> 1] |= (e >> 16) << (16 * (k & 1));
        }
    } else {
        const uint32_t r0 = y > 0 ? __ldg(rowt + y - 1) : 0u, r1 = __ldg(rowt + y), r2 = y + 1 < H ? __ldg(rowt + y + 1) : 0u;
        uint32_t cprev = xv > 0 ? __ldg(colt + xv - 1) : 0u, ccur = __ldg(colt + xv);
        for (int k = 0; k < 16; k++) {   // W % 16 == 0: the strip is inside the row
            const int x = xv + k;
            const uint32_t cnext = x + 1 < W ? __ldg(colt + x + 1) : 0u;
            int t; uint32_t d, a;
            frame_roads_cell(c, g, r0, r1, r2, cprev, ccur, cnext, x, y, t, d, a);
            tw[k >> 2] |= (uint32_t)t << (8 * (k & 3));
            aw[k >> 2] |= a << (8 * (k & 3));
            dw[k >> 1] |= d << (16 * (k & 1));
            cprev = ccur; ccur = cnext;
        }
    }
    const size_t base = (size_t)ly * W + xv;
    *reinterpret_cast<uint4 *>(T + base) = make_uint4(tw[0], tw[1], tw[2], tw[3]);
    *reinterpret_cast<uint4 *>(A + base) = make_uint4(aw[0], aw[1], aw[2], aw[3]);
    *reinterpret_cast<uint4 *>(D + base) = make_uint4(dw[0], dw[1], dw[2], dw[3]);
    *reinterpret_cast<uint4 *>(D + base + 8) = make_uint4(dw[4], dw[5], dw[6], dw[7]);
}

// Bulk strips are a straight copy of the row class's pattern row (~60 rows, cache resident): four 128-bit loads, four
// 128-bit stores per 16 cells.  Frame strips: closed form, as in the look-up kernel.
__global__ void __launch_bounds__(256) frame_roads_rows_kernel(tsim_cfg c, uint8_t *__restrict__ T, uint16_t *__restrict__ D, uint8_t *__restrict__ A,
                                                               const uint32_t *__restrict__ rowt, const uint32_t *__restrict__ colt,
                                                               const uint8_t *__restrict__ rowc, const uint8_t *__restrict__ pt,
                                                               const uint16_t *__restrict__ pd, const uint8_t *__restrict__ pa) {
    const Geo g(c);
    const Bulk bk(g);
    const int W = g.W, H = g.H;
    const int xblocks = (W + 16 * 256 - 1) / (16 * 256);
    const int ly = blockIdx.x / xblocks;
    const int y = c.win_y0 + ly;
    const int xv = ((blockIdx.x % xblocks) * blockDim.x + threadIdx.x) * 16;
    if (xv >= W) return;
    const size_t base = (size_t)ly * W + xv;
    if (y >= bk.y0 && y <= bk.y1 && xv >= bk.x0 && xv + 15 <= bk.x1) {
        const size_t src = (size_t)__ldg(rowc + y) * W + xv;
        *reinterpret_cast<uint4 *>(T + base) = __ldg(reinterpret_cast<const uint4 *>(pt + src));
        *reinterpret_cast<uint4 *>(A + base) = __ldg(reinterpret_cast<const uint4 *>(pa + src));
        *reinterpret_cast<uint4 *>(D + base) = __ldg(reinterpret_cast<const uint4 *>(pd + src));
        *reinterpret_cast<uint4 *>(D + base + 8) = __ldg(reinterpret_cast<const uint4 *>(pd + src + 8));
        return;
    }
    uint32_t tw[4] = {0, 0, 0, 0}, aw[4] = {0, 0, 0, 0}, dw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint32_t r0 = y > 0 ? __ldg(rowt + y - 1) : 0u, r1 = __ldg(rowt + y), r2 = y + 1 < H ? __ldg(rowt + y + 1) : 0u;
    uint32_t cprev = xv > 0 ? __ldg(colt + xv - 1) : 0u, ccur = __ldg(colt + xv);
    for (int k = 0; k < 16; k++) {   // W % 16 == 0: the strip is inside the row
        const int x = xv + k;
        const uint32_t cnext = x + 1 < W ? __ldg(colt + x + 1) : 0u;
        int t; uint32_t d, a;
        frame_roads_cell(c, g, r0, r1, r2, cprev, ccur, cnext, x, y, t, d, a);
        tw[k >> 2] |= (uint32_t)t << (8 * (k & 3));
        aw[k >> 2] |= a << (8 * (k & 3));
        dw[k >> 1] |= d << (16 * (k & 1));
        cprev = ccur; ccur = cnext;
    }
    *reinterpret_cast<uint4 *>(T + base) = make_uint4(tw[0], tw[1], tw[2], tw[3]);
    *reinterpret_cast<uint4 *>(A + base) = make_uint4(aw[0], aw[1], aw[2], aw[3]);
    *reinterpret_cast<uint4 *>(D + base) = make_uint4(dw[0], dw[1], dw[2], dw[3]);
    *reinterpret_cast<uint4 *>(D + base + 8) = make_uint4(dw[4], dw[5], dw[6], dw[7]);
}

// ---- the bulk of the pass as TMA bulk copies ------------------------------------------------------------------------------
// Away from the frame every row of one class IS that class's pattern row, so the bulk of the pass is a broadcast of ~60
// cache-resident rows into 16 384 rows of the planes: pure data movement.  A CTA takes one (row class, x tile, row chunk):
// one thread fetches the tile of the three pattern rows into shared memory with cp.async.bulk (completion on an mbarrier),
// then every thread walks its share of the chunk's rows and, for the rows of that class, issues the three bulk stores shared ->
// global (cp.async.bulk.global.shared::cta, SASS UBLKCP).  No register ever holds a cell: the SM issues a few hundred copy
// descriptors instead of 8 load / store instructions per 16 cells, and the stores leave as whole 2 - 4 KB bursts.
constexpr int TMA_ROWS = 1024;   // rows per CTA; TX = cells per tile (TX bytes of types, TX of aux, 2 TX of dirs in shared memory)

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int TMA_TX>
__global__ void __launch_bounds__(128) frame_roads_tma_kernel(tsim_cfg c, uint8_t *__restrict__ T, uint16_t *__restrict__ D, uint8_t *__restrict__ A,
                                                              const uint8_t *__restrict__ rowc, const uint8_t *__restrict__ pt,
                                                              const uint16_t *__restrict__ pd, const uint8_t *__restrict__ pa, int xs, int xe) {
    __shared__ __align__(128) uint8_t s_t[TMA_TX];
    __shared__ __align__(128) uint8_t s_a[TMA_TX];
    __shared__ __align__(128) uint16_t s_d[TMA_TX];
    __shared__ __align__(8) unsigned long long s_bar;
    const Geo g(c);
    const Bulk bk(g);
    const int W = g.W;
    const int x0 = xs + (int)blockIdx.x * TMA_TX, len = min(TMA_TX, xe - x0);
    const int k = blockIdx.y;
    const int ly0 = (int)blockIdx.z * TMA_ROWS, ly1 = min(ly0 + TMA_ROWS, c.win_rows);
    if (len <= 0) return;
    // does this chunk hold a row of class k at all?  (most (class, chunk) pairs do not: leave before fetching anything)
    bool mine = false;
    for (int ly = ly0 + threadIdx.x; ly < ly1; ly += blockDim.x) {
        const int y = c.win_y0 + ly;
        mine |= y >= bk.y0 && y <= bk.y1 && rowc[y] == k;
    }
    if (!__syncthreads_or(mine)) return;
    const uint32_t bar = smem_addr(&s_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const size_t src = (size_t)k * W + x0;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4u * (uint32_t)len) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(s_t)), "l"(pt + src),
                     "r"((uint32_t)len), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(s_a)), "l"(pa + src),
                     "r"((uint32_t)len), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(s_d)), "l"(pd + src),
                     "r"(2u * (uint32_t)len), "r"(bar) : "memory");
    }
    __syncthreads();   // the barrier is initialised before anybody polls it
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
    bool issued = false;
    for (int ly = ly0 + threadIdx.x; ly < ly1; ly += blockDim.x) {
        const int y = c.win_y0 + ly;
        if (!(y >= bk.y0 && y <= bk.y1 && rowc[y] == k)) continue;
        const size_t dst = (size_t)ly * W + x0;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(T + dst), "r"(smem_addr(s_t)), "r"((uint32_t)len) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(A + dst), "r"(smem_addr(s_a)), "r"((uint32_t)len) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(D + dst), "r"(smem_addr(s_d)), "r"(2u * (uint32_t)len) : "memory");
        issued = true;
    }
    if (issued) {   // the tile must stay in shared memory until the copy engine has read it
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// The bulk as a register copy with little else in the kernel: a thread owns one 16-cell strip column and copies it for R
// consecutive rows -- the row classes first, then 4 loads and 4 stores per row, all independent -- so that a warp has
// R x 4 x 512 bytes in flight and the kernel needs 30-odd registers (the one-strip-per-thread kernel that also evaluated the
// frame's closed form needed 55 and ran at 34 % occupancy, 2.9 TB/s; a write-only probe with hashed data reaches 6 - 6.9 TB/s
// on this device, profiles/r2_write_peak.json).
template <int R>
__global__ void __launch_bounds__(256) frame_copy_kernel(tsim_cfg c, uint8_t *__restrict__ T, uint16_t *__restrict__ D, uint8_t *__restrict__ A,
                                                         const uint8_t *__restrict__ rowc, const uint8_t *__restrict__ pt,
                                                         const uint16_t *__restrict__ pd, const uint8_t *__restrict__ pa, int xs, int xe) {
    const Geo g(c);
    const Bulk bk(g);
    const int W = g.W;
    const int xv = xs + (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (xv >= xe) return;
    const int ly0 = blockIdx.y * R;
    int cls[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int ly = ly0 + r, y = c.win_y0 + ly;
        cls[r] = (ly < c.win_rows && y >= bk.y0 && y <= bk.y1) ? (int)__ldg(rowc + y) : -1;
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        if (cls[r] < 0) continue;
        const size_t src = (size_t)cls[r] * W + xv, dst = (size_t)(ly0 + r) * W + xv;
        const uint4 t = __ldg(reinterpret_cast<const uint4 *>(pt + src)), a = __ldg(reinterpret_cast<const uint4 *>(pa + src));
        const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(pd + src)), d1 = __ldg(reinterpret_cast<const uint4 *>(pd + src + 8));
        *reinterpret_cast<uint4 *>(T + dst) = t;
        *reinterpret_cast<uint4 *>(A + dst) = a;
        *reinterpret_cast<uint4 *>(D + dst) = d0;
        *reinterpret_cast<uint4 *>(D + dst + 8) = d1;
    }
}

// what the bulk copies leave out, closed form, ONE launch, four cells per thread (a quad of lanes = one 16-cell strip; the closed
// form is a few hundred instructions per cell, so a thread that walks a whole strip is a 30 us dependent chain):
//   items [0, n0): in every row of the window, the strips left of xs and right of xe;
//   items [n0, n0 + n1): in the rows outside the bulk rows, the strips between xs and xe.
__global__ void __launch_bounds__(128) frame_edges_kernel(tsim_cfg c, uint8_t *__restrict__ T, uint16_t *__restrict__ D, uint8_t *__restrict__ A,
                                                          const uint32_t *__restrict__ rowt, const uint32_t *__restrict__ colt, int xs, int xe,
                                                          long long n0, long long n1) {
    const Geo g(c);
    const Bulk bk(g);
    const int W = g.W, H = g.H;
    const int n_left = xs / 16, n_all = W / 16, first_right = xe / 16;
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long it = gt >> 2;
    const int q = (int)(gt & 3);
    if (it >= n0 + n1) return;
    int ly, strip;
    if (it < n0) {
        const int per = n_left + (n_all - first_right);
        ly = (int)(it / per);
        const int e = (int)(it % per);
        strip = e < n_left ? e : first_right + (e - n_left);
    } else {
        it -= n0;
        const int per = first_right - n_left;
        const int lo_rows = max(0, min(c.win_rows, bk.y0 - c.win_y0));        // local rows [0, lo_rows) lie below the bulk rows
        const int hi_first = max(0, min(c.win_rows, bk.y1 + 1 - c.win_y0));   // local rows [hi_first, win_rows) lie above them
        const int k = (int)(it / per);
        ly = k < lo_rows ? k : hi_first + (k - lo_rows);
        strip = n_left + (int)(it % per);
    }
    const int y = c.win_y0 + ly, xv = strip * 16 + 4 * q;
    const uint32_t r0 = y > 0 ? __ldg(rowt + y - 1) : 0u, r1 = __ldg(rowt + y), r2 = y + 1 < H ? __ldg(rowt + y + 1) : 0u;
    uint32_t tw = 0, aw = 0, dw[2] = {0, 0};
    uint32_t cprev = xv > 0 ? __ldg(colt + xv - 1) : 0u, ccur = __ldg(colt + xv);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = xv + k;
        const uint32_t cnext = x + 1 < W ? __ldg(colt + x + 1) : 0u;
        int t; uint32_t d, a;
        frame_roads_cell(c, g, r0, r1, r2, cprev, ccur, cnext, x, y, t, d, a);
        tw |= (uint32_t)t << (8 * k);
        aw |= a << (8 * k);
        dw[k >> 1] |= d << (16 * (k & 1));
        cprev = ccur; ccur = cnext;
    }
    const size_t base = (size_t)ly * W + xv;
    *reinterpret_cast<uint32_t *>(T + base) = tw;
    *reinterpret_cast<uint32_t *>(A + base) = aw;
    *reinterpret_cast<uint2 *>(D + base) = make_uint2(dw[0], dw[1]);
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
    const bool aligned = !(((uintptr_t)p->cell_type | (uintptr_t)p->dirs | (uintptr_t)p->aux) & 15);
    const bool patterns = W % 16 == 0 && aligned && lines->row_class && lines->pat_type && lines->pat_dirs && lines->pat_aux &&
                          !(((uintptr_t)lines->pat_type | (uintptr_t)lines->pat_dirs | (uintptr_t)lines->pat_aux) & 15);
    const Geo g(*cfg);
    const Bulk bk(g);
    const int xs = (bk.x0 + 15) & ~15, xe = (bk.x1 + 1) & ~15;   // the 16-cell strips that lie inside the bulk columns
    const char *tma_env = getenv("TSIM_FRAME_TMA");
    const int tma_tx = tma_env ? atoi(tma_env) : 0;               // TSIM_FRAME_TMA=<tile cells>: 2048 or 8192 switch the bulk-copy kernel on
    const char *copy_env = getenv("TSIM_FRAME_COPY");             // "legacy": the one-strip-per-thread kernel (copy + frame in one launch)
    const bool legacy = copy_env && *copy_env == 'l';
    auto edges = [&]() -> tsim_status {   // frame strips of every row, then the rows outside the bulk rows
        const long long n0 = (long long)(xs / 16 + (W / 16 - xe / 16)) * rows;
        const int lo_rows = std::max(0, std::min(rows, bk.y0 - cfg->win_y0)), hi_first = std::max(0, std::min(rows, bk.y1 + 1 - cfg->win_y0));
        const long long n1 = (long long)(xe / 16 - xs / 16) * (lo_rows + (rows - hi_first));
        if (n0 + n1 > 0) {
            frame_edges_kernel<<<div_up(4 * (n0 + n1), 128), 128, 0, st>>>(*cfg, p->cell_type, p->dirs, p->aux, lines->row, lines->col, xs, xe, n0, n1);
            TSIM_LAUNCH_CHECK();
        }
        return TSIM_OK;
    };
    if (patterns && xe - xs >= 64 && lines->n_row_classes > 0 && lines->n_row_classes <= 65535 && (tma_tx == 2048 || tma_tx == 8192)) {
        dim3 grid((unsigned)div_up(xe - xs, tma_tx), (unsigned)lines->n_row_classes, (unsigned)div_up(rows, TMA_ROWS));
        if (tma_tx == 2048)
            frame_roads_tma_kernel<2048><<<grid, 128, 0, st>>>(*cfg, p->cell_type, p->dirs, p->aux, lines->row_class, lines->pat_type, lines->pat_dirs,
                                                               lines->pat_aux, xs, xe);
        else
            frame_roads_tma_kernel<8192><<<grid, 128, 0, st>>>(*cfg, p->cell_type, p->dirs, p->aux, lines->row_class, lines->pat_type, lines->pat_dirs,
                                                               lines->pat_aux, xs, xe);
        TSIM_LAUNCH_CHECK();
        return edges();
    } else if (patterns && xe - xs >= 64 && !legacy) {
        constexpr int R = 8;
        dim3 grid((unsigned)div_up(xe - xs, 16 * 256), (unsigned)div_up(rows, R));
        if (grid.y > 65535) { set_error("tsim_layout_frame_roads: window of %d rows too tall for the copy kernel", rows); return TSIM_ERR_CONFIG; }
        frame_copy_kernel<R><<<grid, 256, 0, st>>>(*cfg, p->cell_type, p->dirs, p->aux, lines->row_class, lines->pat_type, lines->pat_dirs, lines->pat_aux,
                                                   xs, xe);
        TSIM_LAUNCH_CHECK();
        return edges();
    } else if (patterns) {
        dim3 grid((unsigned)div_up(W, 16 * 256) * rows);
        frame_roads_rows_kernel<<<grid, 256, 0, st>>>(*cfg, p->cell_type, p->dirs, p->aux, lines->row, lines->col, lines->row_class, lines->pat_type,
                                                      lines->pat_dirs, lines->pat_aux);
    } else if (W % 16 == 0 && aligned && lines->lut && lines->row_class && lines->col_class && lines->n_col_classes > 0 && lines->n_col_classes <= 256 &&
        !((uintptr_t)lines->col_class & 15)) {
        dim3 grid((unsigned)div_up(W, 16 * 256) * rows);
        frame_roads_lut_kernel<<<grid, 256, 0, st>>>(*cfg, p->cell_type, p->dirs, p->aux, lines->row, lines->col, lines->row_class, lines->col_class,
                                                     lines->lut, lines->n_col_classes);
    } else if (W % 4 == 0) {
        dim3 grid((unsigned)div_up(W, 4 * 256) * rows);
        frame_roads_kernel<4><<<grid, 256, 0, st>>>(*cfg, p->cell_type, p->dirs, p->aux, p->block_id, lines->row, lines->col);
    } else {
        dim3 grid((unsigned)div_up(W, 256) * rows);
        frame_roads_kernel<1><<<grid, 256, 0, st>>>(*cfg, p->cell_type, p->dirs, p->aux, p->block_id, lines->row, lines->col);
    }
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}
