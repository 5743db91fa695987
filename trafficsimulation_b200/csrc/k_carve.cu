// k_carve.cu -- _carve_subblock_roads (city_model.py:563-737): one warp per batch of 32 Nothing blobs.
//
// The reference visits blobs in raster discovery order and carves each with the decisions it draws
// (the carve tape, one row per blob).  Blobs are disjoint and a carve only touches its own blob, the
// sidewalk ring between the blob and the first road it reaches, and that first road cell -- where the
// update (_make_intersection, or one extra arrow) is idempotent -- so blobs can be carved
// concurrently.  Work per blob is O(perimeter): two legs, two extensions, the pivot's 8 neighbours.
#include "cells_stencil.cuh"

namespace tsim {

struct LiveGrid {   // read/write access to the live planes of the window, LOCAL row numbers (own writes are visible to the thread)
    uint8_t *T; uint16_t *D; uint8_t *A;
    int W, H;       // H = rows of the window
    __device__ __forceinline__ bool has(int x, int y) const { return x >= 0 && x < W && y >= 0 && y < H; }
    __device__ __forceinline__ size_t at(int x, int y) const { return (size_t)y * W + x; }
    __device__ __forceinline__ int t(int x, int y) const { return has(x, y) ? (int)((volatile uint8_t *)T)[at(x, y)] : -1; }
    __device__ __forceinline__ uint32_t d(int x, int y) const { return ((volatile uint16_t *)D)[at(x, y)]; }
    __device__ __forceinline__ void place(int x, int y, int t) const {   // place_cell (:1864-1870)
        const size_t i = at(x, y);
        T[i] = (uint8_t)t; D[i] = 0; A[i] &= (AUX_RING | AUX_EVER);
    }
};

// lay_r4_cell (:588-601)
__device__ __forceinline__ void lay_cell(const LiveGrid &g, int sub_t, int x, int y, int arrow) {
    if (!g.has(x, y)) return;
    if (!is_road_like(g.t(x, y))) {
        g.place(x, y, sub_t);
        g.D[g.at(x, y)] = (uint16_t)dl_one(arrow);
    }
    if (g.t(x + 1, y) == T_NOTHING) g.place(x + 1, y, T_SIDEWALK);
    if (g.t(x - 1, y) == T_NOTHING) g.place(x - 1, y, T_SIDEWALK);
    if (g.t(x, y + 1) == T_NOTHING) g.place(x, y + 1, T_SIDEWALK);
    if (g.t(x, y - 1) == T_NOTHING) g.place(x, y - 1, T_SIDEWALK);
}

// extend_to_road (:603-627)
__device__ void extend(const tsim_cfg &c, const LiveGrid &g, const uint32_t *rowt /* + win_y0 */, const uint32_t *colt, int sub_t, int sx, int sy,
                       int march, int arrow, int32_t *err) {
    int cx = sx, cy = sy;
    while (g.has(cx, cy)) {
        const int t = g.t(cx, cy);
        if (is_road_like(t)) {
            const size_t i = g.at(cx, cy);
            if (c.subblock_roads_have_intersections) {
                int t_new; uint32_t d_new;
                const int r = make_intersection_cell(c, g, __ldg(rowt + cy), __ldg(colt + cx), cx, cy, t_new, d_new);
                if (r == 3) { *err = 3; }
                else if (r != 0) { g.place(cx, cy, t_new); g.D[i] = (uint16_t)d_new; }
                g.A[i] |= AUX_EVER;   // :617 adds the cell to _intersection_cells unconditionally
            } else {
                // one extra arrow into the road.  Two blobs may reach the same cell from opposite sides;
                // CAS keeps both arrows.  (List order then follows arrival order, not blob order: flagged.)
                const uint32_t od = g.d(cx, cy);
                if (!dl_has(od, arrow)) {
                    unsigned short *p = (unsigned short *)(g.D + i);
                    unsigned short seen = (unsigned short)od;
                    for (;;) {
                        if (dl_has(seen, arrow)) break;
                        const unsigned short prev = atomicCAS(p, seen, (unsigned short)dl_append(seen, arrow));
                        if (prev == seen) break;
                        seen = prev;
                        *err = 4;   // concurrent append: order may differ from the reference
                    }
                }
            }
            return;
        }
        if (t == T_SIDEWALK || t == T_NOTHING) { lay_cell(g, sub_t, cx, cy, arrow); cx += dx_of(march); cy += dy_of(march); }
        else return;
    }
}

// One warp per BATCH of 32 blobs (lane L holds the decisions of blob L of the batch).  Carving one blob is a chain of a dozen
// dependent memory round trips (tape row, blob box, leg cells, their neighbours, the marches, the pivot's neighbours) for a few
// dozen cells; a warp that takes blobs one after the other spends its time waiting.  So every step is done for all 32 blobs at
// once, over the FLAT list of their leg cells (or pivot neighbours), two items per lane in flight:
//   (1) every leg cell becomes road, (2) every Nothing neighbour of a leg cell becomes Sidewalk -- two steps that commute with
//   the reference's cell-by-cell order (a leg cell that the serial order would first edge as Sidewalk and then overwrite as road
//   ends up road either way); (3) the pivot's arrow; (4) the two extensions of a blob march on different lines outside the blob:
//   lane L marches those of its own blob; (5) the pivot's 8 neighbours.
// Blobs are disjoint (header), so doing a step for 32 of them together orders nothing that was ordered before.
struct CarveJob {   // one blob's taped decisions, checked
    int px, py, hd, vd, nh, nv, h_arrow, v_arrow, inb_h, minx, miny, maxx, maxy;
};

// f(owner lane, index among the owner's items, valid) for the flat list of n_mine items per lane; all lanes call f together
template <class F>
__device__ __forceinline__ void warp_flat(int n_mine, int lane, F f) {
    constexpr uint32_t FULL = 0xffffffffu;
    int incl = n_mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
    const int total = __shfl_sync(FULL, incl, 31);
    for (int i0 = 0; i0 < total; i0 += 64) {
        int own[2], idx[2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int i = i0 + u * 32 + lane;
            int lo = 0;   // first lane with incl > i
#pragma unroll
            for (int step = 16; step; step >>= 1) { const int v = __shfl_sync(FULL, incl, min(lo + step - 1, 31)); if (v <= i) lo += step; }
            lo = min(lo, 31);
            own[u] = lo; idx[u] = i - __shfl_sync(FULL, incl - n_mine, lo); ok[u] = i < total;
        }
        f(own, idx, ok);
    }
}

__global__ void __launch_bounds__(256) carve_kernel(tsim_cfg c, uint8_t *T, uint16_t *D, uint8_t *A, const uint32_t *__restrict__ rowt,
                                                    const uint32_t *__restrict__ colt, const int32_t *__restrict__ blobs,
                                                    const int32_t *__restrict__ n_blobs, int cap_blobs, const int32_t *__restrict__ id_base,
                                                    const int32_t *__restrict__ tape, int n_tape, int32_t *err) {
    constexpr uint32_t FULL = 0xffffffffu;
    const int nb = min(*n_blobs, cap_blobs);
    const int base = id_base ? *id_base : 0;
    const int y0 = c.win_y0;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const LiveGrid g{T, D, A, c.width, c.win_rows};
    const uint32_t *rowl = rowt + y0;   // line table of local row ly = rowl[ly]
    const int sub_t = T_R1 - 1 + c.subblock_road_type;
    const int ms = c.min_subblock_spacing;
    for (int b0 = warp * 32; b0 < nb; b0 += nwarps * 32) {
        const int b = b0 + lane;
        CarveJob j{};
        bool valid = false;
        if (b < nb) {
            const int gid = b + base;            // 0-based global blob id = tape row; < 0: cut by the window's lower edge, owned (and carved) by the shard below
            if (gid >= n_tape) *err = 1;
            else if (gid >= 0) {
                const int4 r0 = *reinterpret_cast<const int4 *>(tape + (size_t)gid * 8), r1 = *reinterpret_cast<const int4 *>(tape + (size_t)gid * 8 + 4);
                const int32_t *bl = blobs + (size_t)b * TSIM_BLOB_STRIDE;
                const int b0x = bl[0], b1y = bl[1], b2x = bl[2], b3y = bl[3], area = bl[4];
                if (r0.y) {   // carved
                    j.minx = b0x; j.miny = b1y - y0; j.maxx = b2x; j.maxy = b3y - y0;   // window-local rows from here on
                    // a blob cut by a window edge that is not a grid edge lies in the halo: the shard that owns its rows sees it whole and carves it
                    if (!((j.miny == 0 && y0 > 0) || (j.maxy == c.win_rows - 1 && y0 + c.win_rows < c.height))) {
                        j.px = r0.z; j.py = r0.w - y0; j.hd = r1.x; j.vd = r1.y; j.inb_h = r1.z;
                        // the tape must hold a decision the reference could have drawn (:659-675)
                        if (!((j.hd == DW || j.hd == DE) && (j.vd == DN || j.vd == DS)) || j.px < j.minx + ms || j.px > j.maxx - ms || j.py < j.miny + ms ||
                            j.py > j.maxy - ms ||
                            (long long)(j.maxx - j.minx + 1) * (j.maxy - j.miny + 1) != area /* non-rectangular blob: carve footprints may interact */) {
                            *err = 2;
                        } else {
                            valid = true;
                            j.h_arrow = j.inb_h ? opp_of(j.hd) : j.hd;   // :683-696
                            j.v_arrow = j.inb_h ? j.vd : opp_of(j.vd);   // :707-708
                            // leg cells: horizontal leg from the pivot's neighbour to the blob edge, vertical leg from the pivot to the blob edge
                            j.nh = j.hd == DW ? j.px - j.minx : j.maxx - j.px;
                            j.nv = j.vd == DS ? j.py - j.miny + 1 : j.maxy - j.py + 1;
                        }
                    }
                }
            }
        }
        if (!__ballot_sync(FULL, valid)) continue;
        // leg cell `i` of the blob lane `o` holds
        auto leg_cell = [&](int o, int i, int &x, int &y, int &arrow) {
            const int px = __shfl_sync(FULL, j.px, o), py = __shfl_sync(FULL, j.py, o), hd = __shfl_sync(FULL, j.hd, o), vd = __shfl_sync(FULL, j.vd, o);
            const int nh = __shfl_sync(FULL, j.nh, o), ha = __shfl_sync(FULL, j.h_arrow, o), va = __shfl_sync(FULL, j.v_arrow, o);
            if (i < nh) { x = hd == DW ? px - 1 - i : px + 1 + i; y = py; arrow = ha; }
            else { const int q = i - nh; x = px; y = vd == DS ? py - q : py + q; arrow = va; }
        };
        const int n_leg = valid ? j.nh + j.nv : 0;
        warp_flat(n_leg, lane, [&](const int (&own)[2], const int (&idx)[2], const bool (&ok)[2]) {   // lay_r4_cell (:588-601), road part
            int x[2], y[2], arrow[2], t[2];
#pragma unroll
            for (int u = 0; u < 2; u++) { leg_cell(own[u], idx[u], x[u], y[u], arrow[u]); t[u] = (ok[u] && g.has(x[u], y[u])) ? g.t(x[u], y[u]) : -1; }
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (t[u] >= 0 && !is_road_like(t[u])) { g.place(x[u], y[u], sub_t); g.D[g.at(x[u], y[u])] = (uint16_t)dl_one(arrow[u]); }
        });
        __syncwarp();
        warp_flat(n_leg, lane, [&](const int (&own)[2], const int (&idx)[2], const bool (&ok)[2]) {   // lay_r4_cell, sidewalk edging
            int x[2], y[2], arrow, t[2][4];
            const int ox[4] = {1, -1, 0, 0}, oy[4] = {0, 0, 1, -1};
#pragma unroll
            for (int u = 0; u < 2; u++) {
                leg_cell(own[u], idx[u], x[u], y[u], arrow);
                const bool in = ok[u] && g.has(x[u], y[u]);
#pragma unroll
                for (int k = 0; k < 4; k++) t[u][k] = in ? g.t(x[u] + ox[k], y[u] + oy[k]) : -1;
            }
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int k = 0; k < 4; k++) if (t[u][k] == T_NOTHING) g.place(x[u] + ox[k], y[u] + oy[k], T_SIDEWALK);
        });
        __syncwarp();
        if (valid) {
            g.D[g.at(j.px, j.py)] = (uint16_t)dl_one(j.inb_h ? j.v_arrow : j.h_arrow);   // pivot shows the outbound arrow only (:713-715)
            const int hx_end = j.hd == DW ? j.minx : j.maxx, vy_end = j.vd == DS ? j.miny : j.maxy;
            extend(c, g, rowl, colt, sub_t, hx_end + dx_of(j.hd), j.py, j.hd, j.h_arrow, err);
            extend(c, g, rowl, colt, sub_t, j.px, vy_end + dy_of(j.vd), j.vd, j.v_arrow, err);
        }
        __syncwarp();
        warp_flat(valid ? 8 : 0, lane, [&](const int (&own)[2], const int (&idx)[2], const bool (&ok)[2]) {   // :731-737
            int x[2], y[2], t[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int px = __shfl_sync(FULL, j.px, own[u]), py = __shfl_sync(FULL, j.py, own[u]);
                const int k = idx[u] < 4 ? idx[u] : idx[u] + 1;
                x[u] = px + k % 3 - 1; y[u] = py + k / 3 - 1;
                t[u] = ok[u] ? g.t(x[u], y[u]) : -1;
            }
#pragma unroll
            for (int u = 0; u < 2; u++) if (t[u] >= 0 && !is_road_like(t[u]) && t[u] != T_WALL) g.place(x[u], y[u], T_SIDEWALK);
        });
        __syncwarp();
    }
}

}  // namespace tsim

using namespace tsim;

extern "C" tsim_status tsim_layout_carve(const tsim_cfg *cfg, const tsim_planes *p, const tsim_lines *lines, const tsim_blobs *blobs,
                                         const int32_t *tape, int32_t n_tape, int32_t *err_flag, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if ((st = check_blobs(blobs, "tsim_layout_carve")) != TSIM_OK) return st;
    if (!p || !p->cell_type || !p->dirs || !p->aux || !lines || !lines->row || !lines->col || !tape || !err_flag || n_tape < 0) {
        set_error("tsim_layout_carve: bad arguments");
        return TSIM_ERR_CONFIG;
    }
    if (n_tape == 0) return TSIM_OK;
    const int grid = div_up(blobs->cap, 8 * 32) < 148 * 8 ? div_up(blobs->cap, 8 * 32) : 148 * 8;   // one warp per 32 blobs
    carve_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*cfg, p->cell_type, p->dirs, p->aux, lines->row, lines->col, blobs->table,
                                                         blobs->count, blobs->cap, blobs->id_base, tape, n_tape, err_flag);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}
