// k_carve.cu -- _carve_subblock_roads (city_model.py:563-737): one warp per Nothing blob.
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

// One WARP per blob.  The cells of the two legs are laid by the lanes in parallel, in two steps that commute with the
// reference's cell-by-cell order: (1) every leg cell becomes road, (2) every Nothing neighbour of a leg cell becomes
// Sidewalk (a leg cell that the serial order would first edge as Sidewalk and then overwrite as road ends up road either
// way).  The two extensions march on different lines outside the blob and run on two lanes; the pivot's 8 neighbours on 8.
__global__ void __launch_bounds__(256) carve_kernel(tsim_cfg c, uint8_t *T, uint16_t *D, uint8_t *A, const uint32_t *__restrict__ rowt,
                                                    const uint32_t *__restrict__ colt, const int32_t *__restrict__ blobs,
                                                    const int32_t *__restrict__ n_blobs, int cap_blobs, const int32_t *__restrict__ id_base,
                                                    const int32_t *__restrict__ tape, int n_tape, int32_t *err) {
    const int nb = min(*n_blobs, cap_blobs);
    const int base = id_base ? *id_base : 0;
    const int y0 = c.win_y0;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int b = warp; b < nb; b += nwarps) {
        const int gid = b + base;            // 0-based global blob id = tape row
        if (gid < 0) continue;               // cut by the window's lower edge: owned (and carved) by the shard below
        if (gid >= n_tape) { if (lane == 0) *err = 1; continue; }
        const int32_t *row = tape + (size_t)gid * 8;
        if (!row[1]) continue;
        const int32_t *bl = blobs + (size_t)b * TSIM_BLOB_STRIDE;
        const int minx = bl[0], miny = bl[1] - y0, maxx = bl[2], maxy = bl[3] - y0;   // window-local rows from here on
        // a blob cut by a window edge that is not a grid edge lies in the halo: the shard that owns its rows sees it whole and carves it
        if ((miny == 0 && y0 > 0) || (maxy == c.win_rows - 1 && y0 + c.win_rows < c.height)) continue;
        const int px = row[2], py = row[3] - y0, hd = row[4], vd = row[5], inb_h = row[6];
        const int ms = c.min_subblock_spacing;
        // the tape must hold a decision the reference could have drawn (:659-675)
        if (!((hd == DW || hd == DE) && (vd == DN || vd == DS)) || px < minx + ms || px > maxx - ms || py < miny + ms || py > maxy - ms ||
            (long long)(maxx - minx + 1) * (maxy - miny + 1) != bl[4] /* non-rectangular blob: carve footprints may interact */) {
            if (lane == 0) *err = 2;
            continue;
        }
        const LiveGrid g{T, D, A, c.width, c.win_rows};
        const uint32_t *rowl = rowt + y0;   // line table of local row ly = rowl[ly]
        const int sub_t = T_R1 - 1 + c.subblock_road_type;
        const int h_arrow = inb_h ? opp_of(hd) : hd;   // :683-696
        const int v_arrow = inb_h ? vd : opp_of(vd);   // :707-708
        // leg cells: horizontal leg from the pivot's neighbour to the blob edge, vertical leg from the pivot to the blob edge
        const int nh = hd == DW ? px - minx : maxx - px, nv = vd == DS ? py - miny + 1 : maxy - py + 1;
        auto leg_cell = [&](int i, int &x, int &y, int &arrow) {
            if (i < nh) { x = hd == DW ? px - 1 - i : px + 1 + i; y = py; arrow = h_arrow; }
            else { const int j = i - nh; x = px; y = vd == DS ? py - j : py + j; arrow = v_arrow; }
        };
        for (int i = lane; i < nh + nv; i += 32) {   // lay_r4_cell (:588-601), road part
            int x, y, arrow;
            leg_cell(i, x, y, arrow);
            if (g.has(x, y) && !is_road_like(g.t(x, y))) { g.place(x, y, sub_t); g.D[g.at(x, y)] = (uint16_t)dl_one(arrow); }
        }
        __syncwarp();
        for (int i = lane; i < nh + nv; i += 32) {   // lay_r4_cell, sidewalk edging
            int x, y, arrow;
            leg_cell(i, x, y, arrow);
            if (!g.has(x, y)) continue;
            if (g.t(x + 1, y) == T_NOTHING) g.place(x + 1, y, T_SIDEWALK);
            if (g.t(x - 1, y) == T_NOTHING) g.place(x - 1, y, T_SIDEWALK);
            if (g.t(x, y + 1) == T_NOTHING) g.place(x, y + 1, T_SIDEWALK);
            if (g.t(x, y - 1) == T_NOTHING) g.place(x, y - 1, T_SIDEWALK);
        }
        __syncwarp();
        if (lane == 0) g.D[g.at(px, py)] = (uint16_t)dl_one(inb_h ? v_arrow : h_arrow);   // pivot shows the outbound arrow only (:713-715)
        const int hx_end = hd == DW ? minx : maxx, vy_end = vd == DS ? miny : maxy;
        if (lane == 0) extend(c, g, rowl, colt, sub_t, hx_end + dx_of(hd), py, hd, h_arrow, err);
        if (lane == 1) extend(c, g, rowl, colt, sub_t, px, vy_end + dy_of(vd), vd, v_arrow, err);
        __syncwarp();
        if (lane < 8) {   // :731-737
            const int k = lane < 4 ? lane : lane + 1, dx = k % 3 - 1, dy = k / 3 - 1;
            const int t = g.t(px + dx, py + dy);
            if (t >= 0 && !is_road_like(t) && t != T_WALL) g.place(px + dx, py + dy, T_SIDEWALK);
        }
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
    const int grid = div_up(blobs->cap, 8) < 148 * 8 ? div_up(blobs->cap, 8) : 148 * 8;   // one warp per blob
    carve_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*cfg, p->cell_type, p->dirs, p->aux, lines->row, lines->col, blobs->table,
                                                         blobs->count, blobs->cap, blobs->id_base, tape, n_tape, err_flag);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}
