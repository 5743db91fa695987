// cells_stencil.cuh -- per-cell logic of the 1-halo stencil passes, written against a "view" so the
// same code runs from global memory, from a shared-memory tile, or on the CPU (tests/hostemu).
#pragma once
#include "common.cuh"

namespace tsim {

// Plane view over one shard allocation: rows [y0, y0 + nrows) of a W x H grid.
struct PlaneView {
    const uint8_t *T;
    const uint16_t *D;
    const uint8_t *A;
    int W, H, y0, nrows;
    __host__ __device__ __forceinline__ bool has(int x, int y) const { return x >= 0 && x < W && y >= y0 && y < y0 + nrows; }
    __host__ __device__ __forceinline__ size_t at(int x, int y) const { return (size_t)(y - y0) * W + x; }
    __host__ __device__ __forceinline__ int t(int x, int y) const { return has(x, y) ? (int)T[at(x, y)] : -1; }
    __host__ __device__ __forceinline__ uint32_t d(int x, int y) const { return D[at(x, y)]; }
};

__host__ __device__ __forceinline__ bool is_road_like(int t) { return t >= 0 && in_set(SET_ROAD_LIKE, t); }

// L4 (city_model.py:811-840): is this cell a dead end right now?
template <class V>
__host__ __device__ __forceinline__ bool dead_end_cell(const V &v, int x, int y) {
    const int t = v.t(x, y);
    if (t < 0 || !in_set(SET_REMOVABLE, t)) return false;
    const int n = (int)is_road_like(v.t(x + 1, y)) + (int)is_road_like(v.t(x - 1, y)) +
                  (int)is_road_like(v.t(x, y + 1)) + (int)is_road_like(v.t(x, y - 1));
    return n < 2;
}

// R2 lane directions (city_model.py:1289-1305); only the types an R2 cell can revert to matter
__host__ __device__ __forceinline__ uint32_t simple_lane_dirs(int rtype, int horiz, int off, int dir) {
    if (rtype == 3) return dir >= 0 ? dl_one(dir) : 0;
    if (rtype == 2) return horiz ? dl_one(off == 0 ? DE : DW) : dl_one(off == 0 ? DS : DN);
    return 0;
}

// _make_intersection (city_model.py:211-306) on the live grid.  Returns 0 = no change,
// 1 = becomes Intersection, 2 = reverts to a plain road cell (t_out/d_out), 3 = unsupported (R1 revert).
template <class V>
__host__ __device__ __forceinline__ int make_intersection_cell(const tsim_cfg &c, const V &v, uint32_t re, uint32_t ce, int x, int y,
                                                               int &t_out, uint32_t &d_out) {
    const int st = T_R1 - 1 + c.subblock_road_type;
    int h_sz, h_off, h_rt, h_bd, v_sz, v_off, v_rt, v_bd;
    bool have_h = false, have_v = false;
    if (lt_valid(re)) { h_sz = lt_size(re); h_off = lt_off(re); h_rt = lt_type(re); h_bd = lt_dir(re); have_h = true; }
    else if (v.t(x, y) == st || v.t(x - 1, y) == st || v.t(x + 1, y) == st) { h_sz = 1; h_off = 0; h_rt = c.subblock_road_type; h_bd = -1; have_h = true; }
    if (lt_valid(ce)) { v_sz = lt_size(ce); v_off = lt_off(ce); v_rt = lt_type(ce); v_bd = lt_dir(ce); have_v = true; }
    else if (v.t(x, y) == st || v.t(x, y - 1) == st || v.t(x, y + 1) == st) { v_sz = 1; v_off = 0; v_rt = c.subblock_road_type; v_bd = -1; have_v = true; }
    if (!(have_h && have_v)) return 0;
    const bool svm = (h_sz == 1 && v_sz > 1) || (v_sz == 1 && h_sz > 1);
    if (c.optimized_intersections && svm) {
        const bool hm = h_sz > 1;
        const int mrt = hm ? h_rt : v_rt, moff = hm ? h_off : v_off, msz = hm ? h_sz : v_sz, mbd = hm ? h_bd : v_bd;
        if (moff != 0 && moff != msz - 1) {
            if (mrt == 1) return 3;
            t_out = T_R1 - 1 + mrt;
            d_out = simple_lane_dirs(mrt, hm ? 1 : 0, moff, mbd);
            return 2;
        }
    }
    if (v.t(x, y) == T_INTER) return 0;
    t_out = T_INTER;
    d_out = DL_NSEW;
    return 1;
}

// L5 (city_model.py:842-879)
template <class V>
__host__ __device__ __forceinline__ int upgrade_r2_cell(const tsim_cfg &c, const V &v, uint32_t re, uint32_t ce, int x, int y,
                                                        int &t_out, uint32_t &d_out) {
    if (v.t(x, y) != T_R2) return 0;
    if (c.ring_road_type == 2 && (lt_in_first(re) || lt_in_last(re)) && (lt_in_first(ce) || lt_in_last(ce))) return 0;
    const int sw = (int)(v.t(x, y + 1) == T_SIDEWALK) + (int)(v.t(x, y - 1) == T_SIDEWALK) +
                   (int)(v.t(x + 1, y) == T_SIDEWALK) + (int)(v.t(x - 1, y) == T_SIDEWALK);
    if (sw < 2) return 0;
    return make_intersection_cell(c, v, re, ce, x, y, t_out, d_out);
}

// L7 (city_model.py:969-1012): new ordered list of an Intersection cell
template <class V>
__host__ __device__ __forceinline__ uint32_t validate_dirs_cell(const V &v, int x, int y, uint32_t od) {
    uint32_t nd = 0;
    const int n = dl_len(od);
    for (int i = 0; i < n; i++) {
        const int d = dl_get(od, i);
        const int nx = x + dx_of(d), ny = y + dy_of(d);
        const int nt = v.t(nx, ny);
        if (!is_road_like(nt)) continue;
        if (nt == T_INTER || dl_has(v.d(nx, ny), d)) nd = dl_append(nd, d);
    }
    return nd;
}

// L8 (city_model.py:1035-1070) in gather form.  For a BlockEntrance: its own list, in neighbour
// order +x,-x,+y,-y.  For any other road-like cell: arrows INTO adjacent entrances are appended in
// the raster order of those entrances (y-1, x-1, x+1, y+1).
template <class V>
__host__ __device__ __forceinline__ uint32_t entrance_dirs_cell(const V &v, int x, int y, int t, uint32_t od) {
    if (t == T_BE) {
        uint32_t ed = 0;
        if (is_road_like(v.t(x + 1, y))) ed = dl_append(ed, DE);
        if (is_road_like(v.t(x - 1, y))) ed = dl_append(ed, DW);
        if (is_road_like(v.t(x, y + 1))) ed = dl_append(ed, DN);
        if (is_road_like(v.t(x, y - 1))) ed = dl_append(ed, DS);
        return ed;
    }
    if (!is_road_like(t)) return od;
    uint32_t nd = od;
    if (v.t(x, y - 1) == T_BE && !dl_has(nd, DS)) nd = dl_append(nd, DS);
    if (v.t(x - 1, y) == T_BE && !dl_has(nd, DW)) nd = dl_append(nd, DW);
    if (v.t(x + 1, y) == T_BE && !dl_has(nd, DE)) nd = dl_append(nd, DE);
    if (v.t(x, y + 1) == T_BE && !dl_has(nd, DN)) nd = dl_append(nd, DN);
    return nd;
}

// L10 (city_model.py:2151-2199)
__host__ __device__ __forceinline__ void maps_cell(int t, uint32_t d, uint32_t a, uint8_t &is_road, uint8_t &road_type,
                                                   uint8_t &inter, uint8_t &allowed) {
    const bool rl = in_set(SET_ROAD_LIKE, t);
    uint8_t rt = 0;
    if (rl) {
        if (t == T_R2) rt = (a & AUX_RING) ? 1 : 2;
        else if (t == T_R3) rt = 3;
        else rt = 1;   // R1, Intersection, HighwayEntrance, HighwayExit, BlockEntrance
    }
    is_road = rl; road_type = rt; inter = (t == T_INTER); allowed = (uint8_t)(d & 0xf);
}

}  // namespace tsim
