// cells_frame.cuh -- closed-form per-cell logic of the frame + road rasterisation pass.
// __host__ __device__ so that tests/hostemu can run the very same code on the CPU.
#pragma once
#include "common.cuh"

namespace tsim {

enum : int { K_NONE = 0, K_ROAD = 1, K_INTER = 2 };

struct Cls {
    int kind, type, horiz, off, size, dir;
    bool ring;
};

__host__ __device__ __forceinline__ int thick_of(int rt) { return rt == 1 ? 4 : (rt == 2 ? 2 : (rt == 3 ? 1 : 0)); }  // config.py:45-49

// city_model.py:401-451 (+ the part of _make_intersection that decides intersection vs inner lane)
__host__ __device__ __forceinline__ Cls classify(const tsim_cfg &c, const Geo &g, uint32_t re, uint32_t ce, int x, int y) {
    Cls o;
    o.kind = K_NONE; o.type = 0; o.horiz = 0; o.off = 0; o.size = 0; o.dir = -1; o.ring = false;
    if (x < 0 || x >= g.W || y < 0 || y >= g.H) return o;
    const bool hv = lt_valid(re), vv = lt_valid(ce);
    if (hv && vv) {
        const int ht = lt_type(re), vt = lt_type(ce);
        if ((ht != 1 || vt != 1) && !g.inside(x, y)) return o;
        if (c.ring_road_type != 0) {   // forced ring corners stay plain road (:415-430)
            const int ft = thick_of(c.ring_road_type);
            const bool yb = y >= g.iymin && y < g.iymin + ft, yt = y >= g.iymax - ft + 1 && y <= g.iymax;
            const bool xl = x >= g.ixmin && x < g.ixmin + ft, xr = x >= g.ixmax - ft + 1 && x <= g.ixmax;
            if ((yb || yt) && (xl || xr)) {
                o.kind = K_ROAD; o.type = ht; o.horiz = 1; o.off = lt_off(re); o.size = lt_size(re); o.dir = lt_dir(re);
                o.ring = true;
                return o;
            }
        }
        const int hs = lt_size(re), vs = lt_size(ce);
        const bool svm = (hs == 1 && vs > 1) || (vs == 1 && hs > 1);
        if (c.optimized_intersections && svm) {   // only the outer lanes of the thick road cross (:277-303)
            if (hs > 1) {
                const int mo = lt_off(re);
                if (mo != 0 && mo != hs - 1) { o.kind = K_ROAD; o.type = ht; o.horiz = 1; o.off = mo; o.size = hs; o.dir = lt_dir(re); return o; }
            } else {
                const int mo = lt_off(ce);
                if (mo != 0 && mo != vs - 1) { o.kind = K_ROAD; o.type = vt; o.horiz = 0; o.off = mo; o.size = vs; o.dir = lt_dir(ce); return o; }
            }
        }
        o.kind = K_INTER;
        return o;
    }
    if (hv) {
        const int ht = lt_type(re);
        if (ht != 1 && !g.inside(x, y)) return o;
        o.kind = K_ROAD; o.type = ht; o.horiz = 1; o.off = lt_off(re); o.size = lt_size(re); o.dir = lt_dir(re);
    } else if (vv) {
        const int vt = lt_type(ce);
        if (vt != 1 && !g.inside(x, y)) return o;
        o.kind = K_ROAD; o.type = vt; o.horiz = 0; o.off = lt_off(ce); o.size = lt_size(ce); o.dir = lt_dir(ce);
    }
    return o;
}

// frame passes in closed form (:315-369)
__host__ __device__ __forceinline__ int base_type(const Geo &g, int x, int y) {
    if (g.inside(x, y)) return T_NOTHING;
    const int W = g.W, H = g.H, ws = g.ws, sr = g.sr;
    const bool xin = x >= ws && x < W - ws, yin = y >= ws && y < H - ws;
    const int a = y - ws, b = H - ws - 1 - y, c2 = x - ws, d = W - ws - 1 - x;
    const bool face_h = xin && ((a >= 0 && a < sr) || (b >= 0 && b < sr));
    const bool face_v = yin && ((c2 >= 0 && c2 < sr) || (d >= 0 && d < sr));
    return (face_h || face_v) ? T_SIDEWALK : T_WALL;
}


// One cell of the fused pass.  r0/r1/r2 = row descriptors of y-1,y,y+1; c0/c1/c2 = column
// descriptors of x-1,x,x+1 (0 when out of range).
__host__ __device__ __forceinline__ void frame_roads_cell(const tsim_cfg &c, const Geo &g, uint32_t r0, uint32_t r1, uint32_t r2,
                                                          uint32_t c0, uint32_t c1, uint32_t c2, int x, int y,
                                                          int &t_out, uint32_t &d_out, uint32_t &a_out) {
    const int W = g.W, H = g.H;
    int t = T_WALL;
    uint32_t d = 0, a = 0;
    const Cls me = classify(c, g, r1, c1, x, y);
    if (me.kind == K_INTER) {
        t = T_INTER; d = DL_NSEW; a = AUX_EVER;
    } else if (me.kind == K_ROAD) {
        t = T_R1 - 1 + me.type;
        a = me.ring ? AUX_RING : 0;
        if (me.type == 3) {
            d = me.dir >= 0 ? dl_one(me.dir) : 0;
        } else if (me.type == 2) {   // :1289-1305
            d = me.horiz ? dl_one(me.off == 0 ? DE : DW) : dl_one(me.off == 0 ? DS : DN);
        } else {                      // R1, :1308-1365
            const int half = me.size / 2;
            // neighbours towards the lower / higher lane offset
            const bool lo_int = me.horiz ? classify(c, g, r0, c1, x, y - 1).kind == K_INTER
                                         : classify(c, g, r1, c0, x - 1, y).kind == K_INTER;
            const bool hi_int = me.horiz ? classify(c, g, r2, c1, x, y + 1).kind == K_INTER
                                         : classify(c, g, r1, c2, x + 1, y).kind == K_INTER;
            const int d_lo = me.horiz ? DS : DW, d_hi = me.horiz ? DN : DE;
            if (me.off < half) {
                d = dl_one(me.horiz ? DE : DS);
                if (me.off > 0 && !lo_int) d = dl_append(d, d_lo);
                if (me.off < half - 1 && !hi_int) d = dl_append(d, d_hi);
            } else {
                d = dl_one(me.horiz ? DW : DN);
                if (me.off < me.size - 1 && !hi_int) d = dl_append(d, d_hi);
                if (me.off > half && !lo_int) d = dl_append(d, d_lo);
            }
        }
        // forced-corner override, ring R2 only (:498-558)
        if (c.ring_road_type == 2) {
            const bool in_b = lt_in_first(r1), in_t = lt_in_last(r1);
            const bool in_l = lt_in_first(c1), in_r = lt_in_last(c1);
            if ((in_b || in_t) && (in_l || in_r)) {
                int lr, lc, code;   // 4 entries of 2 bits, index = local_row*2 + local_col
                // BL {(0,0):E,(0,1):E,(1,0):S,(1,1):N}  BR {E,N,W,N}  TR {S,N,W,W}  TL {S,E,S,W}
                if (in_b && in_l) { lr = lt_off_first(r1); lc = lt_off_first(c1); code = DE | (DE << 2) | (DS << 4) | (DN << 6); }
                else if (in_b && in_r) { lr = lt_off_first(r1); lc = lt_off_last(c1); code = DE | (DN << 2) | (DW << 4) | (DN << 6); }
                else if (in_t && in_r) { lr = lt_off_last(r1); lc = lt_off_last(c1); code = DS | (DN << 2) | (DW << 4) | (DW << 6); }
                else { lr = lt_off_last(r1); lc = lt_off_first(c1); code = DS | (DE << 2) | (DS << 4) | (DW << 6); }
                if (lr < 2 && lc < 2) d = dl_one((code >> (2 * (lr * 2 + lc))) & 3);
            }
        }
        // boundary highway cells (:1370-1420)
        if (t == T_R1 && (x == 0 || x == W - 1 || y == 0 || y == H - 1) &&
            (y < g.ws || y >= H - g.ws || x < g.ws || x >= W - g.ws)) {
            bool ent = false;
            if (x == W - 1) ent |= dl_has(d, DW); else if (x == 0) ent |= dl_has(d, DE);
            if (y == H - 1) ent |= dl_has(d, DS); else if (y == 0) ent |= dl_has(d, DN);
            t = ent ? T_HWY_IN : T_HWY_OUT;
        }
    } else {
        t = base_type(g, x, y);
        if (t != T_SIDEWALK) {   // sidewalk around roads (:470-492)
            const Cls n0 = classify(c, g, r1, c2, x + 1, y), n1 = classify(c, g, r1, c0, x - 1, y);
            const Cls n2 = classify(c, g, r2, c1, x, y + 1), n3 = classify(c, g, r0, c1, x, y - 1);
            if (t == T_NOTHING) {
                if (n0.kind | n1.kind | n2.kind | n3.kind) t = T_SIDEWALK;
            } else {   // Wall next to a highway lane
                if ((n0.kind == K_ROAD && n0.type == 1) || (n1.kind == K_ROAD && n1.type == 1) ||
                    (n2.kind == K_ROAD && n2.type == 1) || (n3.kind == K_ROAD && n3.type == 1)) t = T_SIDEWALK;
            }
        }
    }
    t_out = t; d_out = d; a_out = a;
}

}  // namespace tsim
