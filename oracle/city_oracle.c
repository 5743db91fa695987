/*
 * city_oracle.c -- CPU restatement of the reference's layout passes.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the oracle of SURVEY.md §8(c): a plain-C, sequential, in-place restatement of
 * kurisu-n/TrafficSimulation `Simulation/city_model.py` (layout part of `CityModel.__init__`,
 * lines 124-139) operating on the packed planes described in oracle/refharness/harness.py.
 * It is pinned against the live Python reference by tests/test_oracle_vs_reference.py (runs
 * wherever /root/reference exists) and against the committed fixtures in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * The product (trafficsimulation_b200/) never does.
 *
 * Every function cites the reference lines it follows.  Loops keep the reference's visiting
 * order wherever the order can influence the result.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum {
    T_RES = 0, T_OFF, T_MAR, T_LEI, T_OTH, T_EMPTY, T_NOTHING, T_SIDEWALK, T_WALL,
    T_R1, T_R2, T_R3, T_INTER, T_HWY_IN, T_HWY_OUT, T_TL, T_TL_STOP, T_CR, T_CR_STOP, T_BE
};
enum { DN = 0, DE = 1, DS = 2, DW = 3 };
#define AUX_ORIG 0x1f
#define AUX_RING 0x20
#define AUX_EVER 0x40
#define AUX_LIGHT 0x80

typedef struct {
    int32_t W, H, wall, ring_w;
    int32_t ring_type;      /* 0 = None, 1..3 = R1..R3 (city_model.py:32) */
    int32_t optimized;      /* optimized_intersections */
    int32_t sub_int;        /* subblock_roads_have_intersections */
    int32_t sub_type;       /* subblock_road_type 1..3 */
    int32_t min_sub;        /* min_subblock_spacing */
    int32_t tl_range;       /* traffic_light_range */
    int32_t fwd;            /* forward_traffic_light_range */
    int32_t fwd_mode;       /* 0 Skip, 1 Include in Range, 2 Include as Extra */
    int32_t entr_level;     /* Defaults.BLOCK_ENTRANCE_ROAD_LEVEL */
    int32_t fast_reach;     /* 0: literal BFS leads_to; 1: SCC-accelerated (same answers) */
} ocfg;

typedef struct {
    ocfg c;
    uint8_t *T; uint16_t *D; uint8_t *A; int32_t *B;
    const int32_t *hb; int nh; const int32_t *vb; int nv; /* bands: start,end,type,dir */
} city;

static const int DX[4] = {0, 1, 0, -1};
static const int DY[4] = {1, 0, -1, 0};
static const int OPP[4] = {DS, DW, DN, DE};
static const int RIGHT_OF[4] = {DE, DS, DW, DN};  /* config.py:66 */
static const int THICK[4] = {0, 4, 2, 1};          /* config.py:45-49 */

/* ---- ordered direction lists packed in u16 ---------------------------------------- */
static inline int dl_len(uint16_t c) { return (c >> 12) & 7; }
static inline int dl_get(uint16_t c, int i) { return (c >> (4 + 2 * i)) & 3; }
static inline int dl_has(uint16_t c, int d) { return (c >> d) & 1; }
static inline uint16_t dl_append(uint16_t c, int d) {
    int n = dl_len(c);
    c = (uint16_t)(c & 0x0fff);
    c |= (uint16_t)(1u << d);
    c |= (uint16_t)(d << (4 + 2 * n));
    c |= (uint16_t)((n + 1) << 12);
    return c;
}
static inline uint16_t dl_one(int d) { return dl_append(0, d); }

static inline int inb(const city *m, int x, int y) { return x >= 0 && x < m->c.W && y >= 0 && y < m->c.H; }
static inline int64_t ix(const city *m, int x, int y) { return (int64_t)y * m->c.W + x; }
static inline int typ(const city *m, int x, int y) { return inb(m, x, y) ? m->T[ix(m, x, y)] : -1; }

static inline int road_like(int t) { /* config.py:68 */
    return t == T_R1 || t == T_R2 || t == T_R3 || t == T_INTER || t == T_HWY_IN || t == T_HWY_OUT || t == T_BE;
}

/* city_model.py:1864-1870: a fresh CellAgent -- type set, directions/light/road_type reset */
static void place_cell(city *m, int x, int y, int t) {
    int64_t i = ix(m, x, y);
    m->T[i] = (uint8_t)t;
    m->D[i] = 0;
    m->A[i] &= (AUX_RING | AUX_EVER);
}

static int interior_xmin(const city *m) { return m->c.wall + m->c.ring_w; }
static int interior_xmax(const city *m) { return m->c.W - (m->c.wall + m->c.ring_w) - 1; }
static int interior_ymin(const city *m) { return m->c.wall + m->c.ring_w; }
static int interior_ymax(const city *m) { return m->c.H - (m->c.wall + m->c.ring_w) - 1; }
static int inside_interior(const city *m, int x, int y) { /* :1798-1800 */
    return interior_xmin(m) <= x && x <= interior_xmax(m) && interior_ymin(m) <= y && y <= interior_ymax(m);
}

/* :1269-1273 first band in list order that covers `index`, or -1 */
static int find_band(const int32_t *b, int n, int index) {
    for (int i = 0; i < n; i++)
        if (b[4 * i] <= index && index <= b[4 * i + 1]) return i;
    return -1;
}

/* ===================================================================================== */
/* L0  frame: city_model.py:315-369                                                       */
/* ===================================================================================== */
void oracle_frame(city *m) {
    int W = m->c.W, H = m->c.H, ws = m->c.wall, sr = m->c.ring_w;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) { /* :315-319 */
            int64_t i = ix(m, x, y);
            m->T[i] = T_WALL; m->D[i] = 0; m->A[i] = 0; m->B[i] = 0;
        }
    for (int layer = 0; layer < sr; layer++) { /* :329-360 */
        int y_top = ws + layer, y_bottom = H - ws - 1 - layer;
        for (int x = ws; x < W - ws; x++) {
            if (typ(m, x, y_top) == T_WALL) place_cell(m, x, y_top, T_SIDEWALK);
            if (typ(m, x, y_bottom) == T_WALL) place_cell(m, x, y_bottom, T_SIDEWALK);
        }
        int x_left = ws + layer, x_right = W - ws - 1 - layer;
        for (int y = ws; y < H - ws; y++) {
            if (typ(m, x_left, y) == T_WALL) place_cell(m, x_left, y, T_SIDEWALK);
            if (typ(m, x_right, y) == T_WALL) place_cell(m, x_right, y, T_SIDEWALK);
        }
    }
    for (int y = interior_ymin(m); y <= interior_ymax(m); y++) /* :366-369 */
        for (int x = interior_xmin(m); x <= interior_xmax(m); x++)
            if (inb(m, x, y)) place_cell(m, x, y, T_NOTHING);
}

/* ===================================================================================== */
/* lane directions: city_model.py:1275-1368                                               */
/* ===================================================================================== */
static int next_is_inter(const city *m, int x, int y, int d) { /* :1026-1029 */
    return typ(m, x + DX[d], y + DY[d]) == T_INTER;
}

static uint16_t lane_dirs(const city *m, int x, int y, int rtype, int horizontal, int off, int size, int bdir) {
    if (rtype == 3) return bdir >= 0 ? dl_one(bdir) : 0;
    if (rtype == 2) {
        if (horizontal) return dl_one(off == 0 ? DE : DW);
        return dl_one(off == 0 ? DS : DN);
    }
    if (rtype == 1) {
        int half = size / 2;
        uint16_t c;
        if (horizontal) {
            if (off < half) {
                c = dl_one(DE);
                if (off > 0 && !next_is_inter(m, x, y, DS)) c = dl_append(c, DS);
                if (off < half - 1 && !next_is_inter(m, x, y, DN)) c = dl_append(c, DN);
            } else {
                c = dl_one(DW);
                if (off < size - 1 && !next_is_inter(m, x, y, DN)) c = dl_append(c, DN);
                if (off > half && !next_is_inter(m, x, y, DS)) c = dl_append(c, DS);
            }
        } else {
            if (off < half) {
                c = dl_one(DS);
                if (off > 0 && !next_is_inter(m, x, y, DW)) c = dl_append(c, DW);
                if (off < half - 1 && !next_is_inter(m, x, y, DE)) c = dl_append(c, DE);
            } else {
                c = dl_one(DN);
                if (off < size - 1 && !next_is_inter(m, x, y, DE)) c = dl_append(c, DE);
                if (off > half && !next_is_inter(m, x, y, DW)) c = dl_append(c, DW);
            }
        }
        return c;
    }
    return 0;
}

/* city_model.py:498-558 */
static uint16_t override_corner(const city *m, int rx, int ry, uint16_t def) {
    if (m->c.ring_type != 2 || m->nh == 0 || m->nv == 0) return def;
    const int32_t *hbot = m->hb, *htop = m->hb + 4 * (m->nh - 1);
    const int32_t *vl = m->vb, *vr = m->vb + 4 * (m->nv - 1);
    int in_bottom = hbot[0] <= ry && ry <= hbot[1];
    int in_top = htop[0] <= ry && ry <= htop[1];
    int in_left = vl[0] <= rx && rx <= vl[1];
    int in_right = vr[0] <= rx && rx <= vr[1];
    if (!((in_bottom || in_top) && (in_left || in_right))) return def;
    /* mapping[(local_row, local_col)] */
    static const int BL[2][2] = {{DE, DE}, {DS, DN}};
    static const int BR[2][2] = {{DE, DN}, {DW, DN}};
    static const int TR[2][2] = {{DS, DN}, {DW, DW}};
    static const int TLm[2][2] = {{DS, DE}, {DS, DW}};
    const int (*mp)[2];
    int lr, lc;
    if (in_bottom && in_left) { mp = BL; lr = ry - hbot[0]; lc = rx - vl[0]; }
    else if (in_bottom && in_right) { mp = BR; lr = ry - hbot[0]; lc = rx - vr[0]; }
    else if (in_top && in_right) { mp = TR; lr = ry - htop[0]; lc = rx - vr[0]; }
    else { mp = TLm; lr = ry - htop[0]; lc = rx - vl[0]; }
    if ((lr == 0 || lr == 1) && (lc == 0 || lc == 1)) return dl_one(mp[lr][lc]);
    return def;
}

/* ===================================================================================== */
/* intersection factory: city_model.py:211-306                                            */
/* ===================================================================================== */
static void make_intersection(city *m, int x, int y) {
    int st = m->c.sub_type + (T_R1 - 1);
    int hi = find_band(m->hb, m->nh, y), vi = find_band(m->vb, m->nv, x);
    int h_st = 0, h_en = 0, h_rt = 0, h_bd = -1, v_st = 0, v_en = 0, v_rt = 0, v_bd = -1;
    int have_h = 0, have_v = 0;
    if (hi >= 0) { h_st = m->hb[4 * hi]; h_en = m->hb[4 * hi + 1]; h_rt = m->hb[4 * hi + 2]; h_bd = m->hb[4 * hi + 3]; have_h = 1; }
    else if (typ(m, x, y) == st || typ(m, x - 1, y) == st || typ(m, x + 1, y) == st) {
        h_st = h_en = y; h_rt = m->c.sub_type; h_bd = -1; have_h = 1; }
    if (vi >= 0) { v_st = m->vb[4 * vi]; v_en = m->vb[4 * vi + 1]; v_rt = m->vb[4 * vi + 2]; v_bd = m->vb[4 * vi + 3]; have_v = 1; }
    else if (typ(m, x, y) == st || typ(m, x, y - 1) == st || typ(m, x, y + 1) == st) {
        v_st = v_en = x; v_rt = m->c.sub_type; v_bd = -1; have_v = 1; }
    if (!(have_h && have_v)) return;
    int h_sz = h_en - h_st + 1, h_off = y - h_st;
    int v_sz = v_en - v_st + 1, v_off = x - v_st;
    int svm = (h_sz == 1 && v_sz > 1) || (v_sz == 1 && h_sz > 1);
    if (m->c.optimized && svm) {
        int mrt, mh, moff, msz, mbd;
        if (h_sz > 1) { mrt = h_rt; mh = 1; moff = h_off; msz = h_sz; mbd = h_bd; }
        else { mrt = v_rt; mh = 0; moff = v_off; msz = v_sz; mbd = v_bd; }
        if (moff != 0 && moff != msz - 1) { /* inner lane -> plain road again (:285-299) */
            uint16_t d = lane_dirs(m, x, y, mrt, mh, moff, msz, mbd);
            place_cell(m, x, y, mrt + (T_R1 - 1));
            m->D[ix(m, x, y)] = d;
            m->A[ix(m, x, y)] &= (uint8_t)~AUX_EVER;
            return;
        }
    }
    if (typ(m, x, y) == T_INTER) return; /* :236-245 */
    place_cell(m, x, y, T_INTER);
    /* Defaults.AVAILABLE_DIRECTIONS = N,S,E,W (config.py:62) */
    m->D[ix(m, x, y)] = dl_append(dl_append(dl_append(dl_one(DN), DS), DE), DW);
    m->A[ix(m, x, y)] |= AUX_EVER;
}

/* ===================================================================================== */
/* L1b roads + sidewalks + highway entrances: city_model.py:375-495, 1370-1420             */
/* ===================================================================================== */
void oracle_roads(city *m) {
    int W = m->c.W, H = m->c.H;
    int64_t N = (int64_t)W * H;
    /* _road_cells: rtype,horizontal,offset,size,dir per cell (0 type = absent) */
    int8_t *rc_t = calloc(N, 1), *rc_h = calloc(N, 1), *rc_o = calloc(N, 1), *rc_s = calloc(N, 1), *rc_d = calloc(N, 1);
    uint8_t *isec = calloc(N, 1); /* _intersection_cells */
    for (int y = 0; y < H; y++) {
        int hi = find_band(m->hb, m->nh, y);
        for (int x = 0; x < W; x++) {
            int vi = find_band(m->vb, m->nv, x);
            int64_t i = ix(m, x, y);
            if (hi >= 0 && vi >= 0) {
                const int32_t *hb = m->hb + 4 * hi, *vb = m->vb + 4 * vi;
                if ((hb[2] != 1 || vb[2] != 1) && !inside_interior(m, x, y)) continue;
                if (m->c.ring_type != 0) { /* forced ring corners :415-430 */
                    int ft = THICK[m->c.ring_type];
                    int yb = (y >= interior_ymin(m) && y < interior_ymin(m) + ft);
                    int yt = (y >= interior_ymax(m) - ft + 1 && y <= interior_ymax(m));
                    int xl = (x >= interior_xmin(m) && x < interior_xmin(m) + ft);
                    int xr = (x >= interior_xmax(m) - ft + 1 && x <= interior_xmax(m));
                    if ((yb || yt) && (xl || xr)) {
                        rc_t[i] = (int8_t)hb[2]; rc_h[i] = 1; rc_o[i] = (int8_t)(y - hb[0]);
                        rc_s[i] = (int8_t)(hb[1] - hb[0] + 1); rc_d[i] = (int8_t)hb[3];
                        m->A[i] |= AUX_RING;
                        continue;
                    }
                }
                isec[i] = 1;
            } else if (hi >= 0) {
                const int32_t *hb = m->hb + 4 * hi;
                if (hb[2] != 1 && !inside_interior(m, x, y)) continue;
                rc_t[i] = (int8_t)hb[2]; rc_h[i] = 1; rc_o[i] = (int8_t)(y - hb[0]);
                rc_s[i] = (int8_t)(hb[1] - hb[0] + 1); rc_d[i] = (int8_t)hb[3];
            } else if (vi >= 0) {
                const int32_t *vb = m->vb + 4 * vi;
                if (vb[2] != 1 && !inside_interior(m, x, y)) continue;
                rc_t[i] = (int8_t)vb[2]; rc_h[i] = 0; rc_o[i] = (int8_t)(x - vb[0]);
                rc_s[i] = (int8_t)(vb[1] - vb[0] + 1); rc_d[i] = (int8_t)vb[3];
            }
        }
    }
    /* mark intersections (:454-455); inner lanes of single x multi crossings fall back to road
       cells whose record equals what make_intersection would store (:296-298). */
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int64_t i = ix(m, x, y);
            if (!isec[i]) continue;
            m->A[i] |= AUX_EVER;
            make_intersection(m, x, y);
            if (!(m->A[i] & AUX_EVER)) { /* reverted: record it as a road cell */
                int hi = find_band(m->hb, m->nh, y), vi = find_band(m->vb, m->nv, x);
                const int32_t *hb = m->hb + 4 * hi, *vb = m->vb + 4 * vi;
                int h_sz = hb[1] - hb[0] + 1;
                isec[i] = 0;
                if (h_sz > 1) { rc_t[i] = (int8_t)hb[2]; rc_h[i] = 1; rc_o[i] = (int8_t)(y - hb[0]); rc_s[i] = (int8_t)h_sz; rc_d[i] = (int8_t)hb[3]; }
                else { rc_t[i] = (int8_t)vb[2]; rc_h[i] = 0; rc_o[i] = (int8_t)(x - vb[0]); rc_s[i] = (int8_t)(vb[1] - vb[0] + 1); rc_d[i] = (int8_t)vb[3]; }
            }
        }
    /* mark roads (:458-468) -- lane dirs read the final intersection types */
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int64_t i = ix(m, x, y);
            if (!rc_t[i] || isec[i]) continue;
            place_cell(m, x, y, rc_t[i] + (T_R1 - 1));
        }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int64_t i = ix(m, x, y);
            if (!rc_t[i] || isec[i]) continue;
            uint16_t d = lane_dirs(m, x, y, rc_t[i], rc_h[i], rc_o[i], rc_s[i], rc_d[i]);
            m->D[i] = override_corner(m, x, y, d);
        }
    /* sidewalk around roads (:470-492) */
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int64_t i = ix(m, x, y);
            if (!rc_t[i] && !isec[i]) continue;
            for (int k = 0; k < 4; k++) {
                int nx = x + DX[k], ny = y + DY[k];
                if (!inb(m, nx, ny)) continue;
                int64_t j = ix(m, nx, ny);
                if (rc_t[j] || isec[j]) continue;
                if (m->T[j] == T_NOTHING) place_cell(m, nx, ny, T_SIDEWALK);
                else if (m->T[j] == T_WALL && (m->T[i] == T_R1 || m->T[i] == T_HWY_IN || m->T[i] == T_HWY_OUT))
                    place_cell(m, nx, ny, T_SIDEWALK);
            }
        }
    /* boundary R1 -> HighwayEntrance / HighwayExit (:1370-1420) */
    int ws = m->c.wall;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            if (!(y < ws || y >= H - ws || x < ws || x >= W - ws)) continue;
            int64_t i = ix(m, x, y);
            if (m->T[i] != T_R1) continue;
            if (!(x == 0 || x == W - 1 || y == 0 || y == H - 1)) continue;
            uint16_t od = m->D[i];
            int entrance = 0;
            /* dict lookups: a later key overwrites an earlier equal key (w==1 / h==1) */
            if (x == W - 1) entrance |= dl_has(od, DW); else if (x == 0) entrance |= dl_has(od, DE);
            if (y == H - 1) entrance |= dl_has(od, DS); else if (y == 0) entrance |= dl_has(od, DN);
            place_cell(m, x, y, entrance ? T_HWY_IN : T_HWY_OUT);
            m->D[i] = od;
        }
    free(rc_t); free(rc_h); free(rc_o); free(rc_s); free(rc_d); free(isec);
}

/* ===================================================================================== */
/* flood fill helpers (4-connected components of one type, raster discovery order)          */
/* ===================================================================================== */
typedef struct { int32_t minx, miny, maxx, maxy, size, seed; } blob_t;

/* labels[i] = 1-based component id in raster discovery order (y outer, x inner), 0 elsewhere.
   returns number of components; fills blobs (caller frees). */
static int label_type(const city *m, int t, int32_t *labels, blob_t **out) {
    int W = m->c.W, H = m->c.H;
    int64_t N = (int64_t)W * H;
    memset(labels, 0, N * sizeof(int32_t));
    int32_t *stack = malloc(N * sizeof(int32_t));
    int cap = 1024, n = 0;
    blob_t *bl = malloc(cap * sizeof(blob_t));
    for (int64_t s = 0; s < N; s++) {
        if (m->T[s] != t || labels[s]) continue;
        if (n == cap) { cap *= 2; bl = realloc(bl, cap * sizeof(blob_t)); }
        blob_t b = {W, H, -1, -1, 0, (int32_t)s};
        int sp = 0;
        stack[sp++] = (int32_t)s; labels[s] = n + 1;
        while (sp) {
            int32_t c = stack[--sp];
            int x = c % W, y = c / W;
            if (x < b.minx) b.minx = x; if (x > b.maxx) b.maxx = x;
            if (y < b.miny) b.miny = y; if (y > b.maxy) b.maxy = y;
            b.size++;
            for (int k = 0; k < 4; k++) {
                int nx = x + DX[k], ny = y + DY[k];
                if (!inb(m, nx, ny)) continue;
                int64_t j = ix(m, nx, ny);
                if (m->T[j] == t && !labels[j]) { labels[j] = n + 1; stack[sp++] = (int32_t)j; }
            }
        }
        bl[n++] = b;
    }
    free(stack);
    *out = bl;
    return n;
}

/* ===================================================================================== */
/* L2 carve sub-block roads: city_model.py:563-737                                         */
/* tape row per Nothing blob in discovery order:                                           */
/*   (drawn, carved, px, py, hor_dir, ver_dir, inbound_is_horizontal, tries)               */
/* ===================================================================================== */
static void lay_r4(city *m, int x, int y, int arrow) { /* :588-601 */
    if (!inb(m, x, y)) return;
    if (!road_like(typ(m, x, y))) {
        place_cell(m, x, y, m->c.sub_type + (T_R1 - 1));
        m->D[ix(m, x, y)] = dl_one(arrow);
    }
    for (int k = 0; k < 4; k++) {
        static const int ox[4] = {1, -1, 0, 0}, oy[4] = {0, 0, 1, -1};
        int nx = x + ox[k], ny = y + oy[k];
        if (typ(m, nx, ny) == T_NOTHING) place_cell(m, nx, ny, T_SIDEWALK);
    }
}

static void extend_to_road(city *m, int sx, int sy, int march, int arrow) { /* :603-627 */
    int cx = sx, cy = sy;
    while (inb(m, cx, cy)) {
        int t = typ(m, cx, cy);
        if (road_like(t)) {
            if (m->c.sub_int) {
                make_intersection(m, cx, cy);
                m->A[ix(m, cx, cy)] |= AUX_EVER; /* :617 unconditional add */
            } else if (!dl_has(m->D[ix(m, cx, cy)], arrow)) {
                m->D[ix(m, cx, cy)] = dl_append(m->D[ix(m, cx, cy)], arrow);
            }
            break;
        }
        if (t == T_SIDEWALK || t == T_NOTHING) { lay_r4(m, cx, cy, arrow); cx += DX[march]; cy += DY[march]; }
        else break;
    }
}

/* returns number of blobs visited, or -1 if the tape is too short */
int oracle_carve(city *m, const int32_t *tape, int n_tape) {
    int W = m->c.W, H = m->c.H;
    int64_t N = (int64_t)W * H;
    uint8_t *visited = calloc(N, 1);
    int32_t *stack = malloc(N * sizeof(int32_t));
    int nb = 0;
    for (int64_t s = 0; s < N; s++) { /* :632-635 raster discovery on the LIVE grid */
        if (visited[s] || m->T[s] != T_NOTHING) continue;
        int minx = W, miny = H, maxx = -1, maxy = -1, sp = 0;
        stack[sp++] = (int32_t)s; visited[s] = 1;
        while (sp) {
            int32_t c = stack[--sp];
            int x = c % W, y = c / W;
            if (x < minx) minx = x; if (x > maxx) maxx = x;
            if (y < miny) miny = y; if (y > maxy) maxy = y;
            for (int k = 0; k < 4; k++) {
                int nx = x + DX[k], ny = y + DY[k];
                if (!inb(m, nx, ny)) continue;
                int64_t j = ix(m, nx, ny);
                if (m->T[j] == T_NOTHING && !visited[j]) { visited[j] = 1; stack[sp++] = (int32_t)j; }
            }
        }
        if (nb >= n_tape) { free(visited); free(stack); return -1; }
        const int32_t *row = tape + 8 * nb;
        nb++;
        if (!row[1]) continue;
        int px = row[2], py = row[3], hd = row[4], vd = row[5], inb_h = row[6];
        int h_arrow = inb_h ? OPP[hd] : hd;   /* :683-696 */
        int v_arrow = inb_h ? vd : OPP[vd];   /* :707-708 */
        int hx_end;
        if (hd == DW) { for (int hx = px - 1; hx > minx - 1; hx--) lay_r4(m, hx, py, h_arrow); hx_end = minx; }
        else { for (int hx = px + 1; hx < maxx + 1; hx++) lay_r4(m, hx, py, h_arrow); hx_end = maxx; }
        int vy_end;
        if (vd == DS) { for (int vy = py; vy > miny - 1; vy--) lay_r4(m, px, vy, v_arrow); vy_end = miny; }
        else { for (int vy = py; vy < maxy + 1; vy++) lay_r4(m, px, vy, v_arrow); vy_end = maxy; }
        /* pivot: single outbound arrow (:713-715) */
        m->D[ix(m, px, py)] = dl_one(inb_h ? v_arrow : h_arrow);
        extend_to_road(m, hx_end + DX[hd], py + DY[hd], hd, h_arrow);
        extend_to_road(m, px + DX[vd], vy_end + DY[vd], vd, v_arrow);
        for (int dy = -1; dy <= 1; dy++) /* :731-737 */
            for (int dx = -1; dx <= 1; dx++) {
                if (!dx && !dy) continue;
                int t = typ(m, px + dx, py + dy);
                if (t < 0) continue;
                if (!road_like(t) && t != T_WALL) place_cell(m, px + dx, py + dy, T_SIDEWALK);
            }
    }
    free(visited); free(stack);
    return nb;
}

/* blob table of the current Nothing components (what a host needs to draw a carve tape):
   out rows (minx,miny,maxx,maxy,size,seed_index); returns count (<= cap written). */
int oracle_nothing_blobs(city *m, int32_t *out, int cap) {
    int64_t N = (int64_t)m->c.W * m->c.H;
    int32_t *labels = malloc(N * sizeof(int32_t));
    blob_t *bl;
    int n = label_type(m, T_NOTHING, labels, &bl);
    for (int i = 0; i < n && i < cap; i++) {
        out[6 * i] = bl[i].minx; out[6 * i + 1] = bl[i].miny; out[6 * i + 2] = bl[i].maxx;
        out[6 * i + 3] = bl[i].maxy; out[6 * i + 4] = bl[i].size; out[6 * i + 5] = bl[i].seed;
    }
    free(labels); free(bl);
    return n;
}

/* ===================================================================================== */
/* L3 flood fill + zoning: city_model.py:742-806.  zone_by_block[b-1] in 0..4              */
/* ===================================================================================== */
int oracle_zones(city *m, const uint8_t *zone_by_block, int n_tape) {
    int64_t N = (int64_t)m->c.W * m->c.H;
    int32_t *labels = malloc(N * sizeof(int32_t));
    blob_t *bl;
    int n = label_type(m, T_NOTHING, labels, &bl);
    if (n > n_tape) { free(labels); free(bl); return -1; }
    for (int64_t i = 0; i < N; i++) {
        int b = labels[i];
        if (!b) continue;
        const blob_t *q = &bl[b - 1];
        int t = (q->maxx - q->minx + 1 < 3 || q->maxy - q->miny + 1 < 3) ? T_EMPTY : zone_by_block[b - 1];
        m->T[i] = (uint8_t)t; m->D[i] = 0; m->A[i] &= (AUX_RING | AUX_EVER);
        m->B[i] = b;
    }
    free(labels); free(bl);
    return n;
}

/* ===================================================================================== */
/* L4 dead ends: city_model.py:811-840  (in place, raster sweeps until no change)           */
/* ===================================================================================== */
int oracle_dead_ends(city *m) {
    int W = m->c.W, H = m->c.H, sweeps = 0, changed = 1;
    while (changed) {
        changed = 0; sweeps++;
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                int t = typ(m, x, y);
                if (t != T_R2 && t != T_R3 && t != T_INTER) continue;
                int n = 0;
                for (int k = 0; k < 4; k++) n += road_like(typ(m, x + DX[k], y + DY[k]));
                if (n < 2) { place_cell(m, x, y, T_SIDEWALK); changed = 1; }
            }
    }
    return sweeps;
}

/* ===================================================================================== */
/* L5 R2 -> intersection: city_model.py:842-879                                            */
/* ===================================================================================== */
void oracle_upgrade_r2(city *m) {
    int W = m->c.W, H = m->c.H;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            if (typ(m, x, y) != T_R2) continue;
            if (m->c.ring_type == 2 && m->nh && m->nv) {
                const int32_t *hb0 = m->hb, *hb1 = m->hb + 4 * (m->nh - 1);
                const int32_t *vb0 = m->vb, *vb1 = m->vb + 4 * (m->nv - 1);
                int fy = (hb0[0] <= y && y <= hb0[1]) || (hb1[0] <= y && y <= hb1[1]);
                int fx = (vb0[0] <= x && x <= vb0[1]) || (vb1[0] <= x && x <= vb1[1]);
                if (fy && fx) continue;
            }
            int sw = 0;
            for (int k = 0; k < 4; k++) sw += typ(m, x + DX[k], y + DY[k]) == T_SIDEWALK;
            if (sw >= 2) make_intersection(m, x, y);
        }
}

/* ===================================================================================== */
/* L6 block entrances: city_model.py:884-963, 1783-1796                                    */
/* run_by_block[b-1] = canonical index (runs ordered by min (y,x) cell) among longest runs */
/* entr_out[b-1] = cell index of the entrance or -1                                        */
/* ===================================================================================== */
static int touches_road(const city *m, int x, int y) {
    static const int ox[4] = {1, -1, 0, 0}, oy[4] = {0, 0, 1, -1};
    for (int k = 0; k < 4; k++) {
        int t = typ(m, x + ox[k], y + oy[k]);
        if (t == T_R1 || t == T_R2 || t == T_R3 || t == T_INTER || t == T_HWY_IN || t == T_CR) return 1;
    }
    return 0;
}

static int cmp_xy(const void *a, const void *b) { /* lexicographic (x, y) */
    const int32_t *p = a, *q = b;
    if (p[0] != q[0]) return p[0] - q[0];
    return p[1] - q[1];
}

int oracle_entrances(city *m, int n_blocks, const int32_t *run_by_block, int32_t *entr_out) {
    int W = m->c.W, H = m->c.H;
    int64_t N = (int64_t)W * H;
    /* per-block cell lists via counting sort on B */
    int64_t *start = calloc((size_t)n_blocks + 2, sizeof(int64_t));
    for (int64_t i = 0; i < N; i++) { int b = m->B[i]; if (b > 0 && m->T[i] <= T_EMPTY) start[b + 1]++; }
    for (int b = 1; b <= n_blocks + 1; b++) start[b] += start[b - 1];
    int32_t *cells = malloc((size_t)(start[n_blocks + 1] + 1) * sizeof(int32_t));
    int64_t *fill = malloc(((size_t)n_blocks + 2) * sizeof(int64_t));
    memcpy(fill, start, ((size_t)n_blocks + 2) * sizeof(int64_t));
    for (int64_t i = 0; i < N; i++) { int b = m->B[i]; if (b > 0 && m->T[i] <= T_EMPTY) cells[fill[b]++] = (int32_t)i; }
    int32_t *mark = calloc(N, sizeof(int32_t));   /* ring membership stamp = block id */
    int32_t *runid = calloc(N, sizeof(int32_t));
    int32_t *ring = malloc(N * sizeof(int32_t));
    int32_t *stack = malloc(N * sizeof(int32_t));
    int32_t *run = malloc(2 * N * sizeof(int32_t));
    static const int ox[4] = {1, -1, 0, 0}, oy[4] = {0, 0, 1, -1};
    for (int b = 1; b <= n_blocks; b++) {
        entr_out[b - 1] = -1;
        if (start[b + 1] == start[b]) continue;
        int bt = m->T[cells[start[b]]];
        if (bt > T_OTH) continue; /* only AVAILABLE_CITY_BLOCKS */
        int nr = 0;
        for (int64_t q = start[b]; q < start[b + 1]; q++) {
            int x = cells[q] % W, y = cells[q] / W;
            for (int k = 0; k < 4; k++) {
                int nx = x + ox[k], ny = y + oy[k];
                if (!inb(m, nx, ny)) continue;
                int64_t j = ix(m, nx, ny);
                if (m->B[j] == b && m->T[j] <= T_EMPTY) continue; /* inside region */
                if (mark[j] == b) continue;
                if (!touches_road(m, nx, ny)) continue;
                mark[j] = b; runid[j] = 0; ring[nr++] = (int32_t)j;
            }
        }
        if (!nr) continue;
        if (m->c.entr_level > 0) { /* :911-923 */
            int np = 0;
            int32_t *pref = stack;
            for (int r = 0; r < nr; r++) {
                int x = ring[r] % W, y = ring[r] / W, ok = 0;
                for (int k = 0; k < 4; k++) {
                    int t = typ(m, x + ox[k], y + oy[k]);
                    if (t == T_R1) ok = 1;
                    else if (t == T_R2 && m->c.entr_level < 2) ok = 1;
                }
                if (ok) pref[np++] = ring[r];
            }
            if (np) {
                for (int r = 0; r < nr; r++) mark[ring[r]] = 0;
                for (int r = 0; r < np; r++) { ring[r] = pref[r]; mark[ring[r]] = b; }
                nr = np;
            }
        }
        /* runs = 4-connected components of the ring set; canonical order = by min (y,x) cell,
           i.e. by discovery when the ring is visited in ascending cell index */
        /* sort ring ascending */
        for (int a = 1; a < nr; a++) { int32_t v = ring[a]; int c = a - 1; while (c >= 0 && ring[c] > v) { ring[c + 1] = ring[c]; c--; } ring[c + 1] = v; }
        int nruns = 0, maxlen = 0;
        for (int r = 0; r < nr; r++) {
            if (runid[ring[r]]) continue;
            nruns++;
            int sp = 0, len = 0;
            stack[sp++] = ring[r]; runid[ring[r]] = nruns;
            while (sp) {
                int32_t c = stack[--sp]; len++;
                int x = c % W, y = c / W;
                for (int k = 0; k < 4; k++) {
                    int nx = x + ox[k], ny = y + oy[k];
                    if (!inb(m, nx, ny)) continue;
                    int64_t j = ix(m, nx, ny);
                    if (mark[j] == b && !runid[j]) { runid[j] = nruns; stack[sp++] = (int32_t)j; }
                }
            }
            if (len > maxlen) maxlen = len;
        }
        /* pick the tape's choice among the longest runs */
        int want = run_by_block ? run_by_block[b - 1] : 0, seen = 0, chosen = 0;
        for (int id = 1; id <= nruns && !chosen; id++) {
            int len = 0;
            for (int r = 0; r < nr; r++) len += runid[ring[r]] == id;
            if (len == maxlen) { if (seen == want) chosen = id; seen++; }
        }
        if (!chosen) { /* tape index out of range */
            free(start); free(cells); free(fill); free(mark); free(runid); free(ring); free(stack); free(run);
            return -1;
        }
        int len = 0, same_y = 1, same_x = 1;
        for (int r = 0; r < nr; r++) if (runid[ring[r]] == chosen) { run[2 * len] = ring[r] % W; run[2 * len + 1] = ring[r] / W; len++; }
        for (int r = 1; r < len; r++) { same_y &= run[2 * r + 1] == run[1]; same_x &= run[2 * r] == run[0]; }
        (void)same_x; (void)same_y;
        /* horizontal: sort by x; vertical: by y; else (x,y) -- all equal to (x,y) order here */
        qsort(run, len, 2 * sizeof(int32_t), cmp_xy);
        int cx = run[2 * (len / 2)], cy = run[2 * (len / 2) + 1];
        place_cell(m, cx, cy, T_BE);
        m->B[ix(m, cx, cy)] = b;
        entr_out[b - 1] = (int32_t)ix(m, cx, cy);
    }
    for (int64_t i = 0; i < N; i++) mark[i] = 0;
    free(start); free(cells); free(fill); free(mark); free(runid); free(ring); free(stack); free(run);
    return 0;
}

/* ===================================================================================== */
/* L7 intersection direction validation: city_model.py:969-1012                            */
/* ===================================================================================== */
void oracle_validate_dirs(city *m) {
    int W = m->c.W, H = m->c.H;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            if (typ(m, x, y) != T_INTER) continue;
            uint16_t od = m->D[ix(m, x, y)], nd = 0;
            for (int i = 0; i < dl_len(od); i++) {
                int d = dl_get(od, i);
                int nx = x + DX[d], ny = y + DY[d];
                if (!inb(m, nx, ny)) continue;
                int nt = typ(m, nx, ny);
                if (!road_like(nt)) continue;
                if (nt == T_INTER || dl_has(m->D[ix(m, nx, ny)], d)) nd = dl_append(nd, d);
            }
            m->D[ix(m, x, y)] = nd;
        }
}

/* ===================================================================================== */
/* L8 entrance directions: city_model.py:1035-1070                                         */
/* ===================================================================================== */
void oracle_entrance_dirs(city *m) {
    int W = m->c.W, H = m->c.H;
    static const int ox[4] = {1, -1, 0, 0}, oy[4] = {0, 0, 1, -1};
    static const int need[4] = {DW, DE, DS, DN}; /* arrow on the neighbour INTO the entrance */
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            if (typ(m, x, y) != T_BE) continue;
            uint16_t ed = 0;
            for (int k = 0; k < 4; k++) {
                int nx = x + ox[k], ny = y + oy[k];
                if (!inb(m, nx, ny) || !road_like(typ(m, nx, ny))) continue;
                int64_t j = ix(m, nx, ny);
                if (!dl_has(m->D[j], need[k])) m->D[j] = dl_append(m->D[j], need[k]);
                ed = dl_append(ed, OPP[need[k]]);
            }
            m->D[ix(m, x, y)] = ed;
        }
}

/* ===================================================================================== */
/* L9 traffic lights: city_model.py:1422-1584, cell.py:201-239                             */
/* ===================================================================================== */
typedef struct { int32_t *a; int64_t n, cap; } pairs_t;
static void push_pair(pairs_t *p, int32_t u, int32_t v) {
    if (p->n == p->cap) { p->cap = p->cap ? p->cap * 2 : 4096; p->a = realloc(p->a, (size_t)p->cap * 2 * sizeof(int32_t)); }
    p->a[2 * p->n] = u; p->a[2 * p->n + 1] = v; p->n++;
}

typedef struct {
    city *m;
    pairs_t ctrl, inc, out;
    int32_t *stamp, *queue; int32_t cur;     /* literal BFS scratch */
    int32_t *scc;                             /* SCC ids for fast_reach */
} lights_ctx;

/* cell.py:201-227 literal BFS over the directed arrow graph (any in-bounds neighbour) */
static int leads_to_bfs(lights_ctx *L, int32_t from, int32_t to) {
    city *m = L->m;
    int W = m->c.W;
    int64_t head = 0, tail = 0;
    L->cur++;
    L->queue[tail++] = from; L->stamp[from] = L->cur;
    while (head < tail) {
        int32_t c = L->queue[head++];
        if (c == to) return 1;
        uint16_t d = m->D[c];
        int x = c % W, y = c / W;
        for (int i = 0; i < dl_len(d); i++) {
            int k = dl_get(d, i);
            int nx = x + DX[k], ny = y + DY[k];
            if (!inb(m, nx, ny)) continue;
            int32_t j = (int32_t)ix(m, nx, ny);
            if (L->stamp[j] == L->cur) continue;
            L->stamp[j] = L->cur; L->queue[tail++] = j;
        }
    }
    return 0;
}

/* iterative Tarjan SCC over the arrow graph; scc[i] = component id */
static void build_scc(lights_ctx *L) {
    city *m = L->m;
    int W = m->c.W;
    int64_t N = (int64_t)W * m->c.H;
    int32_t *index = malloc(N * sizeof(int32_t)), *low = malloc(N * sizeof(int32_t));
    int32_t *st = malloc(N * sizeof(int32_t)), *cs = malloc(N * sizeof(int32_t));
    uint8_t *onst = calloc(N, 1), *ci = calloc(N, 1);
    L->scc = malloc(N * sizeof(int32_t));
    for (int64_t i = 0; i < N; i++) { index[i] = -1; L->scc[i] = -1; }
    int32_t idx = 0, ncomp = 0;
    int64_t sp = 0, cp = 0;
    for (int64_t r = 0; r < N; r++) {
        if (index[r] >= 0) continue;
        if (!m->D[r]) { index[r] = low[r] = idx++; L->scc[r] = ncomp++; continue; }
        cs[cp++] = (int32_t)r; index[r] = low[r] = idx++; st[sp++] = (int32_t)r; onst[r] = 1; ci[r] = 0;
        while (cp) {
            int32_t v = cs[cp - 1];
            uint16_t d = m->D[v];
            if (ci[v] < dl_len(d)) {
                int k = dl_get(d, ci[v]++);
                int nx = v % W + DX[k], ny = v / W + DY[k];
                if (!inb(m, nx, ny)) continue;
                int32_t w = (int32_t)ix(m, nx, ny);
                if (index[w] < 0) {
                    index[w] = low[w] = idx++; st[sp++] = w; onst[w] = 1; ci[w] = 0; cs[cp++] = w;
                } else if (onst[w]) { if (index[w] < low[v]) low[v] = index[w]; }
            } else {
                cp--;
                if (cp) { int32_t p = cs[cp - 1]; if (low[v] < low[p]) low[p] = low[v]; }
                if (low[v] == index[v]) {
                    int32_t w;
                    do { w = st[--sp]; onst[w] = 0; L->scc[w] = ncomp; } while (w != v);
                    ncomp++;
                }
            }
        }
    }
    free(index); free(low); free(st); free(cs); free(onst); free(ci);
}

/* same answers as leads_to_bfs: equal SCC => reachable; otherwise BFS that never expands
   inside the target-free part of a component twice (plain BFS, but short-circuited by SCC) */
static int leads_to(lights_ctx *L, int32_t from, int32_t to) {
    if (L->scc) {
        if (L->scc[from] == L->scc[to]) return 1;
        /* Tarjan numbers components in reverse topological order: an edge u->v across
           components has scc[v] < scc[u]; so reachability needs scc[to] < scc[from]. */
        if (L->scc[to] > L->scc[from]) return 0;
    }
    return leads_to_bfs(L, from, to);
}

/* :1528-1548 */
static void scan_reverse(lights_ctx *L, int rx, int ry, uint16_t sdirs, int orig, int32_t tl) {
    city *m = L->m;
    int depth = 0;
    int32_t road = (int32_t)ix(m, rx, ry);
    for (int i = 0; i < dl_len(sdirs); i++) {
        int rd = OPP[dl_get(sdirs, i)];
        int bx = rx + DX[rd], by = ry + DY[rd];
        while (depth <= m->c.tl_range) {
            if (!inb(m, bx, by)) break;
            int32_t nb = (int32_t)ix(m, bx, by);
            if (m->T[nb] == orig && leads_to(L, nb, road)) {
                push_pair(&L->inc, tl, nb);
                m->A[nb] |= AUX_LIGHT;
                bx += DX[rd]; by += DY[rd]; depth++;
            } else break;
        }
    }
}

static int directly_leads_to(const city *m, int32_t from, int32_t to) { /* cell.py:229-239 */
    int W = m->c.W;
    uint16_t d = m->D[from];
    for (int i = 0; i < dl_len(d); i++) {
        int k = dl_get(d, i);
        int nx = from % W + DX[k], ny = from / W + DY[k];
        if (inb(m, nx, ny) && (int32_t)ix(m, nx, ny) == to) return 1;
    }
    return 0;
}

/* :1550-1584 */
static void scan_forward(lights_ctx *L, int rx, int ry, uint16_t sdirs, int orig, int32_t tl, int scan_depth) {
    city *m = L->m;
    int32_t road = (int32_t)ix(m, rx, ry);
    for (int i = 0; i < dl_len(sdirs); i++) {
        int rd = dl_get(sdirs, i);
        int bx = rx + DX[rd], by = ry + DY[rd];
        int depth = scan_depth;
        while (depth <= m->c.tl_range) {
            if (!inb(m, bx, by)) break;
            int32_t c = (int32_t)ix(m, bx, by);
            if (m->T[c] == T_INTER) {
                if (m->c.fwd_mode == 1) { push_pair(&L->out, tl, c); m->A[c] |= AUX_LIGHT; depth++; }
                else if (m->c.fwd_mode == 2) { push_pair(&L->out, tl, c); m->A[c] |= AUX_LIGHT; }
                bx += DX[rd]; by += DY[rd];
            } else if (m->T[c] == orig) {
                if (directly_leads_to(m, c, road)) scan_forward(L, bx, by, sdirs, orig, tl, depth + 1);
                else if (dl_has(m->D[c], rd)) { push_pair(&L->out, tl, c); m->A[c] |= AUX_LIGHT; depth++; }
                bx += DX[rd]; by += DY[rd];
            } else break;
        }
    }
}

/* :1501-1520 */
static void assign_light(lights_ctx *L, int rx, int ry, int ax, int ay, int orig, uint16_t sdirs) {
    city *m = L->m;
    int t = typ(m, ax, ay);
    if (t == T_SIDEWALK) place_cell(m, ax, ay, T_TL);
    else if (t != T_TL) return;
    int32_t tl = (int32_t)ix(m, ax, ay), road = (int32_t)ix(m, rx, ry);
    m->A[road] |= AUX_LIGHT;
    push_pair(&L->ctrl, tl, road);
    scan_reverse(L, rx, ry, sdirs, orig, tl);
    if (m->c.fwd) scan_forward(L, rx, ry, sdirs, orig, tl, 0);
}

static lights_ctx g_lights;

/* returns 0; link tables are fetched with oracle_light_links */
int oracle_lights(city *m) {
    int W = m->c.W, H = m->c.H;
    int64_t N = (int64_t)W * H;
    lights_ctx *L = &g_lights;
    free(L->ctrl.a); free(L->inc.a); free(L->out.a);
    memset(L, 0, sizeof(*L));
    L->m = m;
    L->stamp = calloc(N, sizeof(int32_t));
    L->queue = malloc(N * sizeof(int32_t));
    if (m->c.fast_reach) build_scc(L);
    for (int x = 0; x < W; x++)          /* COLUMN-major (:1432-1433) */
        for (int y = 0; y < H; y++) {
            int t = typ(m, x, y);
            if (!(t == T_R1 || t == T_R2 || t == T_R3 || t == T_HWY_IN || t == T_HWY_OUT || t == T_BE)) continue;
            int64_t i = ix(m, x, y);
            uint16_t rd = m->D[i];
            for (int q = 0; q < dl_len(rd); q++) {
                int d = dl_get(rd, q);
                if (typ(m, x + DX[d], y + DY[d]) != T_INTER) continue;
                int wasbe = (t == T_BE);
                place_cell(m, x, y, T_CR);
                m->D[i] = rd;
                m->A[i] = (uint8_t)((m->A[i] & ~AUX_ORIG) | t);
                if (wasbe) m->B[i] = 0;
                /* valid blocks: cell to the right of every arrow, de-duplicated (:1465-1474) */
                int vx[4], vy[4], nv = 0;
                for (int r = 0; r < dl_len(rd); r++) {
                    int k = RIGHT_OF[dl_get(rd, r)];
                    int bx = x + DX[k], by = y + DY[k], dup = 0;
                    for (int u = 0; u < nv; u++) dup |= (vx[u] == bx && vy[u] == by);
                    if (!dup) { vx[nv] = bx; vy[nv] = by; nv++; }
                }
                for (int u = 0; u < nv; u++) {
                    if (!inb(m, vx[u], vy[u])) continue;
                    int st = typ(m, vx[u], vy[u]);
                    if (st == T_CR || st == t) {
                        /* shares an arrow with the controlled road? (:1483) */
                        if (!(m->D[ix(m, vx[u], vy[u])] & rd & 0xf)) continue;
                        int fx = 2 * vx[u] - x, fy = 2 * vy[u] - y;
                        if (inb(m, fx, fy)) assign_light(L, x, y, fx, fy, t, rd);
                    }
                    assign_light(L, x, y, vx[u], vy[u], t, rd);
                }
                break; /* :1499 */
            }
        }
    free(L->stamp); free(L->queue); free(L->scc);
    L->stamp = L->queue = L->scc = NULL;
    return 0;
}

static int cmp_pair(const void *a, const void *b) {
    const int32_t *p = a, *q = b;
    if (p[0] != q[0]) return p[0] < q[0] ? -1 : 1;
    if (p[1] != q[1]) return p[1] < q[1] ? -1 : 1;
    return 0;
}

/* which: 0 ctrl, 1 incoming, 2 outgoing.  out==NULL -> count only.  Pairs are returned sorted. */
int64_t oracle_light_links(int which, int32_t *out) {
    pairs_t *p = which == 0 ? &g_lights.ctrl : which == 1 ? &g_lights.inc : &g_lights.out;
    if (out) {
        qsort(p->a, (size_t)p->n, 2 * sizeof(int32_t), cmp_pair);
        memcpy(out, p->a, (size_t)p->n * 2 * sizeof(int32_t));
    }
    return p->n;
}

/* ===================================================================================== */
/* L10 derived maps: city_model.py:2151-2199                                               */
/* ===================================================================================== */
void oracle_simple_maps(city *m, uint8_t *is_road, uint8_t *road_type, uint8_t *inter, uint8_t *allowed) {
    int64_t N = (int64_t)m->c.W * m->c.H;
    for (int64_t i = 0; i < N; i++) {
        int t = m->T[i];
        uint8_t rt = 0;
        inter[i] = (t == T_INTER);
        if (t == T_INTER) rt = 1;
        is_road[i] = (uint8_t)road_like(t);
        if (road_like(t)) {
            if (t == T_R1) rt = 1;
            else if (t == T_R2) rt = (m->A[i] & AUX_RING) ? 1 : 2;
            else if (t == T_R3) rt = 3;
            else rt = 1; /* HighwayEntrance / HighwayExit / BlockEntrance / Intersection */
        }
        road_type[i] = rt;
        allowed[i] = (uint8_t)(m->D[i] & 0xf);
    }
}

/* ===================================================================================== */
/* whole pipeline, timing hooks for the CPU baseline                                        */
/* ===================================================================================== */
void oracle_bind(city *m, const ocfg *c, uint8_t *T, uint16_t *D, uint8_t *A, int32_t *B,
                 const int32_t *hb, int nh, const int32_t *vb, int nv) {
    m->c = *c; m->T = T; m->D = D; m->A = A; m->B = B;
    m->hb = hb; m->nh = nh; m->vb = vb; m->nv = nv;
}

size_t oracle_city_sizeof(void) { return sizeof(city); }
