// abi.cu -- host-side part of the C ABI that needs no kernels: version, errors, config checks,
// line tables, workspace sizing.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <map>
#include <vector>
#include <tuple>
#include "cells_frame.cuh"

namespace tsim {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

tsim_status check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return TSIM_OK;
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return TSIM_ERR_CUDA;
}

tsim_status check_cfg(const tsim_cfg *c) {
    if (!c) { set_error("cfg is NULL"); return TSIM_ERR_CONFIG; }
    if (c->width < 1 || c->height < 1) {
        set_error("grid %d x %d out of range", c->width, c->height);
        return TSIM_ERR_CONFIG;
    }
    if (c->wall_thickness < 0 || c->sidewalk_ring_width < 0) { set_error("negative frame widths"); return TSIM_ERR_CONFIG; }
    if (c->ring_road_type < 0 || c->ring_road_type > 3) { set_error("ring_road_type %d", c->ring_road_type); return TSIM_ERR_CONFIG; }
    if (c->subblock_road_type < 1 || c->subblock_road_type > 3) { set_error("subblock_road_type %d", c->subblock_road_type); return TSIM_ERR_CONFIG; }
    if (c->win_rows < 1 || c->win_y0 < 0 || c->win_y0 + c->win_rows > c->height) {
        set_error("bad shard window win_y0=%d win_rows=%d height=%d", c->win_y0, c->win_rows, c->height);
        return TSIM_ERR_CONFIG;
    }
    if ((long long)c->width * c->win_rows > 0x7fffffffLL) {
        set_error("window %d x %d out of range (cell indices are int32)", c->width, c->win_rows);
        return TSIM_ERR_CONFIG;
    }
    return TSIM_OK;
}

tsim_status check_blobs(const tsim_blobs *b, const char *who) {
    if (!b || !b->table || !b->count || b->cap < 1) { set_error("%s: bad tsim_blobs", who); return TSIM_ERR_CONFIG; }
    return TSIM_OK;
}

}  // namespace tsim

using namespace tsim;

extern "C" int tsim_version(void) { return TSIM_ABI_VERSION; }

extern "C" const char *tsim_last_error(void) { return g_err; }

extern "C" long long tsim_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// Restates _find_band_covering (city_model.py:1269-1273): the FIRST band in list order that covers
// an index wins; plus membership in bands[0] / bands[-1], which _override_corner_lane_dirs (:519-527)
// and _upgrade_r2_to_intersections (:859-866) test independently of the covering band.
extern "C" tsim_status tsim_build_line_table(const int32_t *bands, int32_t n, int32_t len, uint32_t *out) {
    if (!out || len < 0 || n < 0 || (n > 0 && !bands)) { set_error("tsim_build_line_table: bad arguments"); return TSIM_ERR_CONFIG; }
    for (int i = 0; i < n; i++) {
        const int32_t *b = bands + 4 * i;
        if (b[2] < 1 || b[2] > 3 || b[1] < b[0] || b[1] - b[0] + 1 > 7 || b[3] < -1 || b[3] > 3) {
            set_error("band %d = (%d,%d,%d,%d) is not representable", i, b[0], b[1], b[2], b[3]);
            return TSIM_ERR_CONFIG;
        }
    }
    for (int idx = 0; idx < len; idx++) {
        uint32_t e = 0;
        for (int i = 0; i < n; i++) {
            const int32_t *b = bands + 4 * i;
            if (b[0] <= idx && idx <= b[1]) {
                uint32_t dir = b[3] < 0 ? 7u : (uint32_t)b[3];
                e = 1u | ((uint32_t)b[2] << 1) | (dir << 3) | ((uint32_t)(idx - b[0]) << 6) | ((uint32_t)(b[1] - b[0] + 1) << 9);
                break;
            }
        }
        if (n > 0) {
            const int32_t *f = bands, *l = bands + 4 * (n - 1);
            if (f[0] <= idx && idx <= f[1]) { int o = idx - f[0]; e |= (1u << 12) | ((uint32_t)(o > 2 ? 2 : o) << 13); }
            if (l[0] <= idx && idx <= l[1]) { int o = idx - l[0]; e |= (1u << 15) | ((uint32_t)(o > 2 ? 2 : o) << 16); }
        }
        out[idx] = e;
    }
    return TSIM_OK;
}

// classes of equal (previous, own, next) descriptor triples along one axis
static int axis_classes(const uint32_t *tab, int len, uint8_t *cls, std::map<std::tuple<uint32_t, uint32_t, uint32_t>, int> &ids,
                        std::tuple<uint32_t, uint32_t, uint32_t> *reps) {
    for (int i = 0; i < len; i++) {
        const auto key = std::make_tuple(i > 0 ? tab[i - 1] : 0u, tab[i], i + 1 < len ? tab[i + 1] : 0u);
        auto it = ids.find(key);
        if (it == ids.end()) {
            if (ids.size() >= 255) return -1;
            const int id = (int)ids.size();
            reps[id] = key;
            it = ids.emplace(key, id).first;
        }
        cls[i] = (uint8_t)it->second;
    }
    return (int)ids.size();
}

extern "C" tsim_status tsim_build_class_tables(const tsim_cfg *cfg, const uint32_t *row, const uint32_t *col, uint8_t *row_class,
                                               uint8_t *col_class, uint32_t *lut, int32_t lut_cap, int32_t *n_row, int32_t *n_col) {
    tsim_status s = check_cfg(cfg);
    if (s != TSIM_OK) return s;
    if (!row || !col || !row_class || !col_class || !lut || !n_row || !n_col) { set_error("tsim_build_class_tables: NULL argument"); return TSIM_ERR_CONFIG; }
    std::map<std::tuple<uint32_t, uint32_t, uint32_t>, int> rid, cid;
    std::tuple<uint32_t, uint32_t, uint32_t> rrep[256], crep[256];
    const int nr = axis_classes(row, cfg->height, row_class, rid, rrep), nc = axis_classes(col, cfg->width, col_class, cid, crep);
    if (nr < 0 || nc < 0 || (long long)nr * nc > lut_cap) { set_error("too many band classes for the look-up table"); return TSIM_ERR_CAPACITY; }
    const Geo g(*cfg);
    // any cell of the bulk region stands for all of them (see bulk_* in k_frame_roads.cu)
    const int x = g.ixmin + 6, y = g.iymin + 6;
    for (int a = 0; a < nr; a++)
        for (int b = 0; b < nc; b++) {
            int t; uint32_t d, au;
            frame_roads_cell(*cfg, g, std::get<0>(rrep[a]), std::get<1>(rrep[a]), std::get<2>(rrep[a]), std::get<0>(crep[b]), std::get<1>(crep[b]),
                             std::get<2>(crep[b]), x, y, t, d, au);
            lut[a * nc + b] = (uint32_t)t | (au << 8) | (d << 16);
        }
    *n_row = nr; *n_col = nc;
    return TSIM_OK;
}

extern "C" tsim_status tsim_build_row_patterns(const tsim_cfg *cfg, const uint32_t *row, const uint32_t *col, const uint8_t *row_class, int32_t nr,
                                               uint8_t *pt, uint16_t *pd, uint8_t *pa) {
    tsim_status s = check_cfg(cfg);
    if (s != TSIM_OK) return s;
    if (!row || !col || !row_class || !pt || !pd || !pa || nr < 1 || nr > 256) { set_error("tsim_build_row_patterns: bad arguments"); return TSIM_ERR_CONFIG; }
    const Geo g(*cfg);
    const int W = cfg->width, H = cfg->height;
    const int ybulk = g.iymin + 6;   // any bulk row stands for all of them (see Bulk in k_frame_roads.cu)
    std::vector<int> rep(nr, -1);
    for (int y = 0; y < H; y++) if (rep[row_class[y]] < 0) rep[row_class[y]] = y;
    for (int k = 0; k < nr; k++) {
        const int y = rep[k];
        if (y < 0) { set_error("row class %d has no row", k); return TSIM_ERR_CONFIG; }
        const uint32_t r0 = y > 0 ? row[y - 1] : 0u, r1 = row[y], r2 = y + 1 < H ? row[y + 1] : 0u;
        for (int x = 0; x < W; x++) {
            int t; uint32_t d, a;
            frame_roads_cell(*cfg, g, r0, r1, r2, x > 0 ? col[x - 1] : 0u, col[x], x + 1 < W ? col[x + 1] : 0u, x, ybulk, t, d, a);
            const size_t i = (size_t)k * W + x;
            pt[i] = (uint8_t)t; pd[i] = (uint16_t)d; pa[i] = (uint8_t)a;
        }
    }
    return TSIM_OK;
}

extern "C" tsim_status tsim_workspace_bytes(const tsim_cfg *cfg, size_t *out) {
    tsim_status s = check_cfg(cfg);
    if (s != TSIM_OK) return s;
    if (!out) { set_error("out is NULL"); return TSIM_ERR_CONFIG; }
    size_t cells = (size_t)cfg->width * (size_t)cfg->win_rows;
    // largest user: tsim_layout_lights (reach bits 1 B + per-cell scratch 4 B) and the labelling
    // passes (scan partials); see each pass for its own layout.  16 B per cell + 1 MiB covers all.
    *out = cells * 16 + (1u << 20);
    return TSIM_OK;
}
