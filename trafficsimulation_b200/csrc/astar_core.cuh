// astar_core.cuh -- one route query of the reference's planner, exactly: astar_core (utilities/pathfinding/astar_numba.py
// :87-236) with compute_fov_inplace (:30-50) and the array heap (:52-85).  Plain C++ that compiles for the device (k_astar.cu,
// one thread per query) and for the host (tests/native/astar_core_host.cpp runs THIS code against the reference's vectors
// without a GPU).
//
// What makes the result bit-exact rather than "a shortest path":
//  * the open list is the reference's binary heap over parallel arrays, ordered by f alone (strict <), ties resolved by
//    array position -- the same sift sequences give the same pop order;
//  * dir_arr is per NODE at initialisation but per HEAP SLOT afterwards, and the sifts do not move it (:52-85 swap only
//    f, g, s, i): the turn penalty of a popped entry uses the direction last written to slot 0 (:130,137,144,229);
//  * VEHICLE_ROAD_TYPES_PENALTY_R1 = 0.5 turns `ng` into a float64 (:207-215) that is compared as such and truncated on
//    every store (:219,225,226).  For integers a, d: a + 0.5 < d  <=>  a < d, and trunc(a + 0.5) = a, so in integer
//    arithmetic the R1 penalty is exactly 0; R2 = 5 and R3 = 50 are integral.  The only floating-point step left is the
//    dynamic vehicle penalty int(1000 * (1.0 + 4.0 * density)) (:193-196), done in IEEE double with explicit roundings.
#pragma once
#include <stdint.h>

#ifndef TSIM_HD
#ifdef __CUDACC__
#define TSIM_HD __host__ __device__ __forceinline__
#else
#define TSIM_HD inline
#endif
#endif

namespace tsim {

constexpr int AS_INF = 0x3F3F3F3F;                       // :118
constexpr int AS_TURN = 10, AS_CONTRA = 5000, AS_VEHICLE = 1000, AS_STOP = 500, AS_R2 = 5, AS_R3 = 50;   // config.py penalties
constexpr double AS_DYN_SCALE = 4.0;
enum { AS_RESPECT_AWARENESS = 1, AS_SOFT_OBSTACLES = 2, AS_IGNORE_FLOW = 4 };
constexpr int AS_ERR_HEAP = -0x40000000;                 // the open list outgrew the reference's own arrays (W * H entries)

struct AstarMaps {      // [H][W] planes of the model (city_model.py:109-115)
    int W, H;
    const uint8_t *occupancy, *stop_map, *is_road, *road_type, *allowed_dirs;
    const double *density;   // may be null (= 0 everywhere)
};

struct AstarWork {      // per query, W * H entries each.  Expected on entry: dist = AS_INF, came = -1, dir = -1, fov = 0
    int32_t *dist, *came, *f, *g, *s, *ix;
    int8_t *dir;
    uint8_t *fov;
};

TSIM_HD int as_abs(int v) { return v < 0 ? -v : v; }

TSIM_HD void as_swap(const AstarWork &w, int a, int b) {
    int32_t t;
    t = w.f[a]; w.f[a] = w.f[b]; w.f[b] = t;
    t = w.g[a]; w.g[a] = w.g[b]; w.g[b] = t;
    t = w.s[a]; w.s[a] = w.s[b]; w.s[b] = t;
    t = w.ix[a]; w.ix[a] = w.ix[b]; w.ix[b] = t;
}

TSIM_HD long long as_dynamic_penalty(double density) {   // int(p * (1.0 + SCALE * local_density)), :193-196
#ifdef __CUDA_ARCH__
    return (long long)__dmul_rn((double)AS_VEHICLE, __dadd_rn(1.0, __dmul_rn(AS_DYN_SCALE, density)));
#else
    volatile double a = AS_DYN_SCALE * density;          // separate roundings (no contraction into an fma)
    volatile double b = 1.0 + a;
    return (long long)((double)AS_VEHICLE * b);
#endif
}

// Returns the number of cells written to out (first step first, goal last), 0 = no path (or start == goal),
// -(needed) if out_cap is too small, AS_ERR_HEAP on open-list overflow.
TSIM_HD int astar_search(const AstarMaps &m, int sx, int sy, int gx, int gy, int flags, int awareness_range, int maximum_steps,
                         const AstarWork &w, int32_t *out, int out_cap) {
    static const int8_t DXY[8] = {0, 1, 1, 0, 0, -1, -1, 0};   // NEIGHBOR_DELTAS N, E, S, W (:9)
    const int W = m.W, H = m.H, n = W * H;
    const int start = sy * W + sx, goal = gy * W + gx;
    const bool respect = flags & AS_RESPECT_AWARENESS, soft = flags & AS_SOFT_OBSTACLES, ignore_flow = flags & AS_IGNORE_FLOW;
    w.dist[start] = 0;
    int heap = 1;
    w.f[0] = as_abs(sx - gx) + as_abs(sy - gy); w.g[0] = 0; w.s[0] = 0; w.ix[0] = start; w.dir[0] = -1;
    if (respect) {   // compute_fov_inplace :30-50 (fov arrives zeroed)
        for (int d = 0; d < 4; d++) {
            const int dx = DXY[2 * d], dy = DXY[2 * d + 1], px = -dy, py = dx;
            for (int off = -awareness_range + 1; off < awareness_range; off++) {
                const int x0 = sx + off * px, y0 = sy + off * py;
                int x = x0, y = y0, step = 0;
                while (x >= 0 && x < W && y >= 0 && y < H && m.is_road[y * W + x] == 1) {
                    w.fov[y * W + x] = 1;
                    step++;
                    x = x0 + dx * step; y = y0 + dy * step;
                }
            }
        }
    }
    while (heap > 0) {
        const int32_t cg = w.g[0], steps = w.s[0], cur = w.ix[0];
        const int prev_dir = w.dir[0];
        heap--;
        if (heap > 0) {
            w.f[0] = w.f[heap]; w.g[0] = w.g[heap]; w.s[0] = w.s[heap]; w.ix[0] = w.ix[heap]; w.dir[0] = w.dir[heap];
            int idx = 0;                                    // heap_sift_down :66-85
            for (;;) {
                const int left = 2 * idx + 1, right = left + 1;
                int smallest = idx;
                if (left < heap && w.f[left] < w.f[smallest]) smallest = left;
                if (right < heap && w.f[right] < w.f[smallest]) smallest = right;
                if (smallest == idx) break;
                as_swap(w, idx, smallest);
                idx = smallest;
            }
        }
        if (cur == goal) {
            int len = 0;
            for (int i = cur; i != start; i = w.came[i]) len++;
            if (len > out_cap) return -len;
            int k = len;
            for (int i = cur; i != start; i = w.came[i]) out[--k] = i;
            return len;
        }
        if (cg > w.dist[cur]) continue;
        const int cx = cur % W, cy = cur / W;
        const int bits = m.allowed_dirs[cur];
        for (int d = 0; d < 4; d++) {
            const int nx = cx + DXY[2 * d], ny = cy + DXY[2 * d + 1];
            if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
            const int ns = steps + 1;
            if (ns > maximum_steps) continue;
            const int nidx = ny * W + nx;
            long long ng = (long long)cg + 1;
            if (prev_dir != -1 && d != prev_dir) ng += AS_TURN;
            const bool road = m.is_road[nidx] == 1;
            if ((bits & (1 << d)) == 0) {
                if (ignore_flow && road) ng += AS_CONTRA;
                else continue;
            }
            const bool seen = !respect || w.fov[nidx] == 1;
            if (m.occupancy[nidx] == 1 && seen) {
                if (soft) ng += as_dynamic_penalty(m.density ? m.density[nidx] : 0.0);
                else continue;
            }
            if (m.stop_map[nidx] == 1 && seen) {
                if (soft) ng += AS_STOP;
                else continue;
            }
            if (road) {
                const int rt = m.road_type[nidx];
                ng += rt == 2 ? AS_R2 : rt == 3 ? AS_R3 : 0;   // R1's 0.5 never survives a comparison or a store (see the header)
            }
            if (ng < (long long)w.dist[nidx]) {
                if (heap >= n) return AS_ERR_HEAP;
                w.dist[nidx] = (int32_t)ng;
                w.came[nidx] = cur;
                int i = heap;
                w.f[i] = (int32_t)(ng + as_abs(nx - gx) + as_abs(ny - gy)); w.g[i] = (int32_t)ng; w.s[i] = ns; w.ix[i] = nidx;
                w.dir[i] = (int8_t)d;
                while (i > 0) {                             // heap_sift_up :52-64
                    const int parent = (i - 1) / 2;
                    if (w.f[i] < w.f[parent]) { as_swap(w, i, parent); i = parent; }
                    else break;
                }
                heap++;
            }
        }
    }
    return 0;
}

}  // namespace tsim
