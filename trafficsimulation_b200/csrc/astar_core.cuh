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

// Memory layout (a query is a chain of dependent loads, so every access counts):
//  * the five byte maps are packed once per batch into ONE 16-bit word per cell (as_pack): a neighbour costs one load;
//  * a heap entry (f, g, steps, cell) is one 16-byte struct: a sift moves whole entries with single 128-bit accesses, and
//    moves a hole instead of swapping (same final arrangement as the reference's swaps, fewer stores);
//  * dist and came_from share a word: cost in the low 30 bits (AS_INF = 0x3F3F3F3F fits), the direction the cell was
//    entered by in the top 2, which is all came_from is ever used for (the parent is the neighbour that direction came from).
constexpr uint32_t AS_COST_MASK = 0x3FFFFFFFu;
enum { ASC_DIRS = 0xF, ASC_ROAD = 0x10, ASC_RT_SHIFT = 5, ASC_OCC = 0x80, ASC_STOP = 0x100, ASC_RANK_SHIFT = 9, ASC_RANK_MAX = 127 };
// bits 9-15: spawn rank of the vehicle on the cell (0 = it was there before this tick's spawner ran).  The spawns of one tick plan one
// after the other (VehicleAgent.__init__ :72-76 inside the generator's loop): spawn k sees spawns 1..k on the grid, not k+1...  A query
// carries k as its rank limit and treats an occupied cell of a higher rank as free, so all spawns of a tick share ONE batch.

TSIM_HD uint16_t as_pack(uint8_t occupancy, uint8_t stop, uint8_t is_road, uint8_t road_type, uint8_t allowed_dirs, uint8_t spawn_rank = 0) {
    return (uint16_t)((allowed_dirs & ASC_DIRS) | (is_road == 1 ? ASC_ROAD : 0) | ((road_type <= 3 ? road_type : 0) << ASC_RT_SHIFT) |
                      (occupancy == 1 ? ASC_OCC : 0) | (stop == 1 ? ASC_STOP : 0) |
                      ((spawn_rank <= ASC_RANK_MAX ? spawn_rank : ASC_RANK_MAX) << ASC_RANK_SHIFT));
}

struct AstarMaps {      // [H][W]
    int W, H;
    const uint16_t *cell;    // as_pack of the model's planes (city_model.py:109-115)
    const double *density;   // may be null (= 0 everywhere)
};

struct alignas(16) AsEntry { int32_t f, g, s, ix; };

struct AstarWork {      // per query.  Expected on entry: dist = AS_INF (0x3F bytes), fov = 0; heap / dir need no initialisation
    uint32_t *dist;     // [W * H]
    AsEntry *heap;      // [cap]
    int8_t *dir;        // [cap] the reference's dir_arr in its per-heap-slot role
    uint8_t *fov;       // [W * H]
    int cap;            // the reference's arrays hold W * H entries; a city's open list stays below a third of its road cells
};

TSIM_HD int as_abs(int v) { return v < 0 ? -v : v; }

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
                         const AstarWork &w, int32_t *out, int out_cap, int rank_limit = 0) {
    static const int8_t DXY[8] = {0, 1, 1, 0, 0, -1, -1, 0};   // NEIGHBOR_DELTAS N, E, S, W (:9)
    const int W = m.W, H = m.H;
    const int start = sy * W + sx, goal = gy * W + gx;
    const bool respect = flags & AS_RESPECT_AWARENESS, soft = flags & AS_SOFT_OBSTACLES, ignore_flow = flags & AS_IGNORE_FLOW;
    w.dist[start] = 0;
    int heap = 1;
    w.heap[0] = AsEntry{as_abs(sx - gx) + as_abs(sy - gy), 0, 0, start};
    w.dir[0] = -1;                                          // dir_arr[i] = -1 for every node, slot 0 included (:121-125,130)
    if (respect) {   // compute_fov_inplace :30-50 (fov arrives zeroed)
        for (int d = 0; d < 4; d++) {
            const int dx = DXY[2 * d], dy = DXY[2 * d + 1], px = -dy, py = dx;
            for (int off = -awareness_range + 1; off < awareness_range; off++) {
                const int x0 = sx + off * px, y0 = sy + off * py;
                int x = x0, y = y0, step = 0;
                while (x >= 0 && x < W && y >= 0 && y < H && (m.cell[y * W + x] & ASC_ROAD)) {
                    w.fov[y * W + x] = 1;
                    step++;
                    x = x0 + dx * step; y = y0 + dy * step;
                }
            }
        }
    }
    // w.dir needs no initialisation: slot i is written by the push that first grows the heap to i + 1, before any pop reads it
    while (heap > 0) {
        const AsEntry top = w.heap[0];
        const int prev_dir = w.dir[0];
        heap--;
        if (heap > 0) {
            const AsEntry last = w.heap[heap];
            w.dir[0] = w.dir[heap];
            int idx = 0;                                    // heap_sift_down :66-85, moving a hole
            for (;;) {
                const int left = 2 * idx + 1, right = left + 1;
                if (left >= heap) break;
                const AsEntry l = w.heap[left];
                int smallest = idx;
                AsEntry pick = last;
                if (l.f < pick.f) { smallest = left; pick = l; }
                if (right < heap) {
                    const AsEntry r = w.heap[right];
                    if (r.f < pick.f) { smallest = right; pick = r; }
                }
                if (smallest == idx) break;
                w.heap[idx] = pick;
                idx = smallest;
            }
            w.heap[idx] = last;
        }
        const int cur = top.ix;
        if (cur == goal) {
            int len = 0;
            for (int i = cur; i != start;) { const int d = (int)(w.dist[i] >> 30); i -= DXY[2 * d + 1] * W + DXY[2 * d]; len++; }
            if (len > out_cap) return -len;
            int k = len;
            for (int i = cur; i != start;) { out[--k] = i; const int d = (int)(w.dist[i] >> 30); i -= DXY[2 * d + 1] * W + DXY[2 * d]; }
            return len;
        }
        if ((uint32_t)top.g > (w.dist[cur] & AS_COST_MASK)) continue;
        const int cx = cur % W, cy = cur / W;
        const int bits = m.cell[cur] & ASC_DIRS;
        for (int d = 0; d < 4; d++) {
            const int nx = cx + DXY[2 * d], ny = cy + DXY[2 * d + 1];
            if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
            const int ns = top.s + 1;
            if (ns > maximum_steps) continue;
            const int nidx = ny * W + nx;
            const uint32_t c = m.cell[nidx];
            long long ng = (long long)top.g + 1;
            if (prev_dir != -1 && d != prev_dir) ng += AS_TURN;
            if ((bits & (1 << d)) == 0) {
                if (ignore_flow && (c & ASC_ROAD)) ng += AS_CONTRA;
                else continue;
            }
            const bool seen = !respect || w.fov[nidx] == 1;
            if ((c & ASC_OCC) && (int)(c >> ASC_RANK_SHIFT) <= rank_limit && seen) {
                if (soft) ng += as_dynamic_penalty(m.density ? m.density[nidx] : 0.0);
                else continue;
            }
            if ((c & ASC_STOP) && seen) {
                if (soft) ng += AS_STOP;
                else continue;
            }
            if (c & ASC_ROAD) {
                const int rt = (c >> ASC_RT_SHIFT) & 3;
                ng += rt == 2 ? AS_R2 : rt == 3 ? AS_R3 : 0;   // R1's 0.5 never survives a comparison or a store (see the header)
            }
            if (ng < (long long)(w.dist[nidx] & AS_COST_MASK)) {
                if (heap >= w.cap) return AS_ERR_HEAP;
                w.dist[nidx] = (uint32_t)ng | ((uint32_t)d << 30);
                const AsEntry e{(int32_t)(ng + as_abs(nx - gx) + as_abs(ny - gy)), (int32_t)ng, ns, nidx};
                int i = heap;
                w.dir[i] = (int8_t)d;                       // stays with the SLOT (:229), whatever the sift does to the entry
                while (i > 0) {                             // heap_sift_up :52-64, moving a hole
                    const int parent = (i - 1) / 2;
                    const AsEntry pe = w.heap[parent];
                    if (e.f < pe.f) { w.heap[i] = pe; i = parent; }
                    else break;
                }
                w.heap[i] = e;
                heap++;
            }
        }
    }
    return 0;
}

// ---- the same search on a WINDOW of the grid (groundwork for r2; not wired into the kernel yet) --------------------------------
// A query's work arrays cost 13.5 bytes per cell, which on a 2048 x 2048 city is 56 MB per query although a route between two
// cells a block apart touches a few hundred cells.  The search restricted to a window [wx0, wx1] x [wy0, wy1] is IDENTICAL to the
// unrestricted one as long as it never tries to relax a cell outside the window (such a cell would have been pushed, since its
// dist is still INF): every heap operation before that point is the same.  So: run in the window, and the first time a cell
// outside it reaches the relaxation test return AS_ERR_WINDOW -- the caller retries with a larger window or the whole grid.
// w.dist / w.fov hold (wx1 - wx0 + 1) * (wy1 - wy0 + 1) entries; heap entries and the output keep GLOBAL cell indices.
constexpr int AS_ERR_WINDOW = -0x40000001;

TSIM_HD int astar_search_window(const AstarMaps &m, int sx, int sy, int gx, int gy, int flags, int awareness_range, int maximum_steps,
                                int wx0, int wy0, int wx1, int wy1, const AstarWork &w, int32_t *out, int out_cap) {
    static const int8_t DXY[8] = {0, 1, 1, 0, 0, -1, -1, 0};
    const int W = m.W, H = m.H, ww = wx1 - wx0 + 1;
    const int start = sy * W + sx, goal = gy * W + gx;
    const bool respect = flags & AS_RESPECT_AWARENESS, soft = flags & AS_SOFT_OBSTACLES, ignore_flow = flags & AS_IGNORE_FLOW;
    auto local = [&](int x, int y) { return (y - wy0) * ww + (x - wx0); };
    auto inside = [&](int x, int y) { return x >= wx0 && x <= wx1 && y >= wy0 && y <= wy1; };
    if (!inside(sx, sy) || !inside(gx, gy)) return AS_ERR_WINDOW;
    w.dist[local(sx, sy)] = 0;
    int heap = 1;
    w.heap[0] = AsEntry{as_abs(sx - gx) + as_abs(sy - gy), 0, 0, start};
    w.dir[0] = -1;
    if (respect) {   // rays are clipped to the window: a cell outside it is never looked up before the search gives up
        for (int d = 0; d < 4; d++) {
            const int dx = DXY[2 * d], dy = DXY[2 * d + 1], px = -dy, py = dx;
            for (int off = -awareness_range + 1; off < awareness_range; off++) {
                const int x0 = sx + off * px, y0 = sy + off * py;
                int x = x0, y = y0, step = 0;
                while (x >= 0 && x < W && y >= 0 && y < H && (m.cell[y * W + x] & ASC_ROAD)) {
                    if (inside(x, y)) w.fov[local(x, y)] = 1;
                    step++;
                    x = x0 + dx * step; y = y0 + dy * step;
                }
            }
        }
    }
    while (heap > 0) {
        const AsEntry top = w.heap[0];
        const int prev_dir = w.dir[0];
        heap--;
        if (heap > 0) {
            const AsEntry last = w.heap[heap];
            w.dir[0] = w.dir[heap];
            int idx = 0;
            for (;;) {
                const int left = 2 * idx + 1, right = left + 1;
                if (left >= heap) break;
                const AsEntry l = w.heap[left];
                int smallest = idx;
                AsEntry pick = last;
                if (l.f < pick.f) { smallest = left; pick = l; }
                if (right < heap) {
                    const AsEntry r = w.heap[right];
                    if (r.f < pick.f) { smallest = right; pick = r; }
                }
                if (smallest == idx) break;
                w.heap[idx] = pick;
                idx = smallest;
            }
            w.heap[idx] = last;
        }
        const int cur = top.ix;
        const int cx = cur % W, cy = cur / W;
        if (cur == goal) {
            int len = 0;
            for (int x = cx, y = cy; y * W + x != start;) { const int d = (int)(w.dist[local(x, y)] >> 30); x -= DXY[2 * d]; y -= DXY[2 * d + 1]; len++; }
            if (len > out_cap) return -len;
            int k = len;
            for (int x = cx, y = cy; y * W + x != start;) { out[--k] = y * W + x; const int d = (int)(w.dist[local(x, y)] >> 30); x -= DXY[2 * d]; y -= DXY[2 * d + 1]; }
            return len;
        }
        if ((uint32_t)top.g > (w.dist[local(cx, cy)] & AS_COST_MASK)) continue;
        const int bits = m.cell[cur] & ASC_DIRS;
        for (int d = 0; d < 4; d++) {
            const int nx = cx + DXY[2 * d], ny = cy + DXY[2 * d + 1];
            if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
            const int ns = top.s + 1;
            if (ns > maximum_steps) continue;
            const int nidx = ny * W + nx;
            const uint32_t c = m.cell[nidx];
            long long ng = (long long)top.g + 1;
            if (prev_dir != -1 && d != prev_dir) ng += AS_TURN;
            if ((bits & (1 << d)) == 0) {
                if (ignore_flow && (c & ASC_ROAD)) ng += AS_CONTRA;
                else continue;
            }
            const bool in = inside(nx, ny);
            if (!in && respect) return AS_ERR_WINDOW;        // its field-of-view bit is not held
            const bool seen = !respect || w.fov[local(nx, ny)] == 1;
            if ((c & ASC_OCC) && seen) {
                if (soft) ng += as_dynamic_penalty(m.density ? m.density[nidx] : 0.0);
                else continue;
            }
            if ((c & ASC_STOP) && seen) {
                if (soft) ng += AS_STOP;
                else continue;
            }
            if (!in) return AS_ERR_WINDOW;                   // the unrestricted search would push this cell
            if (c & ASC_ROAD) {
                const int rt = (c >> ASC_RT_SHIFT) & 3;
                ng += rt == 2 ? AS_R2 : rt == 3 ? AS_R3 : 0;
            }
            const int li = local(nx, ny);
            if (ng < (long long)(w.dist[li] & AS_COST_MASK)) {
                if (heap >= w.cap) return AS_ERR_HEAP;
                w.dist[li] = (uint32_t)ng | ((uint32_t)d << 30);
                const AsEntry e{(int32_t)(ng + as_abs(nx - gx) + as_abs(ny - gy)), (int32_t)ng, ns, nidx};
                int i = heap;
                w.dir[i] = (int8_t)d;
                while (i > 0) {
                    const int parent = (i - 1) / 2;
                    const AsEntry pe = w.heap[parent];
                    if (e.f < pe.f) { w.heap[i] = pe; i = parent; }
                    else break;
                }
                w.heap[i] = e;
                heap++;
            }
        }
    }
    return 0;   // no route, and the search never left the window: the unrestricted search finds none either
}

}  // namespace tsim
