/* astar_oracle.c -- CPU restatement of the reference's route planner.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows Simulation/utilities/pathfinding/astar_numba.py: astar_core (:87-236), compute_fov_inplace (:30-50),
 * heap_sift_up (:52-64), heap_sift_down (:66-85), wrapper astar_numba (:240-281); constants from
 * Simulation/config.py (VEHICLE_TURN_PENALTY 10, VEHICLE_CONTRAFLOW_PENALTY 5000, VEHICLE_OBSTACLE_PENALTY_VEHICLE 1000,
 * VEHICLE_OBSTACLE_PENALTY_STOP 500, VEHICLE_ROAD_TYPES_PENALTY_R1/R2/R3 0.5 / 5 / 50.0, VEHICLE_DYNAMIC_PENALTY_SCALE 4.0).
 *
 * Three things of the reference are kept on purpose because they decide WHICH of several equally cheap paths comes out:
 *  - the binary heap orders by f only (strict <) and breaks ties by array position (:52-85);
 *  - dir_arr is BOTH the per-node "no direction yet" array (:121-125) and a per-HEAP-SLOT array that the sifts do not
 *    move with the entry (:130,137,144,229): the turn penalty of a popped node uses whatever direction was last
 *    written to slot 0;
 *  - VEHICLE_ROAD_TYPES_PENALTY_R1 = 0.5 makes Numba type `ng` as float64 from the road-type block on (:207-215); the
 *    comparison `ng < dist[nidx]` sees the fraction, every store into the int32 arrays (:219,225,226) truncates it.
 * Pinned against the live reference (Numba) by tests/test_astar_vs_reference.py and tests/golden/astar_*.npz.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define AS_INF 0x3F3F3F3F

typedef struct oracle_astar_maps {
    int32_t W, H;
    const uint8_t *occupancy, *stop_map, *is_road, *road_type, *allowed_dirs;   /* [H][W] */
    const double *density;                                                       /* [H][W] or NULL */
} oracle_astar_maps;

static const int DX[4] = {0, 1, 0, -1}, DY[4] = {1, 0, -1, 0};   /* NEIGHBOR_DELTAS: N, E, S, W (:9) */

static void sift_up(int32_t *f, int32_t *g, int32_t *s, int32_t *ix, int i) {
    while (i > 0) {
        const int parent = (i - 1) / 2;
        if (f[i] < f[parent]) {
            int32_t t;
            t = f[i]; f[i] = f[parent]; f[parent] = t;
            t = g[i]; g[i] = g[parent]; g[parent] = t;
            t = s[i]; s[i] = s[parent]; s[parent] = t;
            t = ix[i]; ix[i] = ix[parent]; ix[parent] = t;
            i = parent;
        } else break;
    }
}

static void sift_down(int32_t *f, int32_t *g, int32_t *s, int32_t *ix, int size) {
    int idx = 0;
    for (;;) {
        const int left = 2 * idx + 1, right = left + 1;
        int smallest = idx;
        if (left < size && f[left] < f[smallest]) smallest = left;
        if (right < size && f[right] < f[smallest]) smallest = right;
        if (smallest == idx) break;
        int32_t t;
        t = f[idx]; f[idx] = f[smallest]; f[smallest] = t;
        t = g[idx]; g[idx] = g[smallest]; g[smallest] = t;
        t = s[idx]; s[idx] = s[smallest]; s[smallest] = t;
        t = ix[idx]; ix[idx] = ix[smallest]; ix[smallest] = t;
        idx = smallest;
    }
}

static void compute_fov(const oracle_astar_maps *m, int cx, int cy, int awareness, uint8_t *fov) {
    memset(fov, 0, (size_t)m->W * m->H);
    for (int d = 0; d < 4; d++) {
        const int dx = DX[d], dy = DY[d], px = -dy, py = dx;
        for (int off = -awareness + 1; off < awareness; off++) {
            const int x0 = cx + off * px, y0 = cy + off * py;
            int x = x0, y = y0, step = 0;
            while (x >= 0 && x < m->W && y >= 0 && y < m->H && m->is_road[(size_t)y * m->W + x] == 1) {
                fov[(size_t)y * m->W + x] = 1;
                step++;
                x = x0 + dx * step; y = y0 + dy * step;
            }
        }
    }
}

/* Returns the number of path cells written to out (first step first, goal last; the start is not part of it),
 * 0 if there is no path, -1 if out_cap is too small, -2 if the heap outgrew the reference's own arrays. */
int oracle_astar(const oracle_astar_maps *m, int sx, int sy, int gx, int gy, int respect_awareness, int awareness_range,
                 int soft_obstacles, int ignore_flow, int maximum_steps, int32_t *out, int out_cap) {
    const int W = m->W, H = m->H, n = W * H;
    const int start = sy * W + sx, goal = gy * W + gx;
    int32_t *dist = malloc(sizeof(int32_t) * n), *came = malloc(sizeof(int32_t) * n);
    int32_t *f = malloc(sizeof(int32_t) * n), *g = malloc(sizeof(int32_t) * n), *s = malloc(sizeof(int32_t) * n), *ix = malloc(sizeof(int32_t) * n);
    int8_t *dir = malloc(n);
    uint8_t *fov = calloc(n, 1);
    int result = 0;
    for (int i = 0; i < n; i++) { dist[i] = AS_INF; came[i] = -1; dir[i] = -1; }
    dist[start] = 0;
    int heap = 1;
    f[0] = abs(sx - gx) + abs(sy - gy); g[0] = 0; s[0] = 0; ix[0] = start; dir[0] = -1;
    if (respect_awareness) compute_fov(m, sx, sy, awareness_range, fov);
    while (heap > 0) {
        const int32_t cg = g[0], steps = s[0], cur = ix[0];
        const int prev_dir = dir[0];
        heap--;
        if (heap > 0) {
            f[0] = f[heap]; g[0] = g[heap]; s[0] = s[heap]; ix[0] = ix[heap]; dir[0] = dir[heap];
            sift_down(f, g, s, ix, heap);
        }
        if (cur == goal) {
            int len = 0;
            for (int i = cur; i != start; i = came[i]) len++;
            if (len > out_cap) { result = -1; break; }
            int k = len;
            for (int i = cur; i != start; i = came[i]) out[--k] = i;
            result = len;
            break;
        }
        if (cg > dist[cur]) continue;
        const int cx = cur % W, cy = cur / W;
        for (int d = 0; d < 4; d++) {
            const int nx = cx + DX[d], ny = cy + DY[d];
            if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
            const int ns = steps + 1;
            if (ns > maximum_steps) continue;
            const int nidx = ny * W + nx;
            long long ng_i = (long long)cg + 1;
            if (prev_dir != -1 && d != prev_dir) ng_i += 10;                         /* VEHICLE_TURN_PENALTY */
            const int bits = m->allowed_dirs[cur];
            if ((bits & (1 << d)) == 0) {
                if (ignore_flow && m->is_road[nidx] == 1) ng_i += 5000;             /* VEHICLE_CONTRAFLOW_PENALTY */
                else continue;
            }
            const int seen = !respect_awareness || fov[nidx] == 1;
            if (m->occupancy[nidx] == 1 && seen) {
                if (soft_obstacles) {                                                /* VEHICLE_DYNAMIC_PENALTIES_ENABLED */
                    const double dens = m->density ? m->density[nidx] : 0.0;
                    ng_i += (long long)(1000 * (1.0 + 4.0 * dens));
                } else continue;
            }
            if (m->stop_map[nidx] == 1 && seen) {
                if (soft_obstacles) ng_i += 500;                                     /* VEHICLE_OBSTACLE_PENALTY_STOP */
                else continue;
            }
            double ng = (double)ng_i;                                               /* float64 from here on, like Numba's `ng` */
            if (m->is_road[nidx] == 1) {
                const int rt = m->road_type[nidx];
                if (rt == 1) ng += 0.5; else if (rt == 2) ng += 5; else if (rt == 3) ng += 50.0;
            }
            if (ng < (double)dist[nidx]) {
                if (heap >= n) { result = -2; goto done; }
                dist[nidx] = (int32_t)ng;                                            /* truncating stores */
                came[nidx] = cur;
                const int h = abs(nx - gx) + abs(ny - gy);
                f[heap] = (int32_t)(ng + h); g[heap] = (int32_t)ng; s[heap] = ns; ix[heap] = nidx; dir[heap] = (int8_t)d;
                sift_up(f, g, s, ix, heap);
                heap++;
            }
        }
    }
done:
    free(dist); free(came); free(f); free(g); free(s); free(ix); free(dir); free(fov);
    return result;
}
