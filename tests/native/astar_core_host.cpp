// Host build of the PRODUCT's route-planner core (trafficsimulation_b200/csrc/astar_core.cuh): the same code the CUDA kernel
// runs per thread, driven from pytest through ctypes against the reference's golden vectors (no GPU needed).
#include <cstdlib>
#include <cstring>
#include "../../trafficsimulation_b200/csrc/astar_core.cuh"

// cap <= 0: the capacity tsim_astar_batch uses (half the grid)
// spawn_rank may be null; rank_limit: see tsim_astar_query.spawn_rank_limit
extern "C" int host_astar_ranked(int W, int H, const uint8_t *occ, const uint8_t *stop, const uint8_t *road, const uint8_t *rtype, const uint8_t *adirs,
                                 const double *dens, int sx, int sy, int gx, int gy, int flags, int awareness, int max_steps, int32_t *out,
                                 int out_cap, int cap, const uint8_t *spawn_rank, int rank_limit);

extern "C" int host_astar(int W, int H, const uint8_t *occ, const uint8_t *stop, const uint8_t *road, const uint8_t *rtype, const uint8_t *adirs,
                          const double *dens, int sx, int sy, int gx, int gy, int flags, int awareness, int max_steps, int32_t *out, int out_cap,
                          int cap) {
    return host_astar_ranked(W, H, occ, stop, road, rtype, adirs, dens, sx, sy, gx, gy, flags, awareness, max_steps, out, out_cap, cap, nullptr, 0);
}

extern "C" int host_astar_ranked(int W, int H, const uint8_t *occ, const uint8_t *stop, const uint8_t *road, const uint8_t *rtype, const uint8_t *adirs,
                                 const double *dens, int sx, int sy, int gx, int gy, int flags, int awareness, int max_steps, int32_t *out,
                                 int out_cap, int cap, const uint8_t *spawn_rank, int rank_limit) {
    const size_t n = (size_t)W * H;
    if (cap <= 0) cap = (int)(((n / 2 + 64) + 15) & ~(size_t)15);
    uint16_t *cell = (uint16_t *)malloc(n * sizeof(uint16_t));
    for (size_t i = 0; i < n; i++) cell[i] = tsim::as_pack(occ[i], stop[i], road[i], rtype[i], adirs[i], spawn_rank ? spawn_rank[i] : 0);   // astar_pack_kernel
    tsim::AstarMaps m{W, H, cell, dens};
    tsim::AstarWork w;
    w.dist = (uint32_t *)malloc(n * 4);
    w.heap = (tsim::AsEntry *)aligned_alloc(16, sizeof(tsim::AsEntry) * (size_t)cap);
    w.dir = (int8_t *)malloc(cap);
    w.fov = (uint8_t *)calloc(n, 1);
    w.cap = cap;
    memset(w.dist, 0x3F, n * 4);          // what tsim_astar_batch does with cudaMemsetAsync
    memset(w.heap, 0xA5, sizeof(tsim::AsEntry) * (size_t)cap);   // heap and dir start as garbage on the device
    memset(w.dir, 0x5A, cap);
    const int r = tsim::astar_search(m, sx, sy, gx, gy, flags, awareness, max_steps, w, out, out_cap, rank_limit);
    free(cell); free(w.dist); free(w.heap); free(w.dir); free(w.fov);
    return r;
}

// The windowed search (groundwork, astar_search_window): window = bounding box of start and goal + margin, clipped to the grid.
// Returns what the search returns (AS_ERR_WINDOW = -0x40000001 when it would have left the window); *cells = window area.
extern "C" int host_astar_window(int W, int H, const uint8_t *occ, const uint8_t *stop, const uint8_t *road, const uint8_t *rtype,
                                 const uint8_t *adirs, const double *dens, int sx, int sy, int gx, int gy, int flags, int awareness, int max_steps,
                                 int margin, int32_t *out, int out_cap, int *cells) {
    const size_t n = (size_t)W * H;
    uint16_t *cell = (uint16_t *)malloc(n * sizeof(uint16_t));
    for (size_t i = 0; i < n; i++) cell[i] = tsim::as_pack(occ[i], stop[i], road[i], rtype[i], adirs[i]);
    tsim::AstarMaps m{W, H, cell, dens};
    auto lo = [&](int a, int b) { int v = (a < b ? a : b) - margin; return v < 0 ? 0 : v; };
    auto hi = [&](int a, int b, int lim) { int v = (a > b ? a : b) + margin; return v > lim - 1 ? lim - 1 : v; };
    const int wx0 = lo(sx, gx), wy0 = lo(sy, gy), wx1 = hi(sx, gx, W), wy1 = hi(sy, gy, H);
    const size_t wn = (size_t)(wx1 - wx0 + 1) * (wy1 - wy0 + 1);
    *cells = (int)wn;
    const int cap = (int)(((wn / 2 + 64) + 15) & ~(size_t)15);
    tsim::AstarWork w;
    w.dist = (uint32_t *)malloc(wn * 4);
    w.heap = (tsim::AsEntry *)aligned_alloc(16, sizeof(tsim::AsEntry) * (size_t)cap);
    w.dir = (int8_t *)malloc(cap);
    w.fov = (uint8_t *)calloc(wn, 1);
    w.cap = cap;
    memset(w.dist, 0x3F, wn * 4);
    const int r = tsim::astar_search_window(m, sx, sy, gx, gy, flags, awareness, max_steps, wx0, wy0, wx1, wy1, w, out, out_cap);
    free(cell); free(w.dist); free(w.heap); free(w.dir); free(w.fov);
    return r;
}
