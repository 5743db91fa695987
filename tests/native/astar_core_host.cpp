// Host build of the PRODUCT's route-planner core (trafficsimulation_b200/csrc/astar_core.cuh): the same code the CUDA kernel
// runs per thread, driven from pytest through ctypes against the reference's golden vectors (no GPU needed).
#include <cstdlib>
#include <cstring>
#include "../../trafficsimulation_b200/csrc/astar_core.cuh"

extern "C" int host_astar(int W, int H, const uint8_t *occ, const uint8_t *stop, const uint8_t *road, const uint8_t *rtype, const uint8_t *adirs,
                          const double *dens, int sx, int sy, int gx, int gy, int flags, int awareness, int max_steps, int32_t *out, int out_cap) {
    const size_t n = (size_t)W * H;
    tsim::AstarMaps m{W, H, occ, stop, road, rtype, adirs, dens};
    tsim::AstarWork w;
    int32_t *ints = (int32_t *)malloc(6 * n * sizeof(int32_t));
    w.dist = ints; w.came = ints + n; w.f = ints + 2 * n; w.g = ints + 3 * n; w.s = ints + 4 * n; w.ix = ints + 5 * n;
    w.dir = (int8_t *)malloc(n); w.fov = (uint8_t *)calloc(n, 1);
    memset(w.dist, 0x3F, n * 4); memset(w.came, 0xFF, n * 4); memset(w.dir, 0xFF, n);   // what tsim_astar_batch does with cudaMemsetAsync
    const int r = tsim::astar_search(m, sx, sy, gx, gy, flags, awareness, max_steps, w, out, out_cap);
    free(ints); free(w.dir); free(w.fov);
    return r;
}
