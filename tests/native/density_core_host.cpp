// Host build of the product's density arithmetic (trafficsimulation_b200/csrc/density_core.cuh) in the two-pass structure of
// the CUDA kernels (k_astar.cu: density_columns_kernel / density_rows_kernel), for a CPU check against SciPy.
#include <cstdlib>
#include "../../trafficsimulation_b200/csrc/density_core.cuh"

extern "C" void host_density(int W, int H, const uint8_t *occ, const uint8_t *road, float *out) {
    using namespace tsim;
    uint8_t *ko = (uint8_t *)malloc((size_t)W * H), *kr = (uint8_t *)malloc((size_t)W * H);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int a = 0, b = 0;
            for (int yy = y - DENS_RADIUS; yy <= y + DENS_RADIUS; yy++)
                if (yy >= 0 && yy < H) { a += occ[(size_t)yy * W + x] != 0; b += road[(size_t)yy * W + x] != 0; }
            ko[(size_t)y * W + x] = (uint8_t)a; kr[(size_t)y * W + x] = (uint8_t)b;
        }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            double so = 0.0, sr = 0.0;
            for (int xx = x - DENS_RADIUS; xx <= x + DENS_RADIUS; xx++)
                if (xx >= 0 && xx < W) { so += (double)dens_after_pass1(ko[(size_t)y * W + xx]); sr += (double)dens_after_pass1(kr[(size_t)y * W + xx]); }
            out[(size_t)y * W + x] = dens_ratio(dens_after_pass2(so), dens_after_pass2(sr));
        }
    free(ko); free(kr);
}
