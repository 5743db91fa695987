/* density_oracle.c -- CPU restatement of CityModel._update_density_map (Simulation/city_model.py:1764-1778).
 * TEST INFRASTRUCTURE ONLY.
 *
 * The reference computes, in float32, uniform_filter(occupancy, 21 x 21, mode='constant') * 441 and the same for
 * is_road_map, and divides (0 where no road is in the window).  The arithmetic lives in SciPy (scipy.ndimage.uniform_filter,
 * version unpinned upstream; 1.18.1 in the build image): one 1-D pass per axis, axis 0 (y) first, each pass a running sum
 * in DOUBLE over the zero-extended line, divided by 21 and stored as float32 (ni_filters.c, NI_UniformFilter1D).
 * Restated here order-free: pass 1 sums 0/1 cells (exact), pass 2 sums 21 float32 values of the form fl32(k / 21) -- at most
 * 34 significant bits, exact in a double whatever the order -- so only the three roundings per cell matter
 * (double / 21 -> float32, twice; * 441 in float32) plus the final float32 division.
 * Pinned against SciPy itself and against the live reference method by tests/test_density_oracle.py. */
#include <stdint.h>
#include <stdlib.h>

#define RADIUS 10              /* Defaults.VEHICLE_AWARENESS_RANGE, config.py:279 */
#define SIZE (2 * RADIUS + 1)

static void box_sum_scaled(int W, int H, const uint8_t *src, float *out) {
    float *col = malloc(sizeof(float) * (size_t)W * H);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {          /* axis 0: window over y */
            int k = 0;
            for (int yy = y - RADIUS; yy <= y + RADIUS; yy++)
                if (yy >= 0 && yy < H) k += src[(size_t)yy * W + x] != 0;
            col[(size_t)y * W + x] = (float)((double)k / (double)SIZE);
        }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {          /* axis 1: window over x */
            double t = 0.0;
            for (int xx = x - RADIUS; xx <= x + RADIUS; xx++)
                if (xx >= 0 && xx < W) t += (double)col[(size_t)y * W + xx];
            const float f = (float)(t / (double)SIZE);
            out[(size_t)y * W + x] = f * (float)(SIZE * SIZE);
        }
    free(col);
}

/* occupancy, is_road: [H][W] 0/1; density: [H][W] float32 */
void oracle_density_map(int W, int H, const uint8_t *occupancy, const uint8_t *is_road, float *density) {
    float *so = malloc(sizeof(float) * (size_t)W * H), *sr = malloc(sizeof(float) * (size_t)W * H);
    box_sum_scaled(W, H, occupancy, so);
    box_sum_scaled(W, H, is_road, sr);
    for (size_t i = 0; i < (size_t)W * H; i++) density[i] = sr[i] > 0.0f ? so[i] / sr[i] : 0.0f;
    free(so); free(sr);
}
