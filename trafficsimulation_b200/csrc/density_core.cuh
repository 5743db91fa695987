// density_core.cuh -- the three roundings of CityModel._update_density_map (Simulation/city_model.py:1764-1778), shared by
// the device kernels (k_astar.cu) and a host build (tests/native/density_core_host.cpp).
//
// The reference calls scipy.ndimage.uniform_filter(float32 plane, 21 x 21, mode='constant') * 441 for the occupancy and the
// road plane and divides.  SciPy filters one axis at a time (y first), each pass a running sum in double over the zero-
// extended line, / 21, stored as float32.  Pass 1 sums 0/1 cells, pass 2 sums 21 values fl32(k / 21): both sums are exact in
// a double in ANY order (<= 34 significant bits), so a parallel evaluation is bit-exact as long as these roundings are kept:
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

constexpr int DENS_RADIUS = 10;                 // Defaults.VEHICLE_AWARENESS_RANGE (config.py:279)
constexpr int DENS_SIZE = 2 * DENS_RADIUS + 1;

TSIM_HD float dens_after_pass1(int ones_in_window) {          // double running sum / 21 -> float32
#ifdef __CUDA_ARCH__
    return __double2float_rn(__ddiv_rn((double)ones_in_window, (double)DENS_SIZE));
#else
    return (float)((double)ones_in_window / (double)DENS_SIZE);
#endif
}

TSIM_HD float dens_after_pass2(double sum_of_pass1_values) {  // / 21 -> float32, then * 441 in float32
#ifdef __CUDA_ARCH__
    return __fmul_rn(__double2float_rn(__ddiv_rn(sum_of_pass1_values, (double)DENS_SIZE)), (float)(DENS_SIZE * DENS_SIZE));
#else
    const float f = (float)(sum_of_pass1_values / (double)DENS_SIZE);
    return f * (float)(DENS_SIZE * DENS_SIZE);
#endif
}

TSIM_HD float dens_ratio(float sum_occ, float sum_road) {      // np.where(sum_road > 0, sum_occ / sum_road, 0.0)
#ifdef __CUDA_ARCH__
    return sum_road > 0.0f ? __fdiv_rn(sum_occ, sum_road) : 0.0f;
#else
    return sum_road > 0.0f ? sum_occ / sum_road : 0.0f;
#endif
}

}  // namespace tsim
