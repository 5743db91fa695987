// common.cuh -- shared device helpers for libtsim (sm_100a)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tsim.h"

namespace tsim {

enum : int {
    T_RES = 0, T_OFF, T_MAR, T_LEI, T_OTH, T_EMPTY, T_NOTHING, T_SIDEWALK, T_WALL,
    T_R1, T_R2, T_R3, T_INTER, T_HWY_IN, T_HWY_OUT, T_TL, T_TL_STOP, T_CR, T_CR_STOP, T_BE
};
enum : int { DN = 0, DE = 1, DS = 2, DW = 3 };

constexpr int AUX_ORIG = TSIM_AUX_ORIG_MASK;
constexpr int AUX_RING = TSIM_AUX_RING;
constexpr int AUX_EVER = TSIM_AUX_EVER_INT;
constexpr int AUX_LIGHT = TSIM_AUX_HAS_LIGHT;

// type-set bitmasks (bit t set <=> cell type t is in the set)
constexpr uint32_t M(int t) { return 1u << t; }
constexpr uint32_t SET_ROAD_LIKE = M(T_R1) | M(T_R2) | M(T_R3) | M(T_INTER) | M(T_HWY_IN) | M(T_HWY_OUT) | M(T_BE);   // config.py:68
constexpr uint32_t SET_ROAD_NO_INT = M(T_R1) | M(T_R2) | M(T_R3) | M(T_HWY_IN) | M(T_HWY_OUT) | M(T_BE);             // config.py:69
constexpr uint32_t SET_REMOVABLE = M(T_R2) | M(T_R3) | M(T_INTER);                                                    // config.py:70
constexpr uint32_t SET_TOUCH_ROAD = M(T_R1) | M(T_R2) | M(T_R3) | M(T_INTER) | M(T_HWY_IN) | M(T_CR);                 // city_model.py:1792-1794
constexpr uint32_t SET_ZONE = M(T_RES) | M(T_OFF) | M(T_MAR) | M(T_LEI) | M(T_OTH) | M(T_EMPTY);

__host__ __device__ __forceinline__ bool in_set(uint32_t set, int t) { return (set >> t) & 1u; }

// direction vectors, opposite, right-of (config.py:64-66)
__host__ __device__ __forceinline__ int dx_of(int d) { return d == DE ? 1 : (d == DW ? -1 : 0); }
__host__ __device__ __forceinline__ int dy_of(int d) { return d == DN ? 1 : (d == DS ? -1 : 0); }
__host__ __device__ __forceinline__ int opp_of(int d) { return (d + 2) & 3; }
__host__ __device__ __forceinline__ int right_of(int d) { return (d + 1) & 3; }

// ordered direction lists packed in u16
__host__ __device__ __forceinline__ int dl_len(uint32_t c) { return (c >> 12) & 7; }
__host__ __device__ __forceinline__ int dl_get(uint32_t c, int i) { return (c >> (4 + 2 * i)) & 3; }
__host__ __device__ __forceinline__ bool dl_has(uint32_t c, int d) { return (c >> d) & 1; }
__host__ __device__ __forceinline__ uint32_t dl_append(uint32_t c, int d) {
    int n = dl_len(c);
    c &= 0x0fffu;
    c |= (1u << d) | ((uint32_t)d << (4 + 2 * n)) | ((uint32_t)(n + 1) << 12);
    return c;
}
__host__ __device__ __forceinline__ uint32_t dl_one(int d) { return (1u << d) | ((uint32_t)d << 4) | (1u << 12); }
constexpr uint32_t DL_NSEW = (0xfu) | (DN << 4) | (DS << 6) | (DE << 8) | (DW << 10) | (4u << 12);   // config.py:62 order N,S,E,W

// line-table entry (see tsim_build_line_table)
__host__ __device__ __forceinline__ bool lt_valid(uint32_t e) { return e & 1u; }
__host__ __device__ __forceinline__ int lt_type(uint32_t e) { return (e >> 1) & 3; }
__host__ __device__ __forceinline__ int lt_dir(uint32_t e) { int d = (e >> 3) & 7; return d == 7 ? -1 : d; }
__host__ __device__ __forceinline__ int lt_off(uint32_t e) { return (e >> 6) & 7; }
__host__ __device__ __forceinline__ int lt_size(uint32_t e) { return (e >> 9) & 7; }
__host__ __device__ __forceinline__ bool lt_in_first(uint32_t e) { return (e >> 12) & 1; }
__host__ __device__ __forceinline__ int lt_off_first(uint32_t e) { return (e >> 13) & 3; }
__host__ __device__ __forceinline__ bool lt_in_last(uint32_t e) { return (e >> 15) & 1; }
__host__ __device__ __forceinline__ int lt_off_last(uint32_t e) { return (e >> 16) & 3; }

struct Geo {   // derived geometry of a cfg
    int W, H, ws, sr, ixmin, ixmax, iymin, iymax;
    __host__ __device__ explicit Geo(const tsim_cfg &c) {
        W = c.width; H = c.height; ws = c.wall_thickness; sr = c.sidewalk_ring_width;
        ixmin = ws + sr; ixmax = W - (ws + sr) - 1; iymin = ws + sr; iymax = H - (ws + sr) - 1;   // city_model.py:91-94
    }
    __host__ __device__ bool inside(int x, int y) const { return x >= ixmin && x <= ixmax && y >= iymin && y <= iymax; }
};

// window of a shard allocation: local row ly holds global row y0 + ly
struct Win {
    int W, LH, y0, Hg;
    __host__ __device__ explicit Win(const tsim_cfg &c) : W(c.width), LH(c.win_rows), y0(c.win_y0), Hg(c.height) {}
    __host__ __device__ long long cells() const { return (long long)W * LH; }
    __host__ __device__ bool full() const { return y0 == 0 && LH == Hg; }
};

void set_error(const char *fmt, ...);
tsim_status check_blobs(const tsim_blobs *b, const char *who);
tsim_status check_cuda(cudaError_t e, const char *what);
tsim_status check_cfg(const tsim_cfg *cfg);

#define TSIM_CUDA(call)                                              \
    do {                                                             \
        tsim_status _s = ::tsim::check_cuda((call), #call);          \
        if (_s != TSIM_OK) return _s;                                \
    } while (0)

void count_launch();   // every kernel launch of the library is counted (tsim_launch_count)
#define TSIM_LAUNCH_CHECK()             \
    do {                                \
        ::tsim::count_launch();         \
        TSIM_CUDA(cudaGetLastError());  \
    } while (0)
#define TSIM_COOP_LAUNCH(kernel, grid, block, args, stream)                                              \
    do {                                                                                                 \
        ::tsim::count_launch();                                                                          \
        TSIM_CUDA(cudaLaunchCooperativeKernel((const void *)(kernel), (grid), (block), (args), 0, (stream))); \
    } while (0)

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace tsim
