// k_scan.cu -- exclusive scan building blocks (tile sums -> single-CTA scan of sums -> tile scans)
#include "scan.cuh"

namespace tsim {

__global__ void __launch_bounds__(1024) scan_tiles_kernel(int ntiles, int32_t *tile_count, int32_t *n_out) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < ntiles ? tile_count[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            const int w = s_warp[threadIdx.x];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (threadIdx.x >= o) wi += t; }
            s_warp[threadIdx.x] = wi - w;   // exclusive warp offsets
        }
        __syncthreads();
        const int excl = s_carry + s_warp[threadIdx.x >> 5] + incl - v;
        if (i < ntiles) tile_count[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = s_carry;
}

__global__ void __launch_bounds__(256) tile_sum_kernel(long long n, const int32_t *__restrict__ n_dev, const int32_t *__restrict__ data,
                                                       int32_t *__restrict__ tile_sum) {
    __shared__ int s_sum;
    if (n_dev && *n_dev < n) n = *n_dev;   // the live prefix of a capacity-sized array
    const long long base = (long long)blockIdx.x * SCAN_TILE;
    if (base >= n) { if (threadIdx.x == 0) tile_sum[blockIdx.x] = 0; return; }
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    int c = 0;
#pragma unroll
    for (int k = 0; k < SCAN_TILE / 256; k++) {
        const long long i = base + k * 256 + threadIdx.x;
        if (i < n) c += data[i];
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_sum, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = s_sum;
}

__global__ void __launch_bounds__(256) tile_scan_kernel(long long n, const int32_t *__restrict__ n_dev, int32_t *data,
                                                        const int32_t *__restrict__ tile_off) {
    __shared__ int s_warp[8];
    if (n_dev && *n_dev < n) n = *n_dev;
    const long long base = (long long)blockIdx.x * SCAN_TILE;
    if (base >= n) return;
    int running = tile_off[blockIdx.x];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = 0; k < SCAN_TILE / 256; k++) {
        const long long i = base + k * 256 + threadIdx.x;
        const int v = i < n ? data[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { const int c = s_warp[q]; if (q < w) before += c; total += c; }
        if (i < n) data[i] = running + before + incl - v;
        running += total;
        __syncthreads();
    }
}

tsim_status exclusive_scan_i32(int32_t *data, long long n, int32_t *tmp, int32_t *total_out, cudaStream_t cs, const int32_t *n_dev) {
    if (n <= 0) { TSIM_CUDA(cudaMemsetAsync(total_out, 0, 4, cs)); return TSIM_OK; }
    const int ntiles = div_up(n, SCAN_TILE);
    tile_sum_kernel<<<ntiles, 256, 0, cs>>>(n, n_dev, data, tmp);
    TSIM_LAUNCH_CHECK();
    scan_tiles_kernel<<<1, 1024, 0, cs>>>(ntiles, tmp, total_out);
    TSIM_LAUNCH_CHECK();
    tile_scan_kernel<<<ntiles, 256, 0, cs>>>(n, n_dev, data, tmp);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}

}  // namespace tsim
