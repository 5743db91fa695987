// k_scan.cu -- exclusive scan in one pass over the data: chained scan with decoupled look-back.
//
// Every CTA takes a ticket (so that a CTA only ever waits for CTAs that started before it), scans its tile of 4096 elements
// in registers, publishes the tile's sum as a 64-bit status word (flag | value: one store, no fence needed), then walks back
// over the predecessors' status words -- sums until it meets a word that already is an inclusive prefix -- publishes its own
// inclusive prefix and writes its elements.  One read and one write of the data, one launch (the three-kernel form it
// replaces -- tile sums, single-CTA scan of the sums, tile scans -- read the data twice and cost three launches, nine times
// per city).
#include "scan.cuh"

namespace tsim {

typedef unsigned long long u64;
constexpr int SCAN_ITEMS = SCAN_TILE / 256;
#define ST_SUM (1ull << 32)      // status word = flag << 32 | value: the tile's own sum ...
#define ST_PREFIX (2ull << 32)   // ... or the inclusive prefix up to and including the tile

__global__ void __launch_bounds__(256) scan_lookback_kernel(long long n, const int32_t *__restrict__ n_dev, int32_t *data, u64 *status, int32_t *ticket,
                                                            int32_t *total_out, int ntiles) {
    __shared__ int s_tile, s_prefix;
    __shared__ int s_warp[8];
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
    __syncthreads();
    const int tile = s_tile;
    if (n_dev) { const long long live = *n_dev; if (live < n) n = live < 0 ? 0 : live; }
    // tiles beyond the live prefix have nothing to scan and nobody behind them needs their sum: only the live tiles form the chain
    const int live_tiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    if (tile >= live_tiles) { if (tile == 0 && threadIdx.x == 0) *total_out = 0; return; }
    ntiles = live_tiles;
    const long long base = (long long)tile * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    if (base + SCAN_ITEMS <= n && (((uintptr_t)data) & 15) == 0) {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; k++) {
            const int4 q = *reinterpret_cast<const int4 *>(data + base + 4 * k);
            v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) v[k] = base + k < n ? data[base + k] : 0;
    }
    int sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) sum += v[k];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) { const int c = s_warp[q]; if (q < w) before += c; total += c; }
    if (w == 0) {   // the first warp looks back 32 predecessors at a time (one lane each), not one by one
        int prefix = 0;
        if (tile > 0) {
            if (lane == 0) *((volatile u64 *)(status + tile)) = ST_SUM | (u64)(uint32_t)total;
            for (int hi = tile - 1;; hi -= 32) {   // lane l reads tile hi - l; tiles below 0 count as an inclusive prefix of 0
                const int j = hi - lane;
                u64 sw;
                do {   // a predecessor that has not published yet holds an earlier ticket: it is running
                    sw = j >= 0 ? *((volatile u64 *)(status + j)) : ST_PREFIX;
                } while (__any_sync(0xffffffffu, (sw >> 32) == 0));
                const uint32_t pref = __ballot_sync(0xffffffffu, (sw >> 32) == 2);
                const int first = pref ? __ffs(pref) - 1 : 32;   // nearest predecessor that already is an inclusive prefix
                prefix += __reduce_add_sync(0xffffffffu, lane <= first ? (int)(uint32_t)sw : 0);
                if (pref) break;
            }
        }
        if (lane == 0) {
            *((volatile u64 *)(status + tile)) = ST_PREFIX | (u64)(uint32_t)(prefix + total);
            s_prefix = prefix;
            if (tile == ntiles - 1) *total_out = prefix + total;
        }
    }
    __syncthreads();
    int running = s_prefix + before + incl - sum;
    if (base + SCAN_ITEMS <= n && (((uintptr_t)data) & 15) == 0) {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; k++) {
            int4 q;
            q.x = running; running += v[4 * k];
            q.y = running; running += v[4 * k + 1];
            q.z = running; running += v[4 * k + 2];
            q.w = running; running += v[4 * k + 3];
            *reinterpret_cast<int4 *>(data + base + 4 * k) = q;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) data[base + k] = running; running += v[k]; }
    }
}

tsim_status exclusive_scan_i32(int32_t *data, long long n, int32_t *tmp, int32_t *total_out, cudaStream_t cs, const int32_t *n_dev) {
    if (n <= 0) { TSIM_CUDA(cudaMemsetAsync(total_out, 0, 4, cs)); return TSIM_OK; }
    if (((uintptr_t)tmp) & 7) { set_error("exclusive_scan_i32: scratch must be 8-byte aligned"); return TSIM_ERR_CONFIG; }
    const int ntiles = div_up(n, SCAN_TILE);
    u64 *status = reinterpret_cast<u64 *>(tmp);
    int32_t *ticket = reinterpret_cast<int32_t *>(status + ntiles);
    TSIM_CUDA(cudaMemsetAsync(tmp, 0, (size_t)ntiles * 8 + 8, cs));
    scan_lookback_kernel<<<ntiles, 256, 0, cs>>>(n, n_dev, data, status, ticket, total_out, ntiles);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}

}  // namespace tsim
