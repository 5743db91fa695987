// k_ccl.cu -- union-find connected-component labelling of one cell type, in raster discovery order.
//
// Replaces the DFS flood fills of _carve_subblock_roads (city_model.py:632-647) and
// _flood_fill_blocks_storing_data (:746-763), and the zone fill of the latter (:768-806).
//
// The reference numbers components by the raster position (y outer, x inner) of their first cell.
// Union-find with "smaller index wins" makes every component's root its minimum raster index, so
// id = 1 + (number of roots before it) -- an exclusive scan over root flags.
//
//   1. init    : label = raster index of the start of the cell's horizontal run inside its warp's
//                32-cell segment (warp ballot -- the run pre-merge costs no memory traffic);
//   2. merge   : union with the cell above where the column adjacency is not already implied by the
//                left neighbours, and across warp-segment seams; lock-free atomicMin hooking;
//   3. flatten + count roots per tile -> single-CTA scan of tile counts -> ranks at roots;
//   4. relabel : label -> id, per-run atomics for bbox / size into the component table.
//
// The label plane IS the caller's block_id plane (int32), so no extra 4 B/cell scratch is needed
// besides the per-root rank (workspace, 4 B/cell, written only at roots).
#include "scan.cuh"

namespace tsim {

constexpr int CCL_TILE = SCAN_TILE;   // cells per CTA in the counting / ranking kernels (256 threads x 8)

__device__ __forceinline__ int uf_find(const int32_t *L, int i) {
    int p = __ldcg(L + i);
    while (p != i) { i = p; p = __ldcg(L + i); }
    return i;
}

__device__ __forceinline__ void uf_union(int32_t *L, int a, int b) {
    for (;;) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }   // a > b: hook a under b
        const int old = atomicMin(L + a, b);
        if (old == a) return;
        a = old;
    }
}

// one thread per cell of the OWNED rows; warp = 32 consecutive x of one row (W padded to 32)
__global__ void __launch_bounds__(256) ccl_init_kernel(int W, int nrows, const uint8_t *__restrict__ T, int32_t *__restrict__ L, int target) {
    const int wpr = (W + 31) >> 5;   // warps per row
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (long long)wpr * nrows) return;
    const int y = (int)(gw / wpr), x = (int)(gw % wpr) * 32 + lane;
    const bool in = x < W;
    const size_t i = (size_t)y * W + x;
    const bool tg = in && T[i] == target;
    const uint32_t mask = __ballot_sync(0xffffffffu, tg);
    if (!in) return;
    if (!tg) { L[i] = -1; return; }
    // start of my run inside this warp segment: one past the highest clear bit below my lane
    const uint32_t below = ~mask & ((1u << lane) - 1u);
    const int start = below ? 32 - __clz(below) : 0;
    L[i] = (int32_t)(i - (lane - start));
}

__global__ void __launch_bounds__(256) ccl_merge_kernel(int W, int nrows, const uint8_t *__restrict__ T, int32_t *L, int target) {
    const long long n = (long long)W * nrows;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (T[i] != target) return;
    const int y = (int)(i / W), x = (int)(i % W);
    const bool left = x > 0 && T[i - 1] == target;
    if (left && (x & 31) == 0) uf_union(L, (int)i, (int)i - 1);            // seam between warp segments
    if (y > 0 && T[i - W] == target) {
        const bool upleft = x > 0 && T[i - W - 1] == target;
        if (!(left && upleft)) uf_union(L, (int)i, (int)(i - W));          // else implied by the left pair
    }
}

// flatten labels and count roots per tile
__global__ void __launch_bounds__(256) ccl_flatten_count_kernel(long long n, int32_t *L, int32_t *tile_count) {
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * CCL_TILE;
    int c = 0;
#pragma unroll
    for (int k = 0; k < CCL_TILE / 256; k++) {
        const long long i = base + k * 256 + threadIdx.x;
        if (i < n) {
            const int l = __ldcg(L + i);
            if (l >= 0) {
                const int r = uf_find(L, l);
                if (r != l) L[i] = r;
                c += (r == (int)i);
            }
        }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_count[blockIdx.x] = s_cnt;
}

// rank roots inside each tile (raster order) and initialise their component-table rows
__global__ void __launch_bounds__(256) ccl_rank_kernel(long long n, const int32_t *__restrict__ L, const int32_t *__restrict__ tile_off,
                                                       int32_t *__restrict__ rank, int32_t *__restrict__ blobs, int cap, int32_t *err) {
    __shared__ int s_warp[8];
    const long long base = (long long)blockIdx.x * CCL_TILE;
    int running = tile_off[blockIdx.x];
    for (int k = 0; k < CCL_TILE / 256; k++) {
        const long long i = base + k * 256 + threadIdx.x;
        const bool root = i < n && L[i] == (int)i;
        const uint32_t m = __ballot_sync(0xffffffffu, root);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) s_warp[w] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { const int c = s_warp[q]; if (q < w) before += c; total += c; }
        if (root) {
            const int id = running + before + __popc(m & ((1u << lane) - 1u)) + 1;   // 1-based
            rank[i] = id;
            if (id <= cap) {
                int32_t *b = blobs + (size_t)(id - 1) * TSIM_BLOB_STRIDE;
                b[0] = 0x7fffffff; b[1] = 0x7fffffff; b[2] = -1; b[3] = -1; b[4] = 0; b[5] = (int32_t)i;
            } else {
                *err = 1;
            }
        }
        running += total;
        __syncthreads();
    }
}

// label -> id, bbox / size by per-run atomics
__global__ void __launch_bounds__(256) ccl_relabel_kernel(int W, int nrows, int y_global0, int32_t *L, const int32_t *__restrict__ rank,
                                                          int32_t *blobs, int cap) {
    const int wpr = (W + 31) >> 5;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (long long)wpr * nrows) return;
    const int y = (int)(gw / wpr), x = (int)(gw % wpr) * 32 + lane;
    const bool in = x < W;
    const size_t i = (size_t)y * W + x;
    const int l = in ? L[i] : -1;
    const int id = l >= 0 ? rank[l] : 0;
    if (in) L[i] = id;
    // runs of equal id inside the warp segment
    const int left_id = __shfl_up_sync(0xffffffffu, id, 1);
    const bool start = id > 0 && (lane == 0 || left_id != id);
    const uint32_t smask = __ballot_sync(0xffffffffu, start || id == 0);   // boundaries (run starts and gaps)
    if (start && id <= cap) {
        const uint32_t after = smask & ~((2u << lane) - 1u);               // next boundary after my lane
        const int end = after ? __ffs(after) - 1 : 32;                     // exclusive lane
        const int len = end - lane;
        int32_t *b = blobs + (size_t)(id - 1) * TSIM_BLOB_STRIDE;
        atomicMin(b + 0, x); atomicMax(b + 2, x + len - 1);
        atomicMin(b + 1, y + y_global0); atomicMax(b + 3, y + y_global0);
        atomicAdd(b + 4, len);
    }
}

// zone fill (city_model.py:768-786): Empty when the bbox is thinner than 3, else the taped zone
__global__ void __launch_bounds__(256) zones_fill_kernel(long long n, uint8_t *T, uint16_t *D, uint8_t *A, const int32_t *__restrict__ B,
                                                         const int32_t *__restrict__ blobs, const int32_t *__restrict__ n_blobs,
                                                         const uint8_t *__restrict__ zone, int n_tape, int32_t *err) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = B[i];
    if (id <= 0) return;
    if (id > n_tape || id > *n_blobs) { *err = 1; return; }
    const int32_t *b = blobs + (size_t)(id - 1) * TSIM_BLOB_STRIDE;
    const bool thin = (b[2] - b[0] + 1 < 3) || (b[3] - b[1] + 1 < 3);
    int t = T_EMPTY;
    if (!thin) { t = zone[id - 1]; if (t > T_OTH) { *err = 2; return; } }
    T[i] = (uint8_t)t; D[i] = 0; A[i] &= (AUX_RING | AUX_EVER);
}

tsim_status label_type(const tsim_cfg *cfg, const uint8_t *T, int32_t *L, int target, int32_t *blobs, int32_t cap, int32_t *n_blobs,
                       void *workspace, size_t ws_bytes, cudaStream_t cs) {
    // single-device labelling works on the owned rows; shards call it per device and merge afterwards
    const int W = cfg->width, nrows = cfg->rows;
    const long long n = (long long)W * nrows;
    const int ntiles = div_up(n, CCL_TILE);
    const size_t need = 256 + (size_t)n * 4 + (size_t)ntiles * 4;
    if (!workspace || ws_bytes < need) { set_error("labelling needs %zu workspace bytes, got %zu", need, ws_bytes); return TSIM_ERR_WORKSPACE; }
    int32_t *err = (int32_t *)workspace;                             // [0] capacity flag
    int32_t *rank = (int32_t *)((char *)workspace + 256);
    int32_t *tile_count = rank + n;
    const size_t off = (size_t)cfg->halo * W;
    const uint8_t *To = T + off;
    int32_t *Lo = L + off;
    TSIM_CUDA(cudaMemsetAsync(err, 0, 4, cs));
    const long long warps = (long long)((W + 31) >> 5) * nrows;
    ccl_init_kernel<<<div_up(warps * 32, 256), 256, 0, cs>>>(W, nrows, To, Lo, target);
    TSIM_LAUNCH_CHECK();
    ccl_merge_kernel<<<div_up(n, 256), 256, 0, cs>>>(W, nrows, To, Lo, target);
    TSIM_LAUNCH_CHECK();
    ccl_flatten_count_kernel<<<ntiles, 256, 0, cs>>>(n, Lo, tile_count);
    TSIM_LAUNCH_CHECK();
    scan_tiles_kernel<<<1, 1024, 0, cs>>>(ntiles, tile_count, n_blobs);
    TSIM_LAUNCH_CHECK();
    ccl_rank_kernel<<<ntiles, 256, 0, cs>>>(n, Lo, tile_count, rank, blobs, cap, err);
    TSIM_LAUNCH_CHECK();
    ccl_relabel_kernel<<<div_up(warps * 32, 256), 256, 0, cs>>>(W, nrows, cfg->row0, Lo, rank, blobs, cap);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}

}  // namespace tsim

using namespace tsim;

extern "C" tsim_status tsim_layout_label_nothing(const tsim_cfg *cfg, const tsim_planes *p, int32_t *blobs, int32_t cap, int32_t *n_blobs,
                                                 void *workspace, size_t ws_bytes, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if (!p || !p->cell_type || !p->block_id || !blobs || !n_blobs || cap < 1) { set_error("tsim_layout_label_nothing: bad arguments"); return TSIM_ERR_CONFIG; }
    return label_type(cfg, p->cell_type, p->block_id, T_NOTHING, blobs, cap, n_blobs, workspace, ws_bytes, (cudaStream_t)stream);
}

extern "C" tsim_status tsim_layout_zones(const tsim_cfg *cfg, const tsim_planes *p, const int32_t *blobs, const int32_t *n_blobs,
                                         const uint8_t *zone_by_block, int32_t n_tape, int32_t *err_flag, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if (!p || !p->cell_type || !p->dirs || !p->aux || !p->block_id || !blobs || !n_blobs || !zone_by_block || !err_flag) {
        set_error("tsim_layout_zones: bad arguments");
        return TSIM_ERR_CONFIG;
    }
    const long long n = (long long)cfg->width * cfg->rows;
    const size_t off = (size_t)cfg->halo * cfg->width;
    zones_fill_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(n, p->cell_type + off, p->dirs + off, p->aux + off, p->block_id + off,
                                                                        blobs, n_blobs, zone_by_block, n_tape, err_flag);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}

extern "C" tsim_status tsim_label_mask(const tsim_cfg *cfg, const uint8_t *mask, int32_t *labels, int32_t *blobs, int32_t cap,
                                       int32_t *n_blobs, void *workspace, size_t ws_bytes, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if (!mask || !labels || !blobs || !n_blobs || cap < 1) { set_error("tsim_label_mask: bad arguments"); return TSIM_ERR_CONFIG; }
    return label_type(cfg, mask, labels, 1, blobs, cap, n_blobs, workspace, ws_bytes, (cudaStream_t)stream);
}
