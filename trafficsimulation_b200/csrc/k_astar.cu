// k_astar.cu -- batched route planning: ONE THREAD PER QUERY runs the reference's planner exactly (astar_core.cuh), thousands
// of queries side by side.  Replaces astar_numba(...) (utilities/pathfinding/astar_numba.py:240-281; the reference's only FFI
// precedent, _astar_cpp.astar_numba, utilities/pathfinding/astar_cpp.cpp:35-50, has the same signature) for callers that
// have many routes to plan at once: the spawner's trips of a tick, every stuck vehicle's re-plan (vehicle_base.py:143-420).
//
// A query is a serial heap search -- its order of expansion IS the result, because ties between equally cheap routes are
// broken by heap position -- so the parallelism is across queries.  Every query owns W * H entries of each work array like
// the reference's wrapper allocates per call (:264-272); they are initialised for the whole batch by four memsets.  The
// work is latency-bound pointer chasing in L2 / HBM; what the GPU adds is that tens of thousands of such chains are in
// flight at once.
#include "common.cuh"
#include "astar_core.cuh"
#include "density_core.cuh"

namespace tsim {

struct AstarBatch {
    AstarMaps m;
    const tsim_astar_query *q;
    int n_queries, max_path, cap;
    int32_t *path_len, *path_cells;
    AsEntry *heap;      // [n_queries][cap]
    uint32_t *dist;     // [n_queries][W*H]
    int8_t *dir;        // [n_queries][cap]
    uint8_t *fov;       // [n_queries][W*H]
    int32_t *err;
};

__global__ void __launch_bounds__(256) astar_pack_kernel(long long n, tsim_astar_maps maps, uint16_t *cell) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        cell[i] = as_pack(maps.occupancy[i], maps.stop_map[i], maps.is_road_map[i], maps.road_type_map[i], maps.allowed_dirs_map[i],
                          maps.spawn_rank ? maps.spawn_rank[i] : (uint8_t)0);
}

// query i of the batch; smem / smem_cap: a shared-memory open list of that many entries (16 B entry + 1 B direction each), or null
__device__ __forceinline__ void astar_one_query(const AstarBatch &b, int i, unsigned char *smem, int smem_cap) {
    const size_t n = (size_t)b.m.W * b.m.H;
    const tsim_astar_query q = b.q[i];
    if (q.sx < 0 || q.sx >= b.m.W || q.gx < 0 || q.gx >= b.m.W || q.sy < 0 || q.sy >= b.m.H || q.gy < 0 || q.gy >= b.m.H || q.awareness_range < 0 ||
        q.spawn_rank_limit < 0) {
        *b.err = 52;   // a query outside the grid
        b.path_len[i] = 0;
        return;
    }
    const AstarWork w{b.dist + n * i, b.heap + (size_t)b.cap * i, b.dir + (size_t)b.cap * i, b.fov + n * i, b.cap};
    int32_t *out = b.path_cells + (size_t)i * b.max_path;
    int r = AS_ERR_HEAP;
    if (smem) {
        const AstarWork ws{w.dist, (AsEntry *)smem, (int8_t *)(smem + sizeof(AsEntry) * (size_t)smem_cap), w.fov, smem_cap};
        r = astar_search(b.m, q.sx, q.sy, q.gx, q.gy, q.flags, q.awareness_range, q.maximum_steps, ws, out, b.max_path, q.spawn_rank_limit);
        if (r == AS_ERR_HEAP) {   // the open list outgrew shared memory: once more from scratch on the arrays in global memory
            for (size_t k = 0; k < n; k++) w.dist[k] = 0x3F3F3F3Fu;
            for (size_t k = 0; k < n; k++) w.fov[k] = 0;
        }
    }
    if (r == AS_ERR_HEAP)
        r = astar_search(b.m, q.sx, q.sy, q.gx, q.gy, q.flags, q.awareness_range, q.maximum_steps, w, out, b.max_path, q.spawn_rank_limit);
    if (r == AS_ERR_HEAP) { *b.err = 51; b.path_len[i] = 0; }
    else if (r < 0) { *b.err = 50; b.path_len[i] = r; }   // -(cells needed): the caller's max_path is too small
    else b.path_len[i] = r;
}

// big batches: a query per THREAD (throughput: tens of thousands of dependent chains in flight)
__global__ void __launch_bounds__(64) astar_kernel(AstarBatch b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n_queries) return;
    astar_one_query(b, i, nullptr, 0);
}

// small batches (a tick's re-plans, replan.py): a query per CTA, one lane working.  A serial search is a chain of dependent
// accesses to its open list; 32 queries in one warp execute each other's divergent paths and share one SM's L1, a query alone on
// its SM partition does neither (2 x faster per batch on the re-planning workload).  SMEM: the open list in shared memory -- no
// gain over L1 (kept, tested, off by default; see the launch site).
template <bool SMEM>
__global__ void __launch_bounds__(32) astar_spread_kernel(AstarBatch b, int smem_cap) {
    extern __shared__ __align__(16) unsigned char astar_smem[];
    if (threadIdx.x != 0) return;
    astar_one_query(b, blockIdx.x, SMEM ? astar_smem : nullptr, smem_cap);
}

// ---- CityModel._update_density_map (city_model.py:1764-1778): what the planner's soft vehicle penalty reads ----------------
// two passes like SciPy's separable filter: ones per 21-cell COLUMN window (exact integers, one byte per cell and plane), then
// per cell the sum over a 21-cell ROW window of the float32 values the first pass would have stored (density_core.cuh)
__global__ void __launch_bounds__(256) density_columns_kernel(int W, int H, const uint8_t *__restrict__ occ, const uint8_t *__restrict__ road,
                                                              uint8_t *__restrict__ k_occ, uint8_t *__restrict__ k_road) {
    const long long n = (long long)W * H;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / W);
        int a = 0, b = 0;
        for (int yy = max(y - DENS_RADIUS, 0); yy <= min(y + DENS_RADIUS, H - 1); yy++) {
            const long long j = i + (long long)(yy - y) * W;
            a += occ[j] != 0; b += road[j] != 0;
        }
        k_occ[i] = (uint8_t)a; k_road[i] = (uint8_t)b;
    }
}

__global__ void __launch_bounds__(256) density_rows_kernel(int W, int H, const uint8_t *__restrict__ k_occ, const uint8_t *__restrict__ k_road,
                                                           float *density32, double *density64) {
    const long long n = (long long)W * H;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        double so = 0.0, sr = 0.0;
        for (int xx = max(x - DENS_RADIUS, 0); xx <= min(x + DENS_RADIUS, W - 1); xx++) {
            so += (double)dens_after_pass1(k_occ[i + (xx - x)]);
            sr += (double)dens_after_pass1(k_road[i + (xx - x)]);
        }
        const float d = dens_ratio(dens_after_pass2(so), dens_after_pass2(sr));
        if (density32) density32[i] = d;
        if (density64) density64[i] = (double)d;
    }
}

// open-list capacity per query: the reference's arrays hold W * H entries; on city maps the list peaks below a third of the
// ROAD cells (~ W * H / 14 on the reference's default city), so half the grid is generous -- and overflow is an error, not UB
constexpr int ASTAR_SMEM_CAP = 4096;      // open-list entries held in shared memory (68 KB per CTA: three queries per SM)
constexpr int ASTAR_SPREAD_MAX = 148 * 8;  // batches up to this size run a query per CTA
static int astar_cap(long long n) { return (int)(((n / 2 + 64) + 15) & ~15LL); }
static size_t round16(size_t v) { return (v + 15) & ~(size_t)15; }
static size_t astar_bytes(long long n, int nq) {
    return round16((size_t)n * 2) + (size_t)nq * (sizeof(AsEntry) * astar_cap(n) + round16((size_t)n * 4) + astar_cap(n) + round16((size_t)n)) + 256;
}

}  // namespace tsim

using namespace tsim;

extern "C" tsim_status tsim_astar_scratch_bytes(const tsim_cfg *cfg, int32_t n_queries, size_t *out) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if (!out || n_queries < 0) { set_error("tsim_astar_scratch_bytes: bad arguments"); return TSIM_ERR_CONFIG; }
    *out = astar_bytes((long long)cfg->width * cfg->height, n_queries);
    return TSIM_OK;
}

extern "C" tsim_status tsim_astar_batch(const tsim_cfg *cfg, const tsim_astar_maps *maps, const tsim_astar_query *queries, int32_t n_queries,
                                        int32_t *path_len, int32_t *path_cells, int32_t max_path, int32_t *err_flag, void *scratch,
                                        size_t scratch_bytes, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if (cfg->win_y0 != 0 || cfg->win_rows != cfg->height) { set_error("tsim_astar_batch plans on the whole grid (no shard windows)"); return TSIM_ERR_UNSUPPORTED; }
    if (!maps || !maps->occupancy || !maps->stop_map || !maps->is_road_map || !maps->road_type_map || !maps->allowed_dirs_map || !queries ||
        !path_len || !path_cells || !err_flag || n_queries < 0 || max_path < 1) {
        set_error("tsim_astar_batch: bad arguments");
        return TSIM_ERR_CONFIG;
    }
    if (n_queries == 0) return TSIM_OK;
    const long long n = (long long)cfg->width * cfg->height;
    if (!scratch || scratch_bytes < astar_bytes(n, n_queries)) {
        set_error("tsim_astar_batch needs %zu scratch bytes for %d queries, got %zu", astar_bytes(n, n_queries), n_queries, scratch_bytes);
        return TSIM_ERR_WORKSPACE;
    }
    if (((uintptr_t)scratch & 15) != 0) { set_error("tsim_astar_batch: scratch must be 16-byte aligned"); return TSIM_ERR_CONFIG; }
    cudaStream_t cs = (cudaStream_t)stream;
    const int cap = astar_cap(n);
    char *p = (char *)scratch;
    uint16_t *cell = (uint16_t *)p;                       p += round16((size_t)n * 2);
    AstarBatch b;
    b.heap = (AsEntry *)p;                                p += sizeof(AsEntry) * (size_t)cap * n_queries;
    b.dist = (uint32_t *)p;                               p += round16((size_t)n * 4) * n_queries;
    b.dir = (int8_t *)p;                                  p += (size_t)cap * n_queries;
    b.fov = (uint8_t *)p;
    b.m = AstarMaps{cfg->width, cfg->height, cell, maps->density_map};
    b.q = queries; b.n_queries = n_queries; b.max_path = max_path; b.cap = cap; b.path_len = path_len; b.path_cells = path_cells; b.err = err_flag;
    astar_pack_kernel<<<(int)(div_up(n, 256) < 1184 ? div_up(n, 256) : 1184), 256, 0, cs>>>(n, *maps, cell);
    TSIM_LAUNCH_CHECK();
    TSIM_CUDA(cudaMemsetAsync(b.dist, 0x3F, (size_t)n * 4 * n_queries, cs));   // dist = INF = 0x3F3F3F3F (:118), no parent direction
    TSIM_CUDA(cudaMemsetAsync(b.fov, 0, (size_t)n * n_queries, cs));           // fov_map = 0
    // launch form: TSIM_ASTAR_MODE = 0 a query per thread, 1 a query per CTA, 2 a query per CTA with the open list in shared memory;
    // default: 1 for batches that leave SMs idle in the per-thread form, 0 beyond.  Measured on the planned_trips workload of bench.py
    // (471 batches of 74 queries on average, profiles/r2_planned_modes.json): 9.98 s per thread, 5.05 s per CTA, 6.21 s per CTA with
    // the open list in shared memory -- a query alone in its CTA already finds its open list in L1, and 68 KB of shared memory per
    // CTA only costs residency
    // (TSIM_ASTAR_SMEM_CAP: a smaller shared-memory open list, for the test of the fall-back to global memory)
    const char *mode_env = getenv("TSIM_ASTAR_MODE"), *cap_env = getenv("TSIM_ASTAR_SMEM_CAP");
    const int mode = mode_env ? atoi(mode_env) : (n_queries <= ASTAR_SPREAD_MAX ? 1 : 0);
    if (mode == 2) {
        int smem_cap = cap < ASTAR_SMEM_CAP ? cap : ASTAR_SMEM_CAP;
        if (cap_env && atoi(cap_env) >= 16 && atoi(cap_env) < smem_cap) smem_cap = atoi(cap_env) & ~15;
        const size_t bytes = (sizeof(AsEntry) + 1) * (size_t)smem_cap;
        static bool attr_set = false;
        if (!attr_set) {
            TSIM_CUDA(cudaFuncSetAttribute(astar_spread_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((sizeof(AsEntry) + 1) * ASTAR_SMEM_CAP)));
            attr_set = true;
        }
        astar_spread_kernel<true><<<n_queries, 32, bytes, cs>>>(b, smem_cap);
    } else if (mode == 1) {
        astar_spread_kernel<false><<<n_queries, 32, 0, cs>>>(b, 0);
    } else {
        astar_kernel<<<div_up(n_queries, 64), 64, 0, cs>>>(b);
    }
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}

extern "C" tsim_status tsim_density_map(const tsim_cfg *cfg, const uint8_t *occupancy, const uint8_t *is_road_map, float *density32,
                                        double *density64, void *scratch, size_t scratch_bytes, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if (cfg->win_y0 != 0 || cfg->win_rows != cfg->height) { set_error("tsim_density_map works on the whole grid (no shard windows)"); return TSIM_ERR_UNSUPPORTED; }
    if (!occupancy || !is_road_map || (!density32 && !density64)) { set_error("tsim_density_map: bad arguments"); return TSIM_ERR_CONFIG; }
    const long long n = (long long)cfg->width * cfg->height;
    if (!scratch || scratch_bytes < (size_t)2 * n) { set_error("tsim_density_map needs %lld scratch bytes, got %zu", 2 * n, scratch_bytes); return TSIM_ERR_WORKSPACE; }
    cudaStream_t cs = (cudaStream_t)stream;
    uint8_t *k_occ = (uint8_t *)scratch, *k_road = k_occ + n;
    const int grid = (int)(div_up(n, 256) < 148 * 8 ? div_up(n, 256) : 148 * 8);
    density_columns_kernel<<<grid, 256, 0, cs>>>(cfg->width, cfg->height, occupancy, is_road_map, k_occ, k_road);
    TSIM_LAUNCH_CHECK();
    density_rows_kernel<<<grid, 256, 0, cs>>>(cfg->width, cfg->height, k_occ, k_road, density32, density64);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}
