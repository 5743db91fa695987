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

namespace tsim {

struct AstarBatch {
    AstarMaps m;
    const tsim_astar_query *q;
    int n_queries, max_path;
    int32_t *path_len, *path_cells;
    int32_t *ints;      // [6][n_queries][W*H]: dist, came, f, g, s, ix
    int8_t *dir;        // [n_queries][W*H]
    uint8_t *fov;       // [n_queries][W*H]
    int32_t *err;
};

__global__ void __launch_bounds__(64) astar_kernel(AstarBatch b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n_queries) return;
    const size_t n = (size_t)b.m.W * b.m.H, plane = n * b.n_queries;
    const tsim_astar_query q = b.q[i];
    if (q.sx < 0 || q.sx >= b.m.W || q.gx < 0 || q.gx >= b.m.W || q.sy < 0 || q.sy >= b.m.H || q.gy < 0 || q.gy >= b.m.H || q.awareness_range < 0) {
        *b.err = 52;   // a query outside the grid
        b.path_len[i] = 0;
        return;
    }
    AstarWork w;
    w.dist = b.ints + n * i; w.came = w.dist + plane; w.f = w.came + plane; w.g = w.f + plane; w.s = w.g + plane; w.ix = w.s + plane;
    w.dir = b.dir + n * i; w.fov = b.fov + n * i;
    const int r = astar_search(b.m, q.sx, q.sy, q.gx, q.gy, q.flags, q.awareness_range, q.maximum_steps, w, b.path_cells + (size_t)i * b.max_path,
                               b.max_path);
    if (r == AS_ERR_HEAP) { *b.err = 51; b.path_len[i] = 0; }
    else if (r < 0) { *b.err = 50; b.path_len[i] = r; }   // -(cells needed): the caller's max_path is too small
    else b.path_len[i] = r;
}

static size_t astar_bytes(long long n, int nq) { return (size_t)n * nq * (6 * sizeof(int32_t) + 2) + 256; }

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
    if (((uintptr_t)scratch & 3) != 0) { set_error("tsim_astar_batch: scratch must be 4-byte aligned"); return TSIM_ERR_CONFIG; }
    cudaStream_t cs = (cudaStream_t)stream;
    const size_t plane = (size_t)n * n_queries;
    AstarBatch b;
    b.m = AstarMaps{cfg->width, cfg->height, maps->occupancy, maps->stop_map, maps->is_road_map, maps->road_type_map, maps->allowed_dirs_map,
                    maps->density_map};
    b.q = queries; b.n_queries = n_queries; b.max_path = max_path; b.path_len = path_len; b.path_cells = path_cells; b.err = err_flag;
    b.ints = (int32_t *)scratch;
    b.dir = (int8_t *)(b.ints + 6 * plane);
    b.fov = (uint8_t *)(b.dir + plane);
    TSIM_CUDA(cudaMemsetAsync(b.ints, 0x3F, plane * 4, cs));            // dist = INF = 0x3F3F3F3F (:118)
    TSIM_CUDA(cudaMemsetAsync(b.ints + plane, 0xFF, plane * 4, cs));    // came_from = -1
    TSIM_CUDA(cudaMemsetAsync(b.dir, 0xFF, plane, cs));                 // dir_arr = -1
    TSIM_CUDA(cudaMemsetAsync(b.fov, 0, plane, cs));                    // fov_map = 0
    astar_kernel<<<div_up(n_queries, 64), 64, 0, cs>>>(b);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}
