// k_entrances.cu -- _final_place_block_entrances (city_model.py:884-963, _touches_road :1783-1796).
//
// One CTA per zoned block.  The block's bounding box (from the labelling pass) plus a 1-cell rim is
// staged as a shared-memory tile; ring cells that touch a road are marked, split into 4-connected
// runs by a shared-memory union-find (root = min tile index = min (y,x) cell, the canonical run
// order of the tape), the tape-selected longest run is picked and its (x,y)-lexicographic median
// cell becomes the BlockEntrance.  Blocks are independent: an entrance never changes another
// block's ring or its road contacts (BlockEntrance is neither in the region test nor in
// _touches_road's type list).
#include "common.cuh"

namespace tsim {

constexpr int ENT_CAP = 4096;   // tile cells (bbox + rim) a CTA can stage

__device__ __forceinline__ int s_find(const int *lab, int i) {
    int p = lab[i];
    while (p != i) { i = p; p = lab[i]; }
    return i;
}
__device__ __forceinline__ void s_union(int *lab, int a, int b) {
    for (;;) {
        a = s_find(lab, a); b = s_find(lab, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        const int old = atomicMin(lab + a, b);
        if (old == a) return;
        a = old;
    }
}

__global__ void __launch_bounds__(128) entrances_kernel(tsim_cfg c, uint8_t *T, uint16_t *D, uint8_t *A, int32_t *B,
                                                        const int32_t *__restrict__ blobs, const int32_t *__restrict__ n_blobs, int cap_blobs,
                                                        const int32_t *__restrict__ id_base, const int32_t *__restrict__ run_by_block, int n_tape,
                                                        int32_t *entrances, int32_t *err) {
    __shared__ uint8_t s_mark[ENT_CAP];
    __shared__ int s_lab[ENT_CAP];
    __shared__ int s_cnt[ENT_CAP];
    __shared__ int s_max, s_pref, s_chosen, s_warp[4], s_seen;
    const int nb = min(*n_blobs, cap_blobs);
    const int W = c.width, H = c.win_rows;   // window-local rows throughout
    const int base = id_base ? *id_base : 0;
    for (int bk = blockIdx.x; bk < nb; bk += gridDim.x) {
        __syncthreads();
        const int b = bk + 1 + base;         // block id as stored in block_id
        const int32_t *bl = blobs + (size_t)bk * TSIM_BLOB_STRIDE;
        const int root = bl[5];
        if (threadIdx.x == 0) entrances[bk] = -1;
        if (b < 1) continue;                 // cut by the window's lower edge: not owned here
        if (T[root] > T_OTH) continue;   // Empty blocks get no entrance (:902)
        const int by0 = bl[1] - c.win_y0, by1 = bl[3] - c.win_y0;
        if ((by0 == 0 && c.win_y0 > 0) || (by1 == H - 1 && c.win_y0 + H < c.height)) continue;   // cut by a window edge: the owner sees it whole
        if (b > n_tape) { if (threadIdx.x == 0) *err = 1; continue; }
        const int x0 = max(bl[0] - 1, 0), y0 = max(by0 - 1, 0), x1 = min(bl[2] + 1, W - 1), y1 = min(by1 + 1, H - 1);
        const int tw = x1 - x0 + 1, th = y1 - y0 + 1, n = tw * th;
        if (n > ENT_CAP) { if (threadIdx.x == 0) *err = 2; continue; }
        if (threadIdx.x == 0) { s_max = 0; s_pref = 0; s_chosen = -1; s_seen = 0; }
        // 1. mark ring cells that touch a road (:906); bit1 = preferred by the road-level filter (:911-923)
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int x = x0 + i % tw, y = y0 + i / tw;
            const size_t g = (size_t)y * W + x;
            uint8_t m = 0;
            const bool inreg = B[g] == b && in_set(SET_ZONE, T[g]);
            if (!inreg) {
                bool ring = false, road = false, pref = false;
                const int ox[4] = {1, -1, 0, 0}, oy[4] = {0, 0, 1, -1};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int nx = x + ox[k], ny = y + oy[k];
                    if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
                    const size_t j = (size_t)ny * W + nx;
                    const int t = T[j];
                    ring |= (B[j] == b && in_set(SET_ZONE, t));
                    road |= in_set(SET_TOUCH_ROAD, t);
                    pref |= (t == T_R1) || (t == T_R2 && c.block_entrance_road_level < 2);
                }
                if (ring && road) m = 1 | (pref ? 2 : 0);
            }
            s_mark[i] = m; s_lab[i] = i; s_cnt[i] = 0;
            if ((m & 2) && c.block_entrance_road_level > 0) s_pref = 1;
        }
        __syncthreads();
        const uint8_t need = (c.block_entrance_road_level > 0 && s_pref) ? 3 : 1;
        // 2. runs = 4-connected components of the marked cells
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if ((s_mark[i] & need) != need) continue;
            const int lx = i % tw, ly = i / tw;
            if (lx + 1 < tw && (s_mark[i + 1] & need) == need) s_union(s_lab, i, i + 1);
            if (ly + 1 < th && (s_mark[i + tw] & need) == need) s_union(s_lab, i, i + tw);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if ((s_mark[i] & need) != need) continue;
            const int r = s_find(s_lab, i);
            s_lab[i] = r;
            atomicAdd(s_cnt + r, 1);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            if ((s_mark[i] & need) == need && s_lab[i] == i) atomicMax(&s_max, s_cnt[i]);
        __syncthreads();
        const int maxlen = s_max;
        if (maxlen == 0) continue;   // land-locked block (:907-908)
        // 3. the tape's choice among the longest runs, runs ordered by their root (min (y,x) cell)
        const int want = run_by_block[b - 1];
        for (int base = 0; base < n; base += blockDim.x) {
            const int i = base + threadIdx.x;
            const bool cand = i < n && (s_mark[i] & need) == need && s_lab[i] == i && s_cnt[i] == maxlen;
            const uint32_t m = __ballot_sync(0xffffffffu, cand);
            if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = __popc(m);
            __syncthreads();
            int before = s_seen;
            for (int q = 0; q < (threadIdx.x >> 5); q++) before += s_warp[q];
            if (cand && before + __popc(m & ((1u << (threadIdx.x & 31)) - 1u)) == want) s_chosen = i;
            __syncthreads();
            if (threadIdx.x == 0) s_seen += s_warp[0] + s_warp[1] + s_warp[2] + s_warp[3];
            __syncthreads();
        }
        const int chosen = s_chosen;
        if (chosen < 0) { if (threadIdx.x == 0) *err = 3; continue; }
        // 4. (x,y)-lexicographic element len/2 of the chosen run (:949-956): column histogram, then walk
        for (int i = threadIdx.x; i < tw; i += blockDim.x) s_cnt[i] = 0;   // roots' counts no longer needed
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            if ((s_mark[i] & need) == need && s_lab[i] == chosen) atomicAdd(s_cnt + i % tw, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            int k = maxlen / 2, col = 0;
            while (k >= s_cnt[col]) { k -= s_cnt[col]; col++; }
            int cell = -1;
            for (int ly = 0; ly < th; ly++) {
                const int i = ly * tw + col;
                if ((s_mark[i] & need) == need && s_lab[i] == chosen) { if (k == 0) { cell = i; break; } k--; }
            }
            const int x = x0 + cell % tw, y = y0 + cell / tw;
            const size_t g = (size_t)y * W + x;
            // place_cell(..., "BlockEntrance") (:959-962); the highest block id wins a shared cell, as the
            // reference's later place_cell would
            const int prev = atomicMax(B + g, b);
            if (prev <= b) { T[g] = T_BE; D[g] = 0; A[g] &= (AUX_RING | AUX_EVER); }
            entrances[bk] = (int32_t)g;
        }
    }
}

}  // namespace tsim

using namespace tsim;

extern "C" tsim_status tsim_layout_entrances(const tsim_cfg *cfg, const tsim_planes *p, const tsim_blobs *blobs, const int32_t *run_by_block,
                                             int32_t n_tape, int32_t *entrances, int32_t *err_flag, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if ((st = check_blobs(blobs, "tsim_layout_entrances")) != TSIM_OK) return st;
    if (!p || !p->cell_type || !p->dirs || !p->aux || !p->block_id || !run_by_block || !entrances || !err_flag) {
        set_error("tsim_layout_entrances: bad arguments");
        return TSIM_ERR_CONFIG;
    }
    if (n_tape <= 0) return TSIM_OK;
    int grid = blobs->cap < 148 * 64 ? blobs->cap : 148 * 64;
    entrances_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*cfg, p->cell_type, p->dirs, p->aux, p->block_id, blobs->table, blobs->count,
                                                             blobs->cap, blobs->id_base, run_by_block, n_tape, entrances, err_flag);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}
