// k_entrances.cu -- _final_place_block_entrances (city_model.py:884-963, _touches_road :1783-1796).
//
// One CTA per zoned block.  The block's bounding box (from the labelling pass) plus a 1-cell rim is
// staged as a shared-memory tile; ring cells that touch a road are marked, split into 4-connected
// runs by a shared-memory union-find (root = min tile index = min (y,x) cell, the canonical run
// order of the tape), the tape-selected longest run is picked and its (x,y)-lexicographic median
// cell becomes the BlockEntrance.  Blocks are independent: an entrance never changes another
// block's ring or its road contacts (BlockEntrance is neither in the region test nor in
// _touches_road's type list).
#include <cstdlib>
#include "bitplane.cuh"

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

// big blocks (tile beyond the warp kernel's capacity): one CTA per block, from the list the warp kernel wrote
__global__ void __launch_bounds__(128) entrances_kernel(tsim_cfg c, uint8_t *T, uint16_t *D, uint8_t *A, int32_t *B,
                                                        const int32_t *__restrict__ blobs, const int32_t *__restrict__ n_big,
                                                        const int32_t *__restrict__ big_list, const int32_t *__restrict__ id_base,
                                                        const int32_t *__restrict__ run_by_block, int n_tape, int32_t *entrances, int32_t *err) {
    __shared__ uint8_t s_mark[ENT_CAP];
    __shared__ int s_lab[ENT_CAP];
    __shared__ int s_cnt[ENT_CAP];
    __shared__ int s_max, s_pref, s_chosen, s_warp[4], s_seen;
    const int nbig = *n_big;
    const int W = c.width, H = c.win_rows;   // window-local rows throughout
    const int base = id_base ? *id_base : 0;
    for (int q = blockIdx.x; q < nbig; q += gridDim.x) {
        __syncthreads();
        const int bk = big_list[q];          // the warp kernel already applied the skip rules (:902, window cuts, tape length)
        const int b = bk + 1 + base;         // block id as stored in block_id
        const int32_t *bl = blobs + (size_t)bk * TSIM_BLOB_STRIDE;
        const int by0 = bl[1] - c.win_y0, by1 = bl[3] - c.win_y0;
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

// ------------------------------------------------------------------------------------------------
// Rectangular blocks (size == bounding-box area: almost every block of a city).  Their ring is the four
// sides just outside the box, without the corners, and cells of different sides are never 4-adjacent,
// so the runs are simply the maximal segments of road-touching cells on each side: no tile, no
// union-find, no block_id reads.  The road-touch test runs on BIT-PLANES: TR = cell has a type
// _touches_road accepts (row-major) and its transpose TRt; a side's bit-string is the OR of three
// <= 64-bit extractions (the ring row shifted left and right, and the row outside it).  ONE THREAD per
// block: ~24 word loads from planes that live in L2, then a walk over the handful of runs.
// Anything else (non-rectangular, or a side longer than 64) goes to `gen_list` for the tile kernel.
__device__ __forceinline__ int run_len_at(u64 m, int s) {   // length of the run of ones starting at bit s
    const u64 inv = ~(m >> s);
    const int l = inv ? __ffsll((long long)inv) - 1 : 64;
    return min(l, 64 - s);
}

struct EntPlanes { const u64 *tr, *trt, *pr, *prt; int wp, wpT; };   // pr / prt: preferred road level (NULL when the filter is off)

static_assert(SET_TOUCH_ROAD == (M(T_R1) | M(T_R2) | M(T_R3) | M(T_INTER) | M(T_HWY_IN) | M(T_CR)), "ent_bits_kernel range masks");

// one 16-cell strip per thread -> TR (and PR) words
__global__ void __launch_bounds__(256) ent_bits_kernel(int W, int LH, int wp, const uint8_t *__restrict__ T, int level, u64 *__restrict__ TR,
                                                       u64 *__restrict__ PR) {
    const int spr = wp * 4;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int q = threadIdx.x & 3;
    const long long y = g / spr;
    const int s = (int)(g % spr), x0 = s * 16;
    const bool row_ok = y < LH;
    const uint32_t pref_set = M(T_R1) | (level < 2 ? M(T_R2) : 0u);
    uint32_t mt = 0, mp = 0;
    if (row_ok && x0 < W) {
        const size_t base = (size_t)y * W + x0;
        if ((W & 15) == 0) {
            const uint4 tq = *reinterpret_cast<const uint4 *>(T + base);
            const uint32_t tw[4] = {tq.x, tq.y, tq.z, tq.w};
            mt = strip_range_mask(tw, TypeRanges{T_R1 - 1, T_HWY_IN + 1, T_CR - 1, T_CR + 1});   // SET_TOUCH_ROAD: R1..HighwayEntrance, ControlledRoad
            if (PR) mp = strip_range_mask(tw, TypeRanges{T_R1 - 1, (uint32_t)(level < 2 ? T_R2 + 1 : T_R1 + 1), 0, 0});
        } else {
            for (int k = 0; k < 16 && x0 + k < W; k++) {
                const int t = T[base + k];
                mt |= (uint32_t)in_set(SET_TOUCH_ROAD, t) << k;
                mp |= (uint32_t)in_set(pref_set, t) << k;
            }
        }
    }
    const u64 wt = quad_pack(mt, q);
    if (row_ok && q == 0) TR[(size_t)y * wp + (s >> 2)] = wt;
    if (PR) {   // uniform branch
        const u64 wq = quad_pack(mp, q);
        if (row_ok && q == 0) PR[(size_t)y * wp + (s >> 2)] = wq;
    }
}

// dst ([dst_rows][dst_wp]) = transpose of src ([src_rows][src_wp]); CTA = 4 (bx) x 2 (by) blocks of 64 x 64 bits
__global__ void __launch_bounds__(256) bit_transpose_kernel(const u64 *__restrict__ src, int src_rows, int src_wp, u64 *__restrict__ dst, int dst_rows,
                                                            int dst_wp) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nbx = src_wp, nby = (src_rows + 63) >> 6;
    const int px = (nbx + 3) >> 2;
    const int bx = (blockIdx.x % px) * 4 + (wid & 3), by = (blockIdx.x / px) * 2 + (wid >> 2);
    if (bx < nbx && by < nby) transpose_block(src, src_rows, src_wp, dst, dst_rows, dst_wp, bx, by, lane, false);
}

__global__ void __launch_bounds__(256) entrances_rect_kernel(tsim_cfg c, uint8_t *T, uint16_t *D, uint8_t *A, int32_t *B, EntPlanes ep,
                                                             const int32_t *__restrict__ blobs, const int32_t *__restrict__ n_blobs, int cap_blobs,
                                                             const int32_t *__restrict__ id_base, const int32_t *__restrict__ run_by_block, int n_tape,
                                                             int32_t *entrances, int32_t *n_gen, int32_t *gen_list, int32_t *err) {
    const int nb = min(*n_blobs, cap_blobs);
    const int W = c.width, H = c.win_rows;   // window-local rows throughout
    const int base = id_base ? *id_base : 0;
    const int level = c.block_entrance_road_level;
    for (int bk = blockIdx.x * blockDim.x + threadIdx.x; bk < nb; bk += gridDim.x * blockDim.x) {
        const int b = bk + 1 + base;         // block id as stored in block_id
        const int32_t *bl = blobs + (size_t)bk * TSIM_BLOB_STRIDE;
        const int root = bl[5];
        entrances[bk] = -1;
        if (b < 1) continue;                 // cut by the window's lower edge: not owned here
        if (T[root] > T_OTH) continue;       // Empty blocks get no entrance (:902)
        const int bx0 = bl[0], bx1 = bl[2], by0 = bl[1] - c.win_y0, by1 = bl[3] - c.win_y0;
        if ((by0 == 0 && c.win_y0 > 0) || (by1 == H - 1 && c.win_y0 + H < c.height)) continue;   // cut by a window edge: the owner sees it whole
        if (b > n_tape) { *err = 1; continue; }
        const int w = bx1 - bx0 + 1, h = by1 - by0 + 1;
        if ((long long)w * h != bl[4] || w > 64 || h > 64) { gen_list[atomicAdd(n_gen, 1)] = bk; continue; }
        // sides: 0 bottom (y = by0-1), 1 left (x = bx0-1), 2 right (x = bx1+1), 3 top (y = by1+1); bit j = j-th cell from the low end.
        // A ring cell is marked when one of its three neighbours that are not the block itself has a road type
        // (a zone cell is never a road): its two neighbours along the side and the cell outside.
        auto side = [&](const u64 *pl, const u64 *plt, u64 (&m)[4]) {
            auto rowp = [&](int y) { return (y >= 0 && y < H) ? pl + (size_t)y * ep.wp : nullptr; };
            auto colp = [&](int x) { return (x >= 0 && x < W) ? plt + (size_t)x * ep.wpT : nullptr; };
            const int yb = by0 - 1, yt = by1 + 1, xl = bx0 - 1, xr = bx1 + 1;
            m[0] = yb >= 0 ? extract_bits(rowp(yb), ep.wp, bx0 - 1, w) | extract_bits(rowp(yb), ep.wp, bx0 + 1, w) | extract_bits(rowp(yb - 1), ep.wp, bx0, w) : 0ull;
            m[3] = yt < H ? extract_bits(rowp(yt), ep.wp, bx0 - 1, w) | extract_bits(rowp(yt), ep.wp, bx0 + 1, w) | extract_bits(rowp(yt + 1), ep.wp, bx0, w) : 0ull;
            m[1] = xl >= 0 ? extract_bits(colp(xl), ep.wpT, by0 - 1, h) | extract_bits(colp(xl), ep.wpT, by0 + 1, h) | extract_bits(colp(xl - 1), ep.wpT, by0, h) : 0ull;
            m[2] = xr < W ? extract_bits(colp(xr), ep.wpT, by0 - 1, h) | extract_bits(colp(xr), ep.wpT, by0 + 1, h) | extract_bits(colp(xr + 1), ep.wpT, by0, h) : 0ull;
        };
        u64 mk[4];
        side(ep.tr, ep.trt, mk);
        if (level > 0) {   // :911-923
            u64 pf[4];
            side(ep.pr, ep.prt, pf);
            if ((pf[0] & mk[0]) | (pf[1] & mk[1]) | (pf[2] & mk[2]) | (pf[3] & mk[3])) { mk[0] &= pf[0]; mk[1] &= pf[1]; mk[2] &= pf[2]; mk[3] &= pf[3]; }
        }
        if (!(mk[0] | mk[1] | mk[2] | mk[3])) continue;   // land-locked block (:907-908)
        int maxlen = 0;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            u64 m = mk[s];
            while (m) {
                const int st = __ffsll((long long)m) - 1, l = run_len_at(m, st);
                maxlen = max(maxlen, l);
                m &= ~(((l == 64) ? ~0ull : ((1ull << l) - 1ull)) << st);
            }
        }
        // the tape's choice among the longest runs, in the order of their first (lowest (y, x)) cell:
        // bottom side by x, then left / right runs by their first row (left before right), then the top side
        const int want = run_by_block[b - 1];
        int seen = 0, cs = -1, cst = 0;
        auto visit = [&](int sd, int st, int l) {
            if (l == maxlen) { if (seen == want) { cs = sd; cst = st; } seen++; }
        };
        {
            u64 m = mk[0];
            while (m && cs < 0) { const int st = __ffsll((long long)m) - 1, l = run_len_at(m, st); visit(0, st, l); m &= ~(((l == 64) ? ~0ull : ((1ull << l) - 1ull)) << st); }
            u64 ml = mk[1], mr = mk[2];
            while ((ml | mr) && cs < 0) {
                const int sl = ml ? __ffsll((long long)ml) - 1 : 64, sr = mr ? __ffsll((long long)mr) - 1 : 64;
                if (sl <= sr) { const int l = run_len_at(ml, sl); visit(1, sl, l); ml &= ~(((l == 64) ? ~0ull : ((1ull << l) - 1ull)) << sl); }
                else { const int l = run_len_at(mr, sr); visit(2, sr, l); mr &= ~(((l == 64) ? ~0ull : ((1ull << l) - 1ull)) << sr); }
            }
            m = mk[3];
            while (m && cs < 0) { const int st = __ffsll((long long)m) - 1, l = run_len_at(m, st); visit(3, st, l); m &= ~(((l == 64) ? ~0ull : ((1ull << l) - 1ull)) << st); }
        }
        if (cs < 0) { *err = 3; continue; }
        // element len/2 of the run sorted by x (horizontal) or y (vertical) (:949-956)
        const int off = cst + maxlen / 2;
        const int ex = cs == 1 ? bx0 - 1 : (cs == 2 ? bx1 + 1 : bx0 + off), ey = cs == 0 ? by0 - 1 : (cs == 3 ? by1 + 1 : by0 + off);
        const size_t g = (size_t)ey * W + ex;
        // place_cell(..., "BlockEntrance") (:959-962); the highest block id wins a shared cell, as the
        // reference's later place_cell would
        const int prev = atomicMax(B + g, b);
        if (prev <= b) { T[g] = T_BE; D[g] = 0; A[g] &= (AUX_RING | AUX_EVER); }
        entrances[bk] = (int32_t)g;
    }
}

// ------------------------------------------------------------------------------------------------
// Small blocks (the bounding box grown by 2 cells holds at most ENT_WCAP cells -- every block of a
// carved city): ONE WARP per block, no CTA-wide barriers.  The tile is staged once as a 3-bit code per
// cell (member of this block / a type _touches_road accepts / a type the road-level filter prefers), so
// global memory is read once per tile cell; marks, run union-find, run lengths, the taped choice and the
// median all work on the warp's shared-memory slice.
constexpr int ENT_WARPS = 4;

// ENT_WCAP: tile cells a warp can stage; ENT_LCAP: marked (ring and road) cells of a tile.  <768, 384>: 37 KB of shared memory per CTA, 6
// CTAs per SM; <512, 256>: 25 KB, 9 CTAs per SM -- the kernel is a chain of short warp-wide phases and lives on warps in flight; a block
// whose tile does not fit goes to the CTA kernel either way.
template <int ENT_WCAP, int ENT_LCAP>
__global__ void __launch_bounds__(32 * ENT_WARPS) entrances_warp_kernel(tsim_cfg c, uint8_t *T, uint16_t *D, uint8_t *A, int32_t *B,
                                                                        const int32_t *__restrict__ blobs, const int32_t *__restrict__ n_gen,
                                                                        const int32_t *__restrict__ gen_list, const int32_t *__restrict__ id_base,
                                                                        const int32_t *__restrict__ run_by_block, int n_tape, int32_t *entrances,
                                                                        int32_t *n_big, int32_t *big_list, int32_t *err) {
    __shared__ uint8_t s_code_all[ENT_WARPS][ENT_WCAP];
    __shared__ uint8_t s_mark_all[ENT_WARPS][ENT_WCAP];
    __shared__ int s_lab_all[ENT_WARPS][ENT_WCAP];
    __shared__ int s_cnt_all[ENT_WARPS][ENT_WCAP];
    __shared__ uint32_t s_list_all[ENT_WARPS][ENT_LCAP];   // marked cells in tile order: index | lx << 10 | ly << 20
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint8_t *s_code = s_code_all[wid], *s_mark = s_mark_all[wid];
    int *s_lab = s_lab_all[wid], *s_cnt = s_cnt_all[wid];
    uint32_t *s_list = s_list_all[wid];
    const int ngen = *n_gen;
    const int W = c.width, H = c.win_rows;   // window-local rows throughout
    const int base = id_base ? *id_base : 0;
    const int level = c.block_entrance_road_level;
    const int gw = blockIdx.x * ENT_WARPS + wid, nwarps = gridDim.x * ENT_WARPS;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int q = gw; q < ngen; q += nwarps) {
        __syncwarp();
        const int bk = gen_list[q];          // the rectangle kernel already applied the skip rules (:902, window cuts, tape length)
        const int b = bk + 1 + base;         // block id as stored in block_id
        const int32_t *bl = blobs + (size_t)bk * TSIM_BLOB_STRIDE;
        const int by0 = bl[1] - c.win_y0, by1 = bl[3] - c.win_y0;
        const int x0 = max(bl[0] - 2, 0), y0 = max(by0 - 2, 0), x1 = min(bl[2] + 2, W - 1), y1 = min(by1 + 2, H - 1);
        const int tw = x1 - x0 + 1, th = y1 - y0 + 1, n = tw * th;
        if (n > ENT_WCAP) { if (lane == 0) big_list[atomicAdd(n_big, 1)] = bk; continue; }
        // 0. stage the tile: bit0 member of the block, bit1 _touches_road type, bit2 preferred road level
        for (int i0 = lane; i0 < n; i0 += 128) {   // four cells per lane with both planes' loads in flight together
            size_t g[4];
            int t[4], bid[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = min(i0 + 32 * u, n - 1), ly = i / tw;
                g[u] = (size_t)(y0 + ly) * W + x0 + (i - ly * tw);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) { t[u] = T[g[u]]; bid[u] = B[g[u]]; }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (i0 + 32 * u >= n) continue;
                uint8_t code;
                if (in_set(SET_ZONE, t[u])) code = (bid[u] == b) ? 1 : 0;
                else code = (in_set(SET_TOUCH_ROAD, t[u]) ? 2 : 0) | (((t[u] == T_R1) || (t[u] == T_R2 && level < 2)) ? 4 : 0);
                s_code[i0 + 32 * u] = code;
            }
        }
        __syncwarp();
        // 1. ring cells that touch a road (:906); bit1 = preferred by the road-level filter (:911-923); compacted in tile order
        int nl = 0;
        bool pref_any = false;
        {
            int ly = lane / tw, lx = lane - ly * tw;
            for (int i0 = 0; i0 < n; i0 += 32) {
                const int i = i0 + lane;
                uint8_t m = 0;
                if (i < n && !(s_code[i] & 1)) {
                    uint8_t acc = 0;
                    if (lx + 1 < tw) acc |= s_code[i + 1];
                    if (lx > 0) acc |= s_code[i - 1];
                    if (ly + 1 < th) acc |= s_code[i + tw];
                    if (ly > 0) acc |= s_code[i - tw];
                    if ((acc & 3) == 3) m = 1 | ((acc & 4) ? 2 : 0);
                }
                if (i < n) s_mark[i] = m;
                const uint32_t bal = __ballot_sync(0xffffffffu, m != 0);
                if (m) {
                    const int p = nl + __popc(bal & lt_mask);
                    if (p < ENT_LCAP) s_list[p] = (uint32_t)i | ((uint32_t)lx << 10) | ((uint32_t)ly << 20);
                    s_lab[i] = i; s_cnt[i] = 0;
                    pref_any |= (m & 2) != 0;
                }
                nl += __popc(bal);
                lx += 32;
                while (lx >= tw) { lx -= tw; ly++; }
            }
        }
        if (nl == 0) continue;               // land-locked block (:907-908)
        if (nl > ENT_LCAP) { if (lane == 0) big_list[atomicAdd(n_big, 1)] = bk; continue; }
        pref_any = __any_sync(0xffffffffu, pref_any);
        __syncwarp();
        const uint8_t need = (level > 0 && pref_any) ? 3 : 1;
        // 2. runs = 4-connected components of the marked cells
        for (int p = lane; p < nl; p += 32) {
            const uint32_t e = s_list[p];
            const int i = e & 1023, lx = (e >> 10) & 1023, ly = e >> 20;
            if ((s_mark[i] & need) != need) continue;
            if (lx + 1 < tw && (s_mark[i + 1] & need) == need) s_union(s_lab, i, i + 1);
            if (ly + 1 < th && (s_mark[i + tw] & need) == need) s_union(s_lab, i, i + tw);
        }
        __syncwarp();
        for (int p = lane; p < nl; p += 32) {
            const int i = s_list[p] & 1023;
            if ((s_mark[i] & need) != need) continue;
            const int r = s_find(s_lab, i);
            s_lab[i] = r;
            atomicAdd(s_cnt + r, 1);
        }
        __syncwarp();
        int maxlen = 0;
        for (int p = lane; p < nl; p += 32) {
            const int i = s_list[p] & 1023;
            if ((s_mark[i] & need) == need && s_lab[i] == i) maxlen = max(maxlen, s_cnt[i]);
        }
        maxlen = __reduce_max_sync(0xffffffffu, maxlen);
        if (maxlen == 0) continue;           // no ring cell passes the road-level filter
        // 3. the tape's choice among the longest runs, runs ordered by their root (min (y,x) cell)
        const int want = run_by_block[b - 1];
        int chosen = -1, seen = 0;
        for (int p0 = 0; p0 < nl; p0 += 32) {
            const int p = p0 + lane;
            const int i = p < nl ? (int)(s_list[p] & 1023) : 0;
            const bool cand = p < nl && (s_mark[i] & need) == need && s_lab[i] == i && s_cnt[i] == maxlen;
            const uint32_t m = __ballot_sync(0xffffffffu, cand);
            const int k = want - seen;
            if (k >= 0 && k < __popc(m)) {
                uint32_t mm = m;
                for (int j = 0; j < k; j++) mm &= mm - 1;
                chosen = (int)(s_list[p0 + __ffs(mm) - 1] & 1023);
                break;
            }
            seen += __popc(m);
        }
        if (chosen < 0) { if (lane == 0) *err = 3; continue; }
        // 4. (x,y)-lexicographic element len/2 of the chosen run (:949-956): column histogram, then walk
        __syncwarp();
        for (int i = lane; i < tw; i += 32) s_cnt[i] = 0;   // roots' counts no longer needed
        __syncwarp();
        for (int p = lane; p < nl; p += 32) {
            const uint32_t e = s_list[p];
            const int i = e & 1023;
            if ((s_mark[i] & need) == need && s_lab[i] == chosen) atomicAdd(s_cnt + ((e >> 10) & 1023), 1);
        }
        __syncwarp();
        if (lane == 0) {
            int k = maxlen / 2, col = 0;
            while (k >= s_cnt[col]) { k -= s_cnt[col]; col++; }
            int cell = -1;
            for (int ly = 0; ly < th; ly++) {
                const int i = ly * tw + col;
                if ((s_mark[i] & need) == need && s_lab[i] == chosen) { if (k == 0) { cell = i; break; } k--; }
            }
            const size_t g = (size_t)(y0 + cell / tw) * W + x0 + cell % tw;
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
                                             int32_t n_tape, int32_t *entrances, int32_t *err_flag, void *workspace, size_t ws_bytes, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if ((st = check_blobs(blobs, "tsim_layout_entrances")) != TSIM_OK) return st;
    if (!p || !p->cell_type || !p->dirs || !p->aux || !p->block_id || !run_by_block || !entrances || !err_flag) {
        set_error("tsim_layout_entrances: bad arguments");
        return TSIM_ERR_CONFIG;
    }
    const int W = cfg->width, LH = cfg->win_rows, wp = div_up(W, 64), wpT = div_up(LH, 64);
    const bool pref = cfg->block_entrance_road_level > 0;
    const size_t list_bytes = ((size_t)blobs->cap * 4 + 255) & ~(size_t)255;
    const size_t plane_bytes = ((size_t)wp * LH * 8 + 255) & ~(size_t)255, planeT_bytes = ((size_t)wpT * W * 8 + 255) & ~(size_t)255;
    const size_t need = 256 + 2 * list_bytes + (pref ? 2 : 1) * (plane_bytes + planeT_bytes);
    if (!workspace || ws_bytes < need) { set_error("tsim_layout_entrances needs %zu workspace bytes, got %zu", need, ws_bytes); return TSIM_ERR_WORKSPACE; }
    if (n_tape <= 0) return TSIM_OK;
    cudaStream_t cs = (cudaStream_t)stream;
    char *wsp = (char *)workspace;
    int32_t *n_big = (int32_t *)wsp, *n_gen = n_big + 1, *big_list = (int32_t *)(wsp + 256);
    int32_t *gen_list = (int32_t *)(wsp + 256 + list_bytes);
    u64 *TR = (u64 *)(wsp + 256 + 2 * list_bytes), *TRt = (u64 *)((char *)TR + plane_bytes);
    u64 *PR = pref ? (u64 *)((char *)TRt + planeT_bytes) : nullptr, *PRt = pref ? (u64 *)((char *)PR + plane_bytes) : nullptr;
    TSIM_CUDA(cudaMemsetAsync(n_big, 0, 8, cs));
    ent_bits_kernel<<<div_up((long long)wp * LH * 4, 256), 256, 0, cs>>>(W, LH, wp, p->cell_type, cfg->block_entrance_road_level, TR, PR);
    TSIM_LAUNCH_CHECK();
    const int tgrid = div_up(wp, 4) * div_up(wpT, 2);
    bit_transpose_kernel<<<tgrid, 256, 0, cs>>>(TR, LH, wp, TRt, W, wpT);
    TSIM_LAUNCH_CHECK();
    if (pref) {
        bit_transpose_kernel<<<tgrid, 256, 0, cs>>>(PR, LH, wp, PRt, W, wpT);
        TSIM_LAUNCH_CHECK();
    }
    EntPlanes ep{TR, TRt, PR, PRt, wp, wpT};
    const int rgrid = div_up(blobs->cap, 256) < 148 * 8 ? div_up(blobs->cap, 256) : 148 * 8;
    entrances_rect_kernel<<<rgrid, 256, 0, cs>>>(*cfg, p->cell_type, p->dirs, p->aux, p->block_id, ep, blobs->table, blobs->count, blobs->cap,
                                                 blobs->id_base, run_by_block, n_tape, entrances, n_gen, gen_list, err_flag);
    TSIM_LAUNCH_CHECK();
    static int small_tiles = -1;
    if (small_tiles < 0) { const char *e = getenv("TSIM_ENT_TILE"); small_tiles = (e && atoi(e) == 768) ? 0 : 1; }
    const int per_sm = small_tiles ? 9 : 6;
    const int wgrid = div_up(blobs->cap, ENT_WARPS) < 148 * per_sm ? div_up(blobs->cap, ENT_WARPS) : 148 * per_sm;
    if (small_tiles) entrances_warp_kernel<512, 256><<<wgrid, 32 * ENT_WARPS, 0, cs>>>(*cfg, p->cell_type, p->dirs, p->aux, p->block_id, blobs->table, n_gen, gen_list,
                                                            blobs->id_base, run_by_block, n_tape, entrances, n_big, big_list, err_flag);
    else entrances_warp_kernel<768, 384><<<wgrid, 32 * ENT_WARPS, 0, cs>>>(*cfg, p->cell_type, p->dirs, p->aux, p->block_id, blobs->table, n_gen, gen_list,
                                                            blobs->id_base, run_by_block, n_tape, entrances, n_big, big_list, err_flag);
    TSIM_LAUNCH_CHECK();
    const int grid = blobs->cap < 148 * 6 ? blobs->cap : 148 * 6;
    entrances_kernel<<<grid, 128, 0, cs>>>(*cfg, p->cell_type, p->dirs, p->aux, p->block_id, blobs->table, n_big, big_list, blobs->id_base,
                                           run_by_block, n_tape, entrances, err_flag);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}
