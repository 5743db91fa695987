// k_ccl.cu -- connected-component labelling of one cell type in raster discovery order, as a
// union-find over ROW RUNS of a bit-plane.
//
// Replaces the DFS flood fills of _carve_subblock_roads (city_model.py:632-647),
// _flood_fill_blocks_storing_data (:746-763) with its zone fill (:768-806), and the cluster fill of
// _create_intersection_light_groups (:1595-1611).
//
// The reference numbers components by the raster position (y outer, x inner) of their first cell.
// Union-find with "smaller index wins" over runs numbered in raster order makes every component's
// root its first run, whose first cell is the component's minimum raster cell; so
// id = 1 + (number of root runs before it) -- an exclusive scan over root flags.
//
//   1. bits    : ONE streaming read of the type plane (1 B/cell): target cells -> bit-plane, 64 cells
//                per word (a quad of lanes packs a word with two shuffles);
//   2. runs    : run starts = bit & ~previous bit; popcount -> scan numbers the runs in raster order;
//                per run: parent, first cell, length (dense arrays, ~1 run per 50 cells of a city);
//   3. merge   : per word, the overlap segments of row y and row y-1 (one AND) -> one lock-free
//                atomicMin union per segment;
//   4. flatten : parent -> root, root flags -> scan -> component ids; table rows initialised at roots;
//   5. bbox    : per run atomics (min/max x, y, size) into the component table;
//   6. labels  : only when a label plane is wanted (zoning, clusters): one coalesced 4 B/cell write of
//                the ids, fused with the zone fill of the type plane.
//
// PRODUCT SHORT CUT.  The Nothing cells of a freshly laid-out city are the cells of (rows without a road band) x (columns
// without one): the target set is a PRODUCT R x C of a row set and a column set.  Then its components are exactly the
// rectangles (maximal run of consecutive rows of R) x (maximal run of consecutive columns of C), in raster order row run
// by row run -- no union-find at all.  Whether the plane IS such a product is checked exactly on the device (the plane
// is a subset of rows-hit x columns-hit by construction; it equals it iff the cell counts agree), and every kernel of the
// general path returns at once when it is (and the other way round), so the host enqueues both without reading anything
// back.  Carved cities, intersection clusters and anything irregular take the general path.
//
// Carving needs steps 1-5 only (no per-cell label traffic at all); zoning adds step 6, which is the
// compulsory 4 B/cell write of block_id + the 1 B/cell type update.
#include <cstdlib>
#include "scan.cuh"
#include "bitplane.cuh"

namespace tsim {

struct Prod {            // product short cut (see the header comment)
    u64 *col_any;        // [wp] OR of all rows
    uint32_t *row_any;   // [LH] row holds a target cell
    int32_t *rg_lo, *rg_hi, *cg_lo, *cg_hi;   // row / column runs ("gaps" between the road bands), inclusive bounds
    int32_t *row_gap, *col_gap;               // [LH] / [W] run index of the row / column, -1 outside
    int32_t *scal;       // [1] is product, [2] row runs, [3] column runs, [4..5] target cells (64-bit), [6] 0 if product else INT_MAX
};
enum { PS_FLAG = 1, PS_NRG = 2, PS_NCG = 3, PS_POPC = 4, PS_GATE = 6 };

struct Runs {            // lives in the caller's workspace between the label call and the zoning call
    u64 *M;              // [wp * LH] target bit-plane
    int32_t *sprefix;    // [wp * LH] runs started before this word (raster order)
    int32_t *parent;     // [cap] union-find parent; after flatten: the root run
    int32_t *start;      // [cap] first cell of the run (window index)
    int32_t *len;        // [cap]
    int32_t *rank;       // [cap] root flag, then (scan) roots before this run
    int32_t *n_runs;     // device scalar
    int32_t *err;        // the caller's error flag: 24 = run capacity, 25 = component capacity
    int wp, cap;
    Prod pr;
};

// ---------------------------------------------------------------- 1. bit-plane of the target type
__global__ void __launch_bounds__(256) ccl_bits_kernel(int W, int LH, int wp, const uint8_t *__restrict__ T, int target, u64 *__restrict__ M) {
    // blockIdx.y (+ 65535 * blockIdx.z) = row, blockIdx.x * 256 + thread = 16-cell strip of the row: no division per thread
    const int q = threadIdx.x & 3;
    const int s = blockIdx.x * blockDim.x + threadIdx.x, x0 = s * 16;
    const long long y_ll = (long long)blockIdx.y + (long long)blockIdx.z * 65535;
    const bool row_ok = y_ll < LH && s < wp * 4;
    uint32_t m = 0;
    if (row_ok && x0 < W) {
        const size_t base = (size_t)y_ll * W + x0;
        if ((W & 15) == 0) {
            const uint4 tq = __ldg(reinterpret_cast<const uint4 *>(T + base));
            const uint32_t tw[4] = {tq.x, tq.y, tq.z, tq.w};
            m = strip_range_mask(tw, TypeRanges{(uint32_t)target - 1u, (uint32_t)target + 1u, 0u, 0u});   // target >= 1 (Nothing = 6, mask = 1)
        } else {
            for (int k = 0; k < 16 && x0 + k < W; k++) m |= (uint32_t)(T[base + k] == target) << k;
        }
    }
    u64 v = (u64)m << (16 * q);
    v |= __shfl_xor_sync(0xffffffffu, v, 1);
    v |= __shfl_xor_sync(0xffffffffu, v, 2);
    if (row_ok && q == 0) M[(size_t)y_ll * wp + (s >> 2)] = v;
}

// run starts of word (y, wx): set bit whose left neighbour (previous bit, or bit 63 of the previous word of the row) is clear
__device__ __forceinline__ u64 run_starts(const u64 *__restrict__ M, size_t i, int wx) {
    const u64 m = M[i];
    const u64 prev = wx > 0 ? M[i - 1] >> 63 : 0ull;
    return m & ~((m << 1) | prev);
}

// index of the run that contains bit p of word i
__device__ __forceinline__ int run_of(const u64 *__restrict__ M, const int32_t *__restrict__ sprefix, size_t i, int wx, int p) {
    return sprefix[i] + __popcll(run_starts(M, i, wx) & ((2ull << p) - 1ull)) - 1;
}

// ---------------------------------------------------------------- 2. runs
__global__ void __launch_bounds__(256) ccl_count_kernel(long long nw, int wp, const u64 *__restrict__ M, int32_t *__restrict__ cnt, const int32_t *__restrict__ skip) {
    if (__ldg(skip)) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nw) cnt[i] = __popcll(run_starts(M, (size_t)i, (int)((uint32_t)i % (uint32_t)wp)));   // word indices fit 32 bits
}

__global__ void __launch_bounds__(256) ccl_runs_kernel(int W, long long nw, Runs r) {
    if (__ldg(r.pr.scal + PS_FLAG)) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nw) return;
    const int wp = r.wp;
    const long long y = (uint32_t)i / (uint32_t)wp;   // word indices fit 32 bits
    const int wx = (int)((uint32_t)i - (uint32_t)y * (uint32_t)wp);
    u64 st = run_starts(r.M, (size_t)i, wx);
    if (!st) return;
    const u64 m = r.M[i];
    int k = r.sprefix[i];
    while (st) {
        const int p = __ffsll((long long)st) - 1;
        st &= st - 1;
        // length: ones from p upwards, continuing into the following words of the row
        const u64 rest = ~(m >> p);
        int len = rest ? __ffsll((long long)rest) - 1 : 64;
        if (p + len == 64) {
            for (int w2 = wx + 1; w2 < wp; w2++) {
                const u64 inv = ~r.M[i + (w2 - wx)];
                const int l = inv ? __ffsll((long long)inv) - 1 : 64;
                len += l;
                if (l < 64) break;
            }
        }
        if (k < r.cap) {
            r.parent[k] = k;
            r.start[k] = (int32_t)(y * W + wx * 64 + p);
            r.len[k] = len;
        } else {
            *r.err = 24;
        }
        k++;
    }
}

// ---------------------------------------------------------------- 3. merge
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

__global__ void __launch_bounds__(256) ccl_merge_kernel(long long nw, Runs r) {
    if (__ldg(r.pr.scal + PS_FLAG)) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + r.wp;   // rows 1..
    if (i >= nw) return;
    const int wp = r.wp, wx = (int)((uint32_t)i % (uint32_t)wp);
    const u64 ov = r.M[i] & r.M[i - wp];
    if (!ov) return;
    const u64 ovprev = wx > 0 ? (r.M[i - 1] & r.M[i - wp - 1]) >> 63 : 0ull;
    u64 seg = ov & ~((ov << 1) | ovprev);   // first cell of every overlap segment (a segment continuing from the previous word was merged there)
    while (seg) {
        const int p = __ffsll((long long)seg) - 1;
        seg &= seg - 1;
        const int a = run_of(r.M, r.sprefix, (size_t)i, wx, p), b = run_of(r.M, r.sprefix, (size_t)(i - wp), wx, p);
        if (a < r.cap && b < r.cap) uf_union(r.parent, a, b);
    }
}

// ---------------------------------------------------------------- 4. flatten, rank, table rows
__global__ void __launch_bounds__(256) ccl_flatten_kernel(Runs r) {
    const int n = min(*r.n_runs, r.cap);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int root = uf_find(r.parent, k);
        r.rank[k] = (root == k);
        if (root != k) r.parent[k] = root;   // roots keep pointing at themselves, so concurrent finds stay correct
    }
}

__global__ void __launch_bounds__(256) ccl_roots_kernel(int W, int y_global0, Runs r, int32_t *__restrict__ blobs, int cap_blobs, int32_t *err) {
    const int n = min(*r.n_runs, r.cap);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        if (r.parent[k] != k) continue;
        const int id = r.rank[k];   // 0-based
        if (id < cap_blobs) {
            int32_t *b = blobs + (size_t)id * TSIM_BLOB_STRIDE;
            b[0] = 0x7fffffff; b[1] = r.start[k] / W + y_global0; b[2] = -1; b[3] = -1; b[4] = 0; b[5] = r.start[k];
        } else {
            *err = 25;
        }
    }
}

// ---------------------------------------------------------------- 5. bbox / size; parent[] becomes the 0-based component id of the run
__global__ void __launch_bounds__(256) ccl_bbox_kernel(int W, int y_global0, Runs r, int32_t *blobs, int cap_blobs) {
    const int n = min(*r.n_runs, r.cap);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int id = r.rank[r.parent[k]];
        if (id >= cap_blobs) { r.len[k] = id; continue; }
        const int s = r.start[k], len = r.len[k];
        const int x = s % W, y = s / W + y_global0;
        int32_t *b = blobs + (size_t)id * TSIM_BLOB_STRIDE;
        // min / max only move one way, so a plain look first is safe: a stale value can only cause a redundant atomic.
        // (the root run is the component's first row: miny is written there, no atomic at all)
        if (__ldcg(b + 0) > x) atomicMin(b + 0, x);
        if (__ldcg(b + 2) < x + len - 1) atomicMax(b + 2, x + len - 1);
        if (__ldcg(b + 3) < y) atomicMax(b + 3, y);
        atomicAdd(b + 4, len);
        r.len[k] = id;   // from here on len[] holds the 0-based component id of the run (only this thread reads len[k])
    }
}

// ---------------------------------------------------------------- product short cut
// rows / columns hit by the plane and its cell count.  A CTA takes 32 word columns x 64 rows: lanes = words of a row, warps =
// rows; the column ORs are combined through shared memory: 32 atomics per CTA.
__global__ void __launch_bounds__(256) prod_reduce_kernel(int LH, int wp, const u64 *__restrict__ M, Prod p) {
    __shared__ u64 s_col[8][32];
    __shared__ int s_cnt;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ncx = (wp + 31) >> 5;
    const int wx = (blockIdx.x % ncx) * 32 + lane, y0 = (blockIdx.x / ncx) * 64;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    u64 col = 0;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int y = y0 + k * 8 + w;
        const u64 m = (y < LH && wx < wp) ? M[(size_t)y * wp + wx] : 0ull;
        col |= m;
        cnt += __popcll(m);
        if (__any_sync(0xffffffffu, m != 0ull) && lane == 0) p.row_any[y] = 1u;   // idempotent plain store
    }
    s_col[w][lane] = col;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (w == 0) {
        u64 v = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) v |= s_col[q][lane];
        if (v && wx < wp) atomicOr(p.col_any + wx, v);
        if (lane == 0 && s_cnt) atomicAdd((unsigned long long *)(p.scal + PS_POPC), (unsigned long long)s_cnt);
    }
}

// one CTA: runs of consecutive hit rows / hit columns, numbered in order, and the verdict "the plane is rows x columns".
// Both axes are handled as BIT strings in shared memory (a row / column per bit): run starts and run ends are one shift
// and one AND per word, their numbering a popcount scan by one warp per axis, and the run index of every row / column a
// prefix popcount -- no serial walk anywhere (the first version walked and took 125 us per labelling).
constexpr int PROD_MAX_AXIS = 1 << 16;   // rows / columns per axis this kernel handles (longer: the general path)
__global__ void __launch_bounds__(1024) prod_gaps_kernel(int W, int LH, Prod p) {
    __shared__ uint32_t s_bits[2][PROD_MAX_AXIS / 32 + 1];
    __shared__ int s_pref[2][PROD_MAX_AXIS / 32 + 1];   // run starts before the word
    __shared__ int s_runs[2], s_set[2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (W > PROD_MAX_AXIS || LH > PROD_MAX_AXIS) {
        if (threadIdx.x == 0) { p.scal[PS_FLAG] = 0; p.scal[PS_GATE] = 0x7fffffff; }
        return;
    }
    const int nwords[2] = {(LH + 31) >> 5, (W + 31) >> 5};
    // ---- bit strings (zero padded, plus one zero word at the end)
    for (int base = 0; base < nwords[0] * 32; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned bal = __ballot_sync(0xffffffffu, i < LH && p.row_any[i] != 0u);
        if (lane == 0) s_bits[0][i >> 5] = bal;
    }
    for (int i = threadIdx.x; i < nwords[1]; i += 1024) {
        const u64 v = p.col_any[i >> 1];
        uint32_t b = (uint32_t)(i & 1 ? v >> 32 : v);
        if (i == nwords[1] - 1 && (W & 31)) b &= (1u << (W & 31)) - 1u;
        s_bits[1][i] = b;
    }
    if (threadIdx.x < 2) s_bits[threadIdx.x][nwords[threadIdx.x]] = 0u;
    __syncthreads();
    // ---- one warp per axis: number the run starts / ends, write the run bounds
    if (w < 2) {
        const int axis = w, nw = nwords[axis];
        const uint32_t *bits = s_bits[axis];
        int32_t *lo = axis ? p.cg_lo : p.rg_lo, *hi = axis ? p.cg_hi : p.rg_hi;
        const int chunk = (nw + 31) >> 5, w0 = lane * chunk, w1 = min(w0 + chunk, nw);
        auto starts = [&](int i) { const uint32_t b = bits[i], prev = i > 0 ? bits[i - 1] >> 31 : 0u; return b & ~((b << 1) | prev); };
        auto ends = [&](int i) { const uint32_t b = bits[i], next = bits[i + 1] & 1u; return b & ~((b >> 1) | (next << 31)); };
        int ns = 0, ne = 0, nb = 0;
        for (int i = w0; i < w1; i++) { ns += __popc(starts(i)); ne += __popc(ends(i)); nb += __popc(bits[i]); }
        int is = ns, ie = ne;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, is, o), b = __shfl_up_sync(0xffffffffu, ie, o);
            if (lane >= o) { is += a; ie += b; }
        }
        const int total = __shfl_sync(0xffffffffu, is, 31);
        nb = __reduce_add_sync(0xffffffffu, nb);
        int ks = is - ns, ke = ie - ne;
        for (int i = w0; i < w1; i++) {
            s_pref[axis][i] = ks;
            for (uint32_t m = starts(i); m; m &= m - 1) lo[ks++] = i * 32 + __ffs(m) - 1;
            for (uint32_t m = ends(i); m; m &= m - 1) hi[ke++] = i * 32 + __ffs(m) - 1;
        }
        if (lane == 0) { s_runs[axis] = total; s_set[axis] = nb; p.scal[PS_NRG + axis] = total; }
    }
    __syncthreads();
    // ---- run index of every row / column (-1 outside the runs)
    for (int axis = 0; axis < 2; axis++) {
        const int n = axis ? W : LH;
        int32_t *idx = axis ? p.col_gap : p.row_gap;
        const uint32_t *bits = s_bits[axis];
        for (int i = threadIdx.x; i < n; i += 1024) {
            const int wi = i >> 5, b = i & 31;
            const uint32_t v = bits[wi], prev = wi > 0 ? bits[wi - 1] >> 31 : 0u;
            const uint32_t st = v & ~((v << 1) | prev);
            idx[i] = ((v >> b) & 1u) ? s_pref[axis][wi] + __popc(st & ((2u << b) - 1u)) - 1 : -1;
        }
    }
    if (threadIdx.x == 0) {
        const unsigned long long cells = *(const unsigned long long *)(p.scal + PS_POPC);
        const bool product = cells == (unsigned long long)s_set[0] * (unsigned long long)s_set[1];
        p.scal[PS_FLAG] = product ? 1 : 0;
        p.scal[PS_GATE] = product ? 0 : 0x7fffffff;
    }
}

// component table of a product plane: component k = (row run k / n_cg) x (column run k % n_cg)
__global__ void __launch_bounds__(256) prod_table_kernel(int W, int y_global0, Prod p, int32_t *__restrict__ blobs, int cap_blobs, int32_t *count, int32_t *err) {
    if (!p.scal[PS_FLAG]) return;
    const int n_rg = p.scal[PS_NRG], n_cg = p.scal[PS_NCG];
    const long long n = (long long)n_rg * n_cg;
    if (blockIdx.x == 0 && threadIdx.x == 0) { *count = (int)(n < 0x7fffffffLL ? n : 0x7fffffffLL); if (n > cap_blobs) *err = 25; }
    const long long lim = n < cap_blobs ? n : cap_blobs;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < lim; k += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(k / n_cg), j = (int)(k % n_cg);
        const int x0 = p.cg_lo[j], x1 = p.cg_hi[j], y0 = p.rg_lo[i], y1 = p.rg_hi[i];
        int32_t *b = blobs + (size_t)k * TSIM_BLOB_STRIDE;
        b[0] = x0; b[1] = y0 + y_global0; b[2] = x1; b[3] = y1 + y_global0; b[4] = (x1 - x0 + 1) * (y1 - y0 + 1); b[5] = y0 * W + x0;
    }
}

// ---------------------------------------------------------------- 6. label plane (+ zone fill)
// fill[id] (u8): new type of the component's cells, or 0xff = leave the type plane alone
__global__ void __launch_bounds__(256) zones_table_kernel(const int32_t *__restrict__ blobs, const int32_t *__restrict__ n_blobs, int cap_blobs,
                                                          const int32_t *__restrict__ id_base, const uint8_t *__restrict__ zone, int n_tape,
                                                          uint8_t *__restrict__ fill, int32_t *err) {
    const int n = min(*n_blobs, cap_blobs);
    const int base = id_base ? *id_base : 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int32_t *b = blobs + (size_t)k * TSIM_BLOB_STRIDE;
        const bool thin = (b[2] - b[0] + 1 < 3) || (b[3] - b[1] + 1 < 3);   // :768-770
        int t = T_EMPTY;
        if (!thin) {
            const int gid = k + base;   // 0-based global id
            if (gid < 0) { t = T_EMPTY; }                       // a component cut by the window's lower edge: its rows are not owned here
            else if (gid >= n_tape) { *err = 1; }
            else { t = zone[gid]; if (t > T_OTH) { *err = 2; t = T_EMPTY; } }
        }
        fill[k] = (uint8_t)t;
    }
}

// one group of 8 cells of a row (inside one word): label plane + zone fill
template <bool FILL>
__device__ __forceinline__ void label_group8(const Runs &r, int x, long long y, long long i0, int base, bool prod, int n_cg, int cap_blobs,
                                             int32_t *__restrict__ L, uint8_t *__restrict__ T, const uint8_t *__restrict__ fill) {
    const int wp = r.wp;
    const int wx = x >> 6, sh = x & 63;
    const size_t wi = (size_t)y * wp + wx;
    const uint32_t bits = (uint32_t)(r.M[wi] >> sh) & 0xffu;
    int4 lo = make_int4(0, 0, 0, 0), hi = lo;
    if (bits && prod) {   // product plane: id = row run * column runs + column run (rows / columns of a set cell always have one)
        const int row0 = r.pr.row_gap[y] * n_cg;
        const int4 c0 = *reinterpret_cast<const int4 *>(r.pr.col_gap + x), c1 = *reinterpret_cast<const int4 *>(r.pr.col_gap + x + 4);
        const int cg[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        int ids[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (bits == 0xffu && cg[0] == cg[7]) {
            const int cid = row0 + cg[0], id = cid + base;
            lo = hi = make_int4(id, id, id, id);
            if (FILL && cid < cap_blobs) {
                const uint32_t f = fill[cid];
                if (f != 0xffu) { const uint32_t f4 = f * 0x01010101u; *reinterpret_cast<uint2 *>(T + i0) = make_uint2(f4, f4); }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (!((bits >> k) & 1u)) continue;
                const int cid = row0 + cg[k];
                ids[k] = cid + base;
                if (FILL && cid < cap_blobs) { const uint8_t f = fill[cid]; if (f != 0xff) T[i0 + k] = f; }
            }
            lo = make_int4(ids[0], ids[1], ids[2], ids[3]); hi = make_int4(ids[4], ids[5], ids[6], ids[7]);
        }
    } else if (bits) {
        const u64 st = run_starts(r.M, wi, wx);
        const int sp = r.sprefix[wi];
        const uint32_t inner = (uint32_t)(st >> sh) & 0xfeu;   // run starts strictly inside the group
        if (bits == 0xffu && !inner) {   // the whole group lies in one run: one look-up, splat
            const int run = sp + __popcll(st & ((2ull << sh) - 1ull)) - 1;
            const int cid = run < r.cap ? r.len[run] : 0, id = cid + base;
            lo = hi = make_int4(id, id, id, id);
            if (FILL && cid < cap_blobs) {
                const uint32_t f = fill[cid];
                if (f != 0xffu) { const uint32_t f4 = f * 0x01010101u; *reinterpret_cast<uint2 *>(T + i0) = make_uint2(f4, f4); }
            }
        } else {
            int ids[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            int last_run = -1, last_id = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (!((bits >> k) & 1u)) continue;
                const int run = sp + __popcll(st & ((2ull << (sh + k)) - 1ull)) - 1;
                if (run != last_run) { last_run = run; last_id = run < r.cap ? r.len[run] : 0; }
                ids[k] = last_id + base;
                if (FILL && last_id < cap_blobs) { const uint8_t f = fill[last_id]; if (f != 0xff) T[i0 + k] = f; }
            }
            lo = make_int4(ids[0], ids[1], ids[2], ids[3]); hi = make_int4(ids[4], ids[5], ids[6], ids[7]);
        }
    }
    *reinterpret_cast<int4 *>(L + i0) = lo;
    *reinterpret_cast<int4 *>(L + i0 + 4) = hi;
}

template <bool FILL>
__global__ void __launch_bounds__(256) ccl_labels_kernel(int W, long long n, Runs r, const int32_t *__restrict__ id_base, int cap_blobs,
                                                         int32_t *__restrict__ L, uint8_t *__restrict__ T, const uint8_t *__restrict__ fill) {
    // blockIdx.y (+ 65535 * blockIdx.z) = row, blockIdx.x * 256 + thread = group of 8 cells of the row: no 64-bit division per thread
    const long long y = (long long)blockIdx.y + (long long)blockIdx.z * 65535;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (x >= W || y * W >= n) return;
    const long long i0 = y * W + x;
    const int base = (id_base ? *id_base : 0) + 1;
    const int wp = r.wp;
    const bool prod = r.pr.scal[PS_FLAG] != 0;
    const int n_cg = r.pr.scal[PS_NCG];
    if ((W & 7) == 0) {
        label_group8<FILL>(r, x, y, i0, base, prod, n_cg, cap_blobs, L, T, fill);
    } else {
        for (int k = 0; k < 8 && x + k < W; k++) {
            const long long i = i0 + k;
            const int xk = x + k, wx = xk >> 6, p = xk & 63;
            const size_t wi = (size_t)y * wp + wx;
            int id = 0;
            if ((r.M[wi] >> p) & 1ull) {
                const int run = prod ? 0 : run_of(r.M, r.sprefix, wi, wx, p);
                const int cid = prod ? r.pr.row_gap[y] * n_cg + r.pr.col_gap[xk] : (run < r.cap ? r.len[run] : 0);
                id = cid + base;
                if (FILL && cid < cap_blobs) { const uint8_t f = fill[cid]; if (f != 0xff) T[i] = f; }
            }
            L[i] = id;
        }
    }
}

// 4 bits -> 4 bytes of 0xff / 0x00
__device__ __forceinline__ uint32_t spread4(uint32_t b) {
    const uint32_t x = (b | (b << 7) | (b << 14) | (b << 21)) & 0x01010101u;
    return x * 0xffu;
}

// The same plane, G rows per thread (W % 8 == 0).  The label write is a chain of dependent look-ups per group (word -> run -> component ->
// fill type) ending in 32 bytes of stores: with one group per thread the stores in flight per SM bound the kernel at ~2.9 TB/s of the
// ~6 TB/s a write stream reaches.  Here the G chains of a thread advance together, level by level, and every group's type bytes leave as
// ONE blended 8-byte store instead of up to eight byte stores.
template <bool FILL, int G>
__global__ void __launch_bounds__(256) ccl_labels_rows_kernel(int W, int LH, Runs r, const int32_t *__restrict__ id_base, int cap_blobs,
                                                              int32_t *__restrict__ L, uint8_t *__restrict__ T, const uint8_t *__restrict__ fill) {
    const int yb = (int)(blockIdx.y + blockIdx.z * 65535u) * G;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (x >= W || yb >= LH) return;
    const int base = (id_base ? *id_base : 0) + 1;
    const bool prod = r.pr.scal[PS_FLAG] != 0;
    if (prod) {
        const int n_cg = r.pr.scal[PS_NCG];
        for (int j = 0; j < G && yb + j < LH; j++) label_group8<FILL>(r, x, yb + j, (long long)(yb + j) * W + x, base, true, n_cg, cap_blobs, L, T, fill);
        return;
    }
    const int wp = r.wp, wx = x >> 6, sh = x & 63, cap = r.cap;
    const u64 *__restrict__ M = r.M;
    const int32_t *__restrict__ sprefix = r.sprefix;
    const int32_t *__restrict__ comp = r.len;   // component of every run by now
    // level 1: the words
    u64 m[G], pv[G];
    int sp[G];
    uint2 t8[G];
    bool ok[G];
#pragma unroll
    for (int j = 0; j < G; j++) {
        ok[j] = yb + j < LH;
        const size_t wi = (size_t)(yb + j) * wp + wx;
        m[j] = ok[j] ? M[wi] : 0ull;
        pv[j] = ok[j] && wx > 0 ? M[wi - 1] >> 63 : 0ull;
        sp[j] = ok[j] ? sprefix[wi] : 0;
        if (FILL) t8[j] = ok[j] ? *reinterpret_cast<const uint2 *>(T + (size_t)(yb + j) * W + x) : make_uint2(0u, 0u);
    }
    // level 2: the component of the first run of every group
    uint32_t bits[G], seg[G];
    u64 st[G];
    int id0[G];
#pragma unroll
    for (int j = 0; j < G; j++) {
        bits[j] = (uint32_t)(m[j] >> sh) & 0xffu;
        st[j] = m[j] & ~((m[j] << 1) | pv[j]);
        seg[j] = 0; id0[j] = 0;
        if (bits[j]) {
            const int first = __ffs(bits[j]) - 1;
            const uint32_t gap = ~(bits[j] >> first);             // the run goes on until the first clear bit (bits has 8 bits: there is one)
            seg[j] = ((1u << (__ffs(gap) - 1)) - 1u) << first;
            const int run = sp[j] + __popcll(st[j] & ((2ull << (sh + first)) - 1ull)) - 1;
            id0[j] = run < cap ? comp[run] : 0;
        }
    }
    // level 3: its fill type
    uint32_t f0[G];
#pragma unroll
    for (int j = 0; j < G; j++) f0[j] = (FILL && seg[j] && id0[j] < cap_blobs) ? (uint32_t)fill[id0[j]] : 0xffu;
    // level 4: ids, blended types, the rare further runs of a group, stores
#pragma unroll
    for (int j = 0; j < G; j++) {
        if (!ok[j]) continue;
        int ids[8];
        uint32_t tlo = t8[j].x, thi = t8[j].y;
        bool tch = false;
        auto put = [&](uint32_t sg, int id, uint32_t f) {
#pragma unroll
            for (int k = 0; k < 8; k++) if ((sg >> k) & 1u) ids[k] = id + base;
            if (FILL && f != 0xffu) {
                const uint32_t f4 = f * 0x01010101u, ml = spread4(sg & 15u), mh = spread4(sg >> 4);
                tlo = (tlo & ~ml) | (f4 & ml); thi = (thi & ~mh) | (f4 & mh);
                tch = true;
            }
        };
#pragma unroll
        for (int k = 0; k < 8; k++) ids[k] = 0;
        if (seg[j]) put(seg[j], id0[j], f0[j]);
        uint32_t rem = bits[j] & ~seg[j];
        while (rem) {
            const int first = __ffs(rem) - 1;
            const uint32_t gap = ~(rem >> first);
            const uint32_t sg = ((1u << (__ffs(gap) - 1)) - 1u) << first;
            const int run = sp[j] + __popcll(st[j] & ((2ull << (sh + first)) - 1ull)) - 1;
            const int id = run < cap ? comp[run] : 0;
            put(sg, id, (FILL && id < cap_blobs) ? (uint32_t)fill[id] : 0xffu);
            rem &= ~sg;
        }
        const size_t i0 = (size_t)(yb + j) * W + x;
        *reinterpret_cast<int4 *>(L + i0) = make_int4(ids[0], ids[1], ids[2], ids[3]);
        *reinterpret_cast<int4 *>(L + i0 + 4) = make_int4(ids[4], ids[5], ids[6], ids[7]);
        if (FILL && tch) *reinterpret_cast<uint2 *>(T + i0) = make_uint2(tlo, thi);
    }
}

// shard bookkeeping over the component table (one launch instead of a dozen tensor ops): out[0] = roots below the own
// rows, out[1] = roots inside the own rows, out[2] = 1 if a component meets the own rows AND a cut edge of the window
__global__ void __launch_bounds__(256) shard_counts_kernel(tsim_cfg c, const int32_t *__restrict__ blobs, const int32_t *__restrict__ n_blobs, int cap,
                                                           int own_lo, int own_hi, int32_t *out) {
    const int n = min(*n_blobs, cap);
    const long long lo_cell = (long long)(own_lo - c.win_y0) * c.width, hi_cell = (long long)(own_hi - c.win_y0) * c.width;
    const int win_lo = c.win_y0, win_hi = c.win_y0 + c.win_rows;
    int below = 0, own = 0, bad = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int32_t *b = blobs + (size_t)k * TSIM_BLOB_STRIDE;
        const long long root = b[5];
        below += root < lo_cell;
        own += root >= lo_cell && root < hi_cell;
        const bool meets = b[3] >= own_lo && b[1] < own_hi;
        const bool cut = (b[1] <= win_lo + 1 && win_lo > 0) || (b[3] >= win_hi - 2 && win_hi < c.height);   // ring cells need one more row
        bad |= meets && cut;
    }
    below = __reduce_add_sync(0xffffffffu, below);
    own = __reduce_add_sync(0xffffffffu, own);
    bad = __reduce_or_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        if (below) atomicAdd(out + 0, below);
        if (own) atomicAdd(out + 1, own);
        if (bad) atomicOr(out + 2, 1);
    }
}

// workspace layout shared by the label call and the calls that materialise its result
static tsim_status runs_layout(const tsim_cfg *cfg, void *workspace, size_t ws_bytes, Runs &r, int32_t *&scan_tmp, uint8_t *&fill, int cap_blobs) {
    const Win win(*cfg);
    const long long n = win.cells();
    const int wp = div_up(win.W, 64);
    const long long nw = (long long)wp * win.LH;
    const int cap = (int)(n / 4 + 1024);
    char *w = (char *)workspace;
    size_t o = 0;
    auto take = [&](size_t bytes) { char *q = w + o; o += (bytes + 255) & ~(size_t)255; return q; };
    int32_t *scal = (int32_t *)take(64 * 4);
    r.n_runs = scal; r.err = nullptr;
    r.M = (u64 *)take((size_t)nw * 8);
    r.sprefix = (int32_t *)take((size_t)nw * 4);
    r.parent = (int32_t *)take((size_t)cap * 4);
    r.start = (int32_t *)take((size_t)cap * 4);
    r.len = (int32_t *)take((size_t)cap * 4);
    r.rank = (int32_t *)take((size_t)cap * 4);
    scan_tmp = (int32_t *)take(scan_tmp_bytes(nw > cap ? nw : cap));
    fill = (uint8_t *)take((size_t)(cap_blobs > 0 ? cap_blobs : 1));
    r.wp = wp; r.cap = cap;
    r.pr.scal = scal;
    r.pr.col_any = (u64 *)take((size_t)wp * 8 + (size_t)win.LH * 4);   // col_any and row_any in one piece (one memset)
    r.pr.row_any = (uint32_t *)(r.pr.col_any + wp);
    r.pr.rg_lo = (int32_t *)take((size_t)(win.LH / 2 + 2) * 4); r.pr.rg_hi = (int32_t *)take((size_t)(win.LH / 2 + 2) * 4);
    r.pr.cg_lo = (int32_t *)take((size_t)(win.W / 2 + 2) * 4); r.pr.cg_hi = (int32_t *)take((size_t)(win.W / 2 + 2) * 4);
    r.pr.row_gap = (int32_t *)take((size_t)win.LH * 4); r.pr.col_gap = (int32_t *)take((size_t)(win.W + 8) * 4);
    if (!workspace || o > ws_bytes) { set_error("labelling needs %zu workspace bytes, got %zu", o, ws_bytes); return TSIM_ERR_WORKSPACE; }
    return TSIM_OK;
}

__global__ void set_i32_kernel(int32_t *p, int32_t v) { *p = v; }

// TSIM_CCL_PRODUCT=0 switches the product short cut off (tests: both paths must label alike)
static bool product_shortcut_enabled() {
    const char *e = getenv("TSIM_CCL_PRODUCT");
    return !(e && *e == '0');
}

static int list_grid(long long n) {
    long long b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    return (int)(b < 1 ? 1 : b);
}

// steps 1-5
tsim_status label_type(const tsim_cfg *cfg, const uint8_t *T, int target, const tsim_blobs *blobs, int32_t *err_flag, void *workspace,
                       size_t ws_bytes, cudaStream_t cs) {
    Runs r; int32_t *scan_tmp; uint8_t *fill;
    tsim_status st = runs_layout(cfg, workspace, ws_bytes, r, scan_tmp, fill, blobs->cap);
    if (st != TSIM_OK) return st;
    r.err = err_flag;
    const Win win(*cfg);
    const long long nw = (long long)r.wp * win.LH;
    TSIM_CUDA(cudaMemsetAsync(r.n_runs, 0, 64 * 4, cs));
    ccl_bits_kernel<<<dim3(div_up(r.wp * 4, 256), win.LH < 65535 ? win.LH : 65535, div_up(win.LH, 65535)), 256, 0, cs>>>(win.W, win.LH, r.wp, T, target, r.M);
    TSIM_LAUNCH_CHECK();
    // product short cut: is the plane (rows hit) x (columns hit)?  Then the general path below is skipped on the device.
    const bool try_product = product_shortcut_enabled();
    if (try_product) {
        TSIM_CUDA(cudaMemsetAsync(r.pr.col_any, 0, (size_t)r.wp * 8 + (size_t)win.LH * 4, cs));
        prod_reduce_kernel<<<div_up(r.wp, 32) * div_up(win.LH, 64), 256, 0, cs>>>(win.LH, r.wp, r.M, r.pr);
        TSIM_LAUNCH_CHECK();
        prod_gaps_kernel<<<1, 1024, 0, cs>>>(win.W, win.LH, r.pr);
        TSIM_LAUNCH_CHECK();
    } else {
        set_i32_kernel<<<1, 1, 0, cs>>>(r.pr.scal + PS_GATE, 0x7fffffff);   // (PS_FLAG is 0 from the memset)
        TSIM_LAUNCH_CHECK();
    }
    ccl_count_kernel<<<div_up(nw, 256), 256, 0, cs>>>(nw, r.wp, r.M, r.sprefix, r.pr.scal + PS_FLAG);
    TSIM_LAUNCH_CHECK();
    if ((st = exclusive_scan_i32(r.sprefix, nw, scan_tmp, r.n_runs, cs, r.pr.scal + PS_GATE)) != TSIM_OK) return st;
    ccl_runs_kernel<<<div_up(nw, 256), 256, 0, cs>>>(win.W, nw, r);
    TSIM_LAUNCH_CHECK();
    if (win.LH > 1) {
        ccl_merge_kernel<<<div_up(nw - r.wp, 256), 256, 0, cs>>>(nw, r);
        TSIM_LAUNCH_CHECK();
    }
    const int g = list_grid(r.cap);
    ccl_flatten_kernel<<<g, 256, 0, cs>>>(r);
    TSIM_LAUNCH_CHECK();
    if ((st = exclusive_scan_i32(r.rank, r.cap, scan_tmp, blobs->count, cs, r.n_runs)) != TSIM_OK) return st;
    ccl_roots_kernel<<<g, 256, 0, cs>>>(win.W, win.y0, r, blobs->table, blobs->cap, r.err);
    TSIM_LAUNCH_CHECK();
    ccl_bbox_kernel<<<g, 256, 0, cs>>>(win.W, win.y0, r, blobs->table, blobs->cap);
    TSIM_LAUNCH_CHECK();
    if (try_product) {   // last: the general path's scan has just written a count of 0
        prod_table_kernel<<<list_grid(blobs->cap), 256, 0, cs>>>(win.W, win.y0, r.pr, blobs->table, blobs->cap, blobs->count, r.err);
        TSIM_LAUNCH_CHECK();
    }
    return TSIM_OK;
}

// label plane launcher: G rows per thread when the rows are made of whole 8-cell groups (TSIM_LABEL_ROWS=1: one row per thread, the r1 kernel)
template <bool FILL>
static void launch_labels(int W, int LH, long long n, const Runs &r, const int32_t *id_base, int cap_blobs, int32_t *L, uint8_t *T, const uint8_t *fill,
                          cudaStream_t cs) {
    static int rows = -1;
    if (rows < 0) { const char *e = getenv("TSIM_LABEL_ROWS"); rows = e ? atoi(e) : 4; }
    if ((W & 7) == 0 && rows == 4) {
        const int yb = div_up(LH, 4);
        ccl_labels_rows_kernel<FILL, 4><<<dim3(div_up(div_up(W, 8), 256), yb < 65535 ? yb : 65535, div_up(yb, 65535)), 256, 0, cs>>>(W, LH, r, id_base, cap_blobs,
                                                                                                                                  L, T, fill);
    } else if ((W & 7) == 0 && rows == 2) {
        const int yb = div_up(LH, 2);
        ccl_labels_rows_kernel<FILL, 2><<<dim3(div_up(div_up(W, 8), 256), yb < 65535 ? yb : 65535, div_up(yb, 65535)), 256, 0, cs>>>(W, LH, r, id_base, cap_blobs,
                                                                                                                                  L, T, fill);
    } else {
        ccl_labels_kernel<FILL><<<dim3(div_up(div_up(W, 8), 256), LH < 65535 ? LH : 65535, div_up(LH, 65535)), 256, 0, cs>>>(W, n, r, id_base, cap_blobs, L, T, fill);
    }
}

}  // namespace tsim

using namespace tsim;

extern "C" tsim_status tsim_layout_label_nothing(const tsim_cfg *cfg, const tsim_planes *p, const tsim_blobs *blobs, int32_t *err_flag,
                                                 void *workspace, size_t ws_bytes, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if ((st = check_blobs(blobs, "tsim_layout_label_nothing")) != TSIM_OK) return st;
    if (!p || !p->cell_type || !err_flag) { set_error("tsim_layout_label_nothing: NULL plane or err_flag"); return TSIM_ERR_CONFIG; }
    return label_type(cfg, p->cell_type, T_NOTHING, blobs, err_flag, workspace, ws_bytes, (cudaStream_t)stream);
}

extern "C" tsim_status tsim_layout_zones(const tsim_cfg *cfg, const tsim_planes *p, const tsim_blobs *blobs, const uint8_t *zone_by_block,
                                         int32_t n_tape, int32_t *err_flag, void *workspace, size_t ws_bytes, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if ((st = check_blobs(blobs, "tsim_layout_zones")) != TSIM_OK) return st;
    if (!p || !p->cell_type || !p->block_id || !zone_by_block || !err_flag || n_tape < 0) { set_error("tsim_layout_zones: bad arguments"); return TSIM_ERR_CONFIG; }
    if ((((uintptr_t)p->block_id) & 15) != 0 || (((uintptr_t)p->cell_type) & 7) != 0) { set_error("tsim_layout_zones: block_id must be 16-byte, cell_type 8-byte aligned"); return TSIM_ERR_CONFIG; }
    Runs r; int32_t *scan_tmp; uint8_t *fill;
    if ((st = runs_layout(cfg, workspace, ws_bytes, r, scan_tmp, fill, blobs->cap)) != TSIM_OK) return st;
    cudaStream_t cs = (cudaStream_t)stream;
    const Win win(*cfg);
    const long long n = win.cells();
    zones_table_kernel<<<list_grid(blobs->cap), 256, 0, cs>>>(blobs->table, blobs->count, blobs->cap, blobs->id_base, zone_by_block, n_tape, fill, err_flag);
    TSIM_LAUNCH_CHECK();
    // Nothing cells carry no arrows and no aux bits (frame pass / place_cell), so only the type changes
    launch_labels<true>(win.W, win.LH, n, r, blobs->id_base, blobs->cap, p->block_id, p->cell_type, fill, cs);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}

extern "C" tsim_status tsim_label_mask(const tsim_cfg *cfg, const uint8_t *mask, int32_t *labels, const tsim_blobs *blobs, int32_t *err_flag,
                                       void *workspace, size_t ws_bytes, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if ((st = check_blobs(blobs, "tsim_label_mask")) != TSIM_OK) return st;
    if (!mask || !labels || !err_flag || ((uintptr_t)labels & 15)) { set_error("tsim_label_mask: bad arguments"); return TSIM_ERR_CONFIG; }
    cudaStream_t cs = (cudaStream_t)stream;
    if ((st = label_type(cfg, mask, 1, blobs, err_flag, workspace, ws_bytes, cs)) != TSIM_OK) return st;
    Runs r; int32_t *scan_tmp; uint8_t *fill;
    if ((st = runs_layout(cfg, workspace, ws_bytes, r, scan_tmp, fill, blobs->cap)) != TSIM_OK) return st;
    const Win win(*cfg);
    const long long n = win.cells();
    launch_labels<false>(win.W, win.LH, n, r, blobs->id_base, blobs->cap, labels, nullptr, nullptr, cs);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}

extern "C" tsim_status tsim_shard_counts(const tsim_cfg *cfg, const tsim_blobs *blobs, int32_t own_lo, int32_t own_hi, int32_t *out, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if ((st = check_blobs(blobs, "tsim_shard_counts")) != TSIM_OK) return st;
    if (!out || own_lo < cfg->win_y0 || own_hi > cfg->win_y0 + cfg->win_rows || own_lo >= own_hi) { set_error("tsim_shard_counts: bad arguments"); return TSIM_ERR_CONFIG; }
    cudaStream_t cs = (cudaStream_t)stream;
    TSIM_CUDA(cudaMemsetAsync(out, 0, 3 * sizeof(int32_t), cs));
    shard_counts_kernel<<<list_grid(blobs->cap), 256, 0, cs>>>(*cfg, blobs->table, blobs->count, blobs->cap, own_lo, own_hi, out);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}
