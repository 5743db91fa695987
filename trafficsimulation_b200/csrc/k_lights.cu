// k_lights.cu -- _add_traffic_lights (city_model.py:1422-1499), _assign_traffic_light (:1501-1520),
// _scan_for_traffic_flow_reverse (:1528-1548) and CellAgent.leads_to (cell.py:201-227).
//
// The reference sweeps the grid COLUMN-major and mutates it while it goes; three things depend on
// that order and are reproduced with explicit predicates instead of a serial sweep:
//   * a neighbour already converted to ControlledRoad (it precedes the cell in column-major order and
//     itself points into an Intersection) no longer has its original type -- `type_at_time`;
//   * Sidewalk -> TrafficLight conversions never influence a test (both are accepted), so they commute;
//   * `nb.light` of a cell that is converted later is reset by the conversion (a fresh CellAgent).
//
// Everything except the first kernel works on BIT-PLANES (64 cells per 64-bit word, 1/8 B per cell):
//   1. bits    : one streaming read of T and D (3 B/cell) -> arrow planes N/E/S/W, Intersection plane,
//                road-without-intersection plane; picks the pivot intersection;
//   2. cr      : ControlledRoad candidates = R & (arrow planes & shifted Intersection plane): pure word
//                logic; popcount -> scan -> compact list of the (~2.5 % of all) candidate cells;
//   3. reach   : `leads_to` is a BFS over the whole directed arrow graph.  It is answered in O(1) from
//                two reachability planes, FW = reachable from the pivot, BW = can reach the pivot
//                (BW(a) && FW(b) implies a -> pivot -> b).  The planes are the fixed point of alternating
//                LINE closures: a row sweep closes every row under its E/W arrows (64 cells per
//                Kogge-Stone fill, carries across the words of a row resolved like an adder's carry chain
//                from two warp ballots), a column sweep closes every column under its N/S arrows (64
//                columns per word in parallel, rows chunked over the threads of a CTA with a
//                generate/propagate scan).  A path with k turns is covered after k sweeps, and a grid
//                city needs a handful; the loop runs until a whole alternation changes nothing, so the
//                result is the exact closure whatever the number of turns.  Queries whose source cannot
//                reach the pivot or whose target the pivot cannot reach (sink / source regions such as
//                highway exit lanes) are searched exactly with a bounded visited list that raises
//                TSIM_ERR_CAPACITY rather than guessing;
//   4. eval    : ONE thread per candidate evaluates the reference's per-road logic once and keeps the
//                result as a 64-bit record (light cells as 5-bit offsets, reverse-scan lengths per
//                direction); lights are marked in a bit-plane with atomicOr;
//   5. lights in ascending cell order = popcount scan of that plane; CSR link tables by
//      count -> scan -> fill from the records; type / aux updates are written for the candidate and
//      light cells only (sparse), never as a full-plane sweep.
#include <cooperative_groups.h>
#include <cstdlib>
#include <cstring>
#include "scan.cuh"
#include "bitplane.cuh"

namespace cg = cooperative_groups;

namespace tsim {


constexpr int BFS_CAP = 160;        // visited cells of an exact fallback search
constexpr int MAX_TL_RANGE = 30;    // reverse-scan lengths are kept in 5 bits

struct Bits {
    u64 *aN, *aE, *aS, *aW;   // arrow planes
    u64 *I, *R;               // Intersection / road-like-without-intersection
    u64 *fw, *bw;             // reachable from the pivot / reaches the pivot
    u64 *cr, *tl;             // ControlledRoad candidates / TrafficLight cells
    u64 *lm;                  // cells of the reverse marches of the accepted candidates ("nb.light = tl", :1542), merged into aux after the evaluation
    int wp;                   // words per row
};

struct LightsCtx {
    int W, H, tl_range, cap_lights, fwd, fwd_mode;
    int stages;           // bit 0: Z witnesses off, bit 1: window closures off (TSIM_LIGHTS_STAGES, tests only)
    int cut_lo, cut_hi;   // rows at the window's low / high end that belong to a neighbour shard (0: that end is a grid edge)
    const uint8_t *T; const uint16_t *D;
    Bits b;
    int32_t *err;
    __device__ __forceinline__ bool has(int x, int y) const { return x >= 0 && x < W && y >= 0 && y < H; }
    __device__ __forceinline__ int at(int x, int y) const { return y * W + x; }
    __device__ __forceinline__ bool bit(const u64 *pl, int x, int y) const { return (pl[(size_t)y * b.wp + (x >> 6)] >> (x & 63)) & 1ull; }
};

// ---------------------------------------------------------------- 1. bit-planes from T and D
template <bool MUL>
__global__ void __launch_bounds__(256) lights_bits_kernel(int W, int H, const uint8_t *__restrict__ T, const uint16_t *__restrict__ D, Bits bp,
                                                          long long mid, int32_t *piv /* [0] first >= mid, [1] first */) {
    // blockIdx.y = row, blockIdx.x * 256 + thread = 16-cell strip of the row (a quad of lanes = one word).  (As one flat index
    // this took a 64-bit division per thread: 26 % of the kernel's stall samples.)
    const int lane = threadIdx.x & 31, q = lane & 3;
    const int s = blockIdx.x * blockDim.x + threadIdx.x, x0 = s * 16;
    const int y = blockIdx.y + blockIdx.z * 65535;
    const bool row_ok = y < H && s < bp.wp * 4;
    uint32_t mN = 0, mE = 0, mS = 0, mW = 0, mI = 0, mR = 0;
    int p_all = 0x7fffffff, p_mid = 0x7fffffff;
    if (row_ok && x0 < W) {
        const size_t base = (size_t)y * W + x0;
        uint32_t tw[4], dw[8];
        if ((W & 15) == 0) {
            const uint4 tq = __ldg(reinterpret_cast<const uint4 *>(T + base));
            const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(D + base)), d1 = __ldg(reinterpret_cast<const uint4 *>(D + base + 8));
            tw[0] = tq.x; tw[1] = tq.y; tw[2] = tq.z; tw[3] = tq.w;
            dw[0] = d0.x; dw[1] = d0.y; dw[2] = d0.z; dw[3] = d0.w; dw[4] = d1.x; dw[5] = d1.y; dw[6] = d1.z; dw[7] = d1.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) tw[k] = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) dw[k] = 0;
            for (int k = 0; k < 16 && x0 + k < W; k++) {
                tw[k >> 2] |= (uint32_t)T[base + k] << (8 * (k & 3));
                dw[k >> 1] |= (uint32_t)D[base + k] << (16 * (k & 1));
            }
            for (int k = W - x0; k < 16; k++) if (k >= 0) tw[k >> 2] |= (uint32_t)T_WALL << (8 * (k & 3));
        }
        // SWAR: type sets as byte-range tests on 4 cells at once, arrow bits from 2 cells per dirs word
        mI = strip_range_mask(tw, TypeRanges{T_INTER - 1, T_INTER + 1, 0, 0});
        mR = strip_range_mask(tw, TypeRanges{T_R1 - 1, T_HWY_OUT + 1, T_BE - 1, T_BE + 1}) & ~mI;   // ROAD_LIKE_TYPES_WITHOUT_INTERSECTIONS (config.py:69)
        uint32_t nz = 0;   // cells with at least one arrow
        if (MUL) {
            // the arrow mask is the low nibble of every cell's dirs word: low bytes of 4 cells -> one word (PRMT), then one multiplication
            // gathers bit d of the four bytes (632 -> 504 SASS instructions; TSIM_LIGHTS_BITS=shift selects the older form)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t lb = __byte_perm(dw[2 * j], dw[2 * j + 1], 0x6420);
                auto four = [&](int d) { return ((((lb >> d) & 0x01010101u) * 0x01020408u) >> 24) << (4 * j); };
                mN |= four(DN); mE |= four(DE); mS |= four(DS); mW |= four(DW);
            }
            nz = mN | mE | mS | mW;   // a dirs word is 0 or carries a non-empty mask (dl_append)
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t w = dw[k];
                auto two = [&](int d) { const uint32_t v = (w >> d) & 0x00010001u; return ((v | (v >> 15)) & 3u) << (2 * k); };
                mN |= two(DN); mE |= two(DE); mS |= two(DS); mW |= two(DW);
                nz |= ((uint32_t)((w & 0xffffu) != 0u) | ((uint32_t)((w >> 16) != 0u) << 1)) << (2 * k);
            }
        }
        uint32_t cand = mI & nz;   // Intersection cells that still have an arrow: pivot candidates
        if (cand) {
            p_all = (int)base + __ffs(cand) - 1;
            const long long rel = mid - (long long)base;
            if (rel > 0) cand = rel < 16 ? cand & ~((1u << rel) - 1u) : 0u;
            if (cand) p_mid = (int)base + __ffs(cand) - 1;
        }
    }
    // quad of lanes -> one 64-bit word per plane
    auto pack = [&](uint32_t m) {
        u64 v = (u64)m << (16 * q);
        v |= __shfl_xor_sync(0xffffffffu, v, 1);
        v |= __shfl_xor_sync(0xffffffffu, v, 2);
        return v;
    };
    const u64 wN = pack(mN), wE = pack(mE), wS = pack(mS), wW = pack(mW), wI = pack(mI), wR = pack(mR);
    if (row_ok && q == 0) {
        const size_t o = (size_t)y * bp.wp + (s >> 2);
        bp.aN[o] = wN; bp.aE[o] = wE; bp.aS[o] = wS; bp.aW[o] = wW; bp.I[o] = wI; bp.R[o] = wR;
        bp.fw[o] = 0ull; bp.bw[o] = 0ull; bp.tl[o] = 0ull; bp.lm[o] = 0ull;
    }
    p_all = __reduce_min_sync(0xffffffffu, p_all);
    p_mid = __reduce_min_sync(0xffffffffu, p_mid);
    if (lane == 0) {
        if (p_all != 0x7fffffff) atomicMin(piv + 1, p_all);
        if (p_mid != 0x7fffffff) atomicMin(piv + 0, p_mid);
    }
}

__global__ void init_pivot_kernel(int32_t *piv) { piv[0] = 0x7fffffff; piv[1] = 0x7fffffff; }

// ---------------------------------------------------------------- 2. ControlledRoad candidates (:1439-1452)
__global__ void __launch_bounds__(256) cr_bits_kernel(int H, Bits bp, int32_t *__restrict__ cnt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nw = (long long)bp.wp * H;
    if (i >= nw) return;
    const int y = (int)((uint32_t)i / (uint32_t)bp.wp), wx = (int)((uint32_t)i - (uint32_t)y * (uint32_t)bp.wp);   // word indices fit 32 bits
    const u64 I0 = bp.I[i];
    const u64 Iup = y + 1 < H ? bp.I[i + bp.wp] : 0ull, Idn = y > 0 ? bp.I[i - bp.wp] : 0ull;
    const u64 Inx = wx + 1 < bp.wp ? bp.I[i + 1] : 0ull, Ipv = wx > 0 ? bp.I[i - 1] : 0ull;
    const u64 cr = bp.R[i] & ((bp.aN[i] & Iup) | (bp.aS[i] & Idn) | (bp.aE[i] & ((I0 >> 1) | (Inx << 63))) | (bp.aW[i] & ((I0 << 1) | (Ipv >> 63))));
    bp.cr[i] = cr;
    cnt[i] = __popcll(cr);
}

__global__ void __launch_bounds__(256) bit_count_kernel(long long nw, const u64 *__restrict__ plane, int32_t *__restrict__ cnt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nw) cnt[i] = __popcll(plane[i]);
}

// cells of the set bits, ascending cell index; prefix = exclusive scan of the word popcounts
__global__ void __launch_bounds__(256) bit_list_kernel(int W, int H, int wp, const u64 *__restrict__ plane, const int32_t *__restrict__ prefix,
                                                       int32_t *__restrict__ list, int cap, int32_t *err, int err_code) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)wp * H) return;
    u64 m = plane[i];
    if (!m) return;
    const int y = (int)((uint32_t)i / (uint32_t)wp), wx = (int)((uint32_t)i - (uint32_t)y * (uint32_t)wp);   // word indices fit 32 bits
    int k = prefix[i];
    const int base = y * W + wx * 64;
    while (m) {
        const int b = __ffsll((long long)m) - 1;
        m &= m - 1;
        if (k < cap) list[k] = base + b; else *err = err_code;
        k++;
    }
}

// ---------------------------------------------------------------- 3. reach
// pivot: device scalar(s).  own == true: the candidates of lights_bits_kernel ([0] first at or after the middle row,
// [1] first overall); own == false: one window cell index chosen by the caller (negative: the pivot lies outside this window)
__global__ void reach_seed_kernel(int W, const int32_t *piv, bool own, Bits bp) {
    int p = piv[0];
    if (own) { if (p == 0x7fffffff) p = piv[1]; if (p == 0x7fffffff) return; }   // no intersection at all: every query goes to the exact search
    if (p < 0) return;
    const int x = p % W, y = p / W;
    bp.fw[(size_t)y * bp.wp + (x >> 6)] = 1ull << (x & 63);
    bp.bw[(size_t)y * bp.wp + (x >> 6)] = 1ull << (x & 63);
}

// Occluded fills inside a 64-bit word: every set bit of f spreads upwards (fill_up) / downwards (fill_down) through
// the consecutive cells of p, p = cells that may be entered from the lower / higher neighbour.  One subtraction does
// it: with o = blockers | sources, o - 2f lets the borrow of every source ripple up to its first blocker, so
// o ^ (o - 2f) is the set of cells passed (the chess-engine "o ^ (o - 2r)" ray trick; checked exhaustively against the
// Kogge-Stone form on the host).
__device__ __forceinline__ u64 fill_up(u64 f, u64 p) {
    const u64 o = ~p | f;
    return f | ((o ^ (o - (f << 1))) & p);
}
__device__ __forceinline__ u64 fill_down(u64 f, u64 p) { return __brevll(fill_up(__brevll(f), __brevll(p))); }

// carry chain over the 32 words of a warp: c[0] = cin, c[i+1] = g[i] | (p[i] & c[i]) -- exactly a binary
// adder's carries, so one 64-bit add resolves all of them.  Returns the carry INTO every lane.
__device__ __forceinline__ uint32_t carry_chain(uint32_t g, uint32_t p, uint32_t cin, uint32_t &cout) {
    const u64 X = (u64)(g | (p & ~g)), Y = (u64)g;
    const u64 C = (X + Y + cin) ^ X ^ Y;
    cout = (uint32_t)(C >> 32) & 1u;
    return (uint32_t)C;
}

// Closes one LINE of `wp` words under the arrows along it: aUp = arrow towards the next cell (E in a row of the
// row-major planes, N in a row of the transposed planes), aDn = arrow towards the previous cell.  `f` moves WITH the
// arrows, `b` AGAINST them.  One warp per line; returns true if anything changed.
__device__ bool line_closure(u64 *__restrict__ fpl, u64 *__restrict__ bpl, const u64 *__restrict__ up, const u64 *__restrict__ dn, size_t base, int wp,
                             int nbits, int lane, uint8_t *__restrict__ dirty /* + word * dstride: the 64x64 block the word belongs to */, int dstride) {
    bool any = false;
    const u64 tail = (nbits & 63) ? ((1ull << (nbits & 63)) - 1ull) : ~0ull;   // an arrow out of the grid must not set a padding bit
    // Alternate low -> high and high -> low passes.  The line is closed as soon as a pass changes nothing and the
    // opposite pass has run on the same state (so: at least two passes, stop at the first quiet one).
    for (int pass = 0; pass < 256; pass++) {
        bool ch = false;
        uint32_t cf = 0, cb = 0;
        if ((pass & 1) == 0) {   // ---- low -> high: f along aUp, b against aDn
            for (int c0 = 0; c0 < wp; c0 += 32) {
                const int w = c0 + lane;
                const bool in = w < wp;
                const u64 aE = in ? up[base + w] : 0ull, aW = in ? dn[base + w] : 0ull;
                const u64 f = in ? __ldcg(fpl + base + w) : 0ull, b = in ? __ldcg(bpl + base + w) : 0ull;
                u64 f1 = fill_up(f, aE << 1), b1 = fill_up(b, aW);
                const uint32_t gf = __ballot_sync(0xffffffffu, (f1 & aE) >> 63), pf = __ballot_sync(0xffffffffu, aE == ~0ull);
                const uint32_t gb = __ballot_sync(0xffffffffu, b1 >> 63), pb = __ballot_sync(0xffffffffu, aW == ~0ull);
                uint32_t nf, nb;
                const uint32_t inf = carry_chain(gf, pf, cf, nf), inb = carry_chain(gb, pb, cb, nb);
                cf = nf; cb = nb;
                if ((inf >> lane) & 1u) f1 = fill_up(f1 | 1ull, aE << 1);
                if ((inb >> lane) & 1u) b1 = fill_up(b1 | (aW & 1ull), aW);
                if (w == wp - 1) { f1 &= tail; b1 &= tail; }
                if (in && (f1 != f || b1 != b)) { fpl[base + w] = f1; bpl[base + w] = b1; dirty[(size_t)w * dstride] = 1; ch = true; }
            }
        } else {                 // ---- high -> low: f along aDn, b against aUp
            for (int c0 = ((wp - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
                const int w = c0 + lane;
                const bool in = w < wp;
                const u64 aE = in ? up[base + w] : 0ull, aW = in ? dn[base + w] : 0ull;
                const u64 f = in ? __ldcg(fpl + base + w) : 0ull, b = in ? __ldcg(bpl + base + w) : 0ull;
                u64 f1 = fill_down(f, aW >> 1), b1 = fill_down(b, aE);
                const uint32_t gf = __brev(__ballot_sync(0xffffffffu, f1 & aW & 1ull)), pf = __brev(__ballot_sync(0xffffffffu, aW == ~0ull));
                const uint32_t gb = __brev(__ballot_sync(0xffffffffu, b1 & 1ull)), pb = __brev(__ballot_sync(0xffffffffu, aE == ~0ull));
                uint32_t nf, nb;
                const uint32_t inf = __brev(carry_chain(gf, pf, cf, nf)), inb = __brev(carry_chain(gb, pb, cb, nb));
                cf = nf; cb = nb;
                if ((inf >> lane) & 1u) f1 = fill_down(f1 | (1ull << 63), aW >> 1);
                if ((inb >> lane) & 1u) b1 = fill_down(b1 | (aE & (1ull << 63)), aE);
                if (w == wp - 1) { f1 &= tail; b1 &= tail; }
                if (in && (f1 != f || b1 != b)) { fpl[base + w] = f1; bpl[base + w] = b1; dirty[(size_t)w * dstride] = 1; ch = true; }
            }
        }
        ch = __any_sync(0xffffffffu, ch);
        any |= ch;
        if (!ch && pass >= 1) break;
    }
    return any;
}

// What a carry INTO a word adds, without running the fill again: the carry enters at the lowest (highest) cell and runs
// through the consecutive arrows from there.  (Checked against fill_up / fill_down on the host.)
__device__ __forceinline__ u64 low_run_incl(u64 a) { return a ^ (a + 1ull); }                 // cells 0 .. first cell without an arrow, inclusive
__device__ __forceinline__ u64 low_run(u64 a) { return a & ~(a + 1ull); }                     // the arrows below the first missing one
__device__ __forceinline__ u64 high_run_incl(u64 a) { const u64 m = ~a; return m ? (~0ull << (63 - __clzll((long long)m))) : ~0ull; }
__device__ __forceinline__ u64 high_run(u64 a) { const u64 m = ~a; return m ? ~(~0ull >> __clzll((long long)m)) : ~0ull; }

// The same closure with the whole line in REGISTERS (lines of up to 32 * NCH words): the words of the four planes are
// fetched with independent loads (one memory latency per line instead of one per chunk and pass), the passes run on
// registers, and only the words that changed are written back.  A pass only runs when its FRONTIER is not empty (some
// set cell has an arrow of that direction into an unset cell -- a few bit operations and ballots per chunk instead of a
// whole pass), so a line that is already closed costs two tests and no pass, and no pass is spent on noticing the end.
template <int NCH>
__device__ bool line_closure_reg(u64 *__restrict__ fpl, u64 *__restrict__ bpl, const u64 *__restrict__ up, const u64 *__restrict__ dn, size_t base, int wp,
                                 int nbits, int lane, uint8_t *__restrict__ dirty, int dstride) {
    constexpr uint32_t FULL = 0xffffffffu;
    u64 f[NCH], b[NCH], aU[NCH], aD[NCH];
    bool live = false;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        const int w = c * 32 + lane;
        const bool in = w < wp;
        f[c] = in ? __ldcg(fpl + base + w) : 0ull; b[c] = in ? __ldcg(bpl + base + w) : 0ull;
        aU[c] = in ? up[base + w] : 0ull; aD[c] = in ? dn[base + w] : 0ull;
        live |= (f[c] | b[c]) != 0ull;
    }
    // all four planes' loads are in flight before the first use: one memory round trip per line, not two (the profile showed
    // the arrow loads sunk below the early exit and a second full latency exposed on every live line)
#pragma unroll
    for (int c = 0; c < NCH; c++) asm volatile("" ::"l"(aU[c]), "l"(aD[c]));
    if (!__any_sync(FULL, live)) return false;   // nothing on this line yet: nothing can spread
    const u64 tail = (nbits & 63) ? ((1ull << (nbits & 63)) - 1ull) : ~0ull;
    uint32_t allU[NCH], allD[NCH];   // lanes whose word has the arrow on every cell: a carry runs straight through
#pragma unroll
    for (int c = 0; c < NCH; c++) { allU[c] = __ballot_sync(FULL, aU[c] == ~0ull); allD[c] = __ballot_sync(FULL, aD[c] == ~0ull); }

    auto frontier_up = [&]() {     // f moves with aU, b against aD, towards the higher cells
        uint32_t hit = 0, cf = 0, cb = 0;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const int w = c * 32 + lane;
            const bool in = w < wp;
            const u64 fa = f[c] & aU[c];
            u64 inner = ((fa << 1) & ~f[c]) | ((b[c] << 1) & aD[c] & ~b[c]);
            if (w == wp - 1) inner &= tail;
            const uint32_t sf = __ballot_sync(FULL, fa >> 63), df = __ballot_sync(FULL, in && !(f[c] & 1ull));
            const uint32_t sb = __ballot_sync(FULL, b[c] >> 63), db = __ballot_sync(FULL, in && ((aD[c] & ~b[c]) & 1ull));
            hit |= __ballot_sync(FULL, inner != 0ull) | (((sf << 1) | cf) & df) | (((sb << 1) | cb) & db);
            cf = sf >> 31; cb = sb >> 31;
        }
        return hit != 0;
    };
    auto frontier_down = [&]() {   // f moves with aD, b against aU, towards the lower cells
        uint32_t hit = 0, cf = 0, cb = 0;
#pragma unroll
        for (int c = NCH - 1; c >= 0; c--) {
            const int w = c * 32 + lane;
            const bool in = w < wp;
            const u64 fa = f[c] & aD[c];
            const u64 inner = ((fa >> 1) & ~f[c]) | ((b[c] >> 1) & aU[c] & ~b[c]);
            const uint32_t sf = __ballot_sync(FULL, fa & 1ull), df = __ballot_sync(FULL, in && !(f[c] >> 63));
            const uint32_t sb = __ballot_sync(FULL, b[c] & 1ull), db = __ballot_sync(FULL, in && ((aU[c] & ~b[c]) >> 63));
            hit |= __ballot_sync(FULL, inner != 0ull) | (((sf >> 1) | (cf << 31)) & df) | (((sb >> 1) | (cb << 31)) & db);
            cf = sf & 1u; cb = sb & 1u;
        }
        return hit != 0;
    };

    uint32_t chm = 0;   // bit c: my word of chunk c changed
    bool up_closed = false, dn_closed = false;
    for (int guard = 0; guard < 128 && !(up_closed && dn_closed); guard++) {
        if (!up_closed) {
            if (frontier_up()) {
                uint32_t cf = 0, cb = 0;
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    const int w = c * 32 + lane;
                    u64 f1 = fill_up(f[c], aU[c] << 1), b1 = fill_up(b[c], aD[c]);
                    const uint32_t gf = __ballot_sync(FULL, (f1 & aU[c]) >> 63), gb = __ballot_sync(FULL, b1 >> 63);
                    uint32_t nf, nb;
                    const uint32_t inf = carry_chain(gf, allU[c], cf, nf), inb = carry_chain(gb, allD[c], cb, nb);
                    cf = nf; cb = nb;
                    if ((inf >> lane) & 1u) f1 |= low_run_incl(aU[c]);
                    if ((inb >> lane) & 1u) b1 |= low_run(aD[c]);
                    if (w == wp - 1) { f1 &= tail; b1 &= tail; }
                    if (w < wp && (f1 != f[c] || b1 != b[c])) { f[c] = f1; b[c] = b1; chm |= 1u << c; }
                }
                dn_closed = false;
            }
            up_closed = true;
        }
        if (!dn_closed) {
            if (frontier_down()) {
                uint32_t cf = 0, cb = 0;
#pragma unroll
                for (int c = NCH - 1; c >= 0; c--) {
                    const int w = c * 32 + lane;
                    u64 f1 = fill_down(f[c], aD[c] >> 1), b1 = fill_down(b[c], aU[c]);
                    const uint32_t gf = __brev(__ballot_sync(FULL, f1 & aD[c] & 1ull)), gb = __brev(__ballot_sync(FULL, b1 & 1ull));
                    uint32_t nf, nb;
                    const uint32_t inf = __brev(carry_chain(gf, __brev(allD[c]), cf, nf)), inb = __brev(carry_chain(gb, __brev(allU[c]), cb, nb));
                    cf = nf; cb = nb;
                    if ((inf >> lane) & 1u) f1 |= high_run_incl(aD[c]);
                    if ((inb >> lane) & 1u) b1 |= high_run(aU[c]);
                    if (w == wp - 1) { f1 &= tail; b1 &= tail; }
                    if (w < wp && (f1 != f[c] || b1 != b[c])) { f[c] = f1; b[c] = b1; chm |= 1u << c; }
                }
                up_closed = false;
            }
            dn_closed = true;
        }
    }
#pragma unroll
    for (int c = 0; c < NCH; c++)
        if ((chm >> c) & 1u) { const int w = c * 32 + lane; fpl[base + w] = f[c]; bpl[base + w] = b[c]; dirty[(size_t)w * dstride] = 1; }
    return __any_sync(FULL, chm != 0);
}

// ---- reachability as a sequence of ordinary launches (one per phase, each with its own grid and occupancy) ----------
// ctl[0] = something changed in this alternation, ctl[1] = done, ctl[2] = alternations run.  Every phase kernel returns
// at once when ctl[1] is set, so the host can enqueue a fixed number of alternations without reading anything back.
struct ReachT;   // below

template <int NCH>
__global__ void __launch_bounds__(256, 2) reach_lines_kernel(u64 *fpl, u64 *bpl, const u64 *__restrict__ up, const u64 *__restrict__ dn, int nlines, int wp,
                                                          int nbits, const uint8_t *line_dirty, bool all, uint8_t *bd, int bd_line_stride,
                                                          int bd_word_stride, int32_t *ctl) {
    // ctl[1] was written by the previous launch and is constant during this one: a cached read-only load.  (As a volatile
    // load it was served by L2 for each of the 16 k warps and was the top stall of this kernel: 46 % of the samples.)
    if (__ldg(ctl + 1)) return;
    const int line = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (line >= nlines) return;
    if (!all && !line_dirty[line >> 6]) return;
    uint8_t *dirty = bd + (size_t)(line >> 6) * bd_line_stride;
    const bool ch = NCH > 0 ? line_closure_reg<(NCH > 0 ? NCH : 1)>(fpl, bpl, up, dn, (size_t)line * wp, wp, nbits, lane, dirty, bd_word_stride)
                            : line_closure(fpl, bpl, up, dn, (size_t)line * wp, wp, nbits, lane, dirty, bd_word_stride);
    if (ch && lane == 0) ctl[0] = 1;
}

// 512 x 512-cell tiles of the SOURCE plane pair that hold a dirty 64 x 64 block (or all tiles) -> destination pair; marks the
// destination's lines.  Dynamic shared memory: 2 x [512][9] words.
__global__ void __launch_bounds__(256, 3) reach_transpose_kernel(const u64 *srcA, const u64 *srcB, int src_rows, int src_wp, u64 *dstA, u64 *dstB, int dst_rows,
                                                              int dst_wp, uint8_t *bd /* [nby][nbx], row-major block grid */, int nbx, int nby,
                                                              bool src_is_rowmajor, bool all, bool compare, uint8_t *dst_line_dirty,
                                                              uint8_t *clear_flags, int n_clear, const int32_t *ctl) {
    extern __shared__ u64 s_tile[];
    u64 (*s_in)[9] = reinterpret_cast<u64 (*)[9]>(s_tile);
    u64 (*s_out)[9] = reinterpret_cast<u64 (*)[9]>(s_tile + 512 * 9);
    if (__ldg(ctl + 1)) return;
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < n_clear; i += blockDim.x) clear_flags[i] = 0;   // the previous phase has read them
    const int ntx = (nbx + 7) >> 3;
    const int tx = blockIdx.x % ntx, ty = blockIdx.x / ntx;   // tile of the row-major block grid (8 x 8 blocks)
    if (!all) {
        bool any = false;
        for (int k = threadIdx.x; k < 64; k += blockDim.x) {
            const int by = ty * 8 + (k >> 3), bx = tx * 8 + (k & 7);
            if (by < nby && bx < nbx && bd[(size_t)by * nbx + bx]) { any = true; bd[(size_t)by * nbx + bx] = 0; }   // each flag has one reader
        }
        if (!__syncthreads_or(any)) return;
    }
    // a tile (tx, ty) of the row-major grid is tile (ty, tx) of the transposed planes
    const int sx = src_is_rowmajor ? tx : ty, sy = src_is_rowmajor ? ty : tx;
    Tile512Regs qa, qb;
    tile512_load(srcA, src_rows, src_wp, sx, sy, true, qa);
    tile512_stage(qa, s_in);
    tile512_load(srcB, src_rows, src_wp, sx, sy, true, qb);   // in flight while plane A is transposed and written
    __syncthreads();
    bool tch = tile512_finish(dstA, dst_rows, dst_wp, sx, sy, compare, s_in, s_out);
    tile512_stage(qb, s_in);
    __syncthreads();
    tch |= tile512_finish(dstB, dst_rows, dst_wp, sx, sy, compare, s_in, s_out);
    if (tch && threadIdx.x < 8) {
        const int k = (src_is_rowmajor ? tx : ty) * 8 + threadIdx.x;   // destination lines = source bit columns
        if (k < (src_is_rowmajor ? nbx : nby)) dst_line_dirty[k] = 1;
    }
}

// ctl[1] = "done": set from the start when nobody needs the planes (*gate == 0), so that every phase kernel returns at once
__global__ void reach_gate_kernel(int32_t *ctl, const int32_t *gate) { ctl[0] = 0; ctl[1] = (gate && *gate == 0) ? 1 : 0; ctl[2] = 0; ctl[3] = 0; }

__global__ void reach_ctl_kernel(int32_t *ctl, int32_t *changed) {
    if (ctl[1]) return;
    ctl[2]++;
    if (ctl[0]) { if (changed) *changed = 1; } else ctl[1] = 1;
    ctl[0] = 0;
}

// resumed call: only `edge_rows` rows at each end of the window received bits from outside
__global__ void __launch_bounds__(256) reach_mark_edges_kernel(int H, int edge_rows, int nbx, int nby, uint8_t *bdR, uint8_t *rd) {
    const int lo_blocks = (min(edge_rows, H) + 63) >> 6, hi_first = max(H - edge_rows, 0) >> 6;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nby * nbx; i += gridDim.x * blockDim.x) {
        const int by = i / nbx;
        if (by < lo_blocks || by >= hi_first) { bdR[i] = 1; rd[by] = 1; }
    }
}

// Persistent cooperative kernel: alternate row closures of the row-major planes and row closures of the TRANSPOSED
// planes (= column closures) until an alternation changes nothing.  Both sweeps are the same warp-per-line kernel
// with fully coalesced word loads.  Work follows the frontier: a sweep marks the 64 x 64 blocks it changed, only
// those are re-transposed, and only the lines that cross a re-transposed block are closed again (the first
// alternation of a launch does everything once, whatever was seeded or merged into the planes from outside).
struct ReachT {
    u64 *aNt, *aSt, *fwT, *bwT;   // transposed planes: [W][wpT], bit y & 63 of word y >> 6
    uint8_t *bdR, *bdT;           // [nby][nbx] block changed in the row-major / transposed planes since it was last transposed
    uint8_t *rd, *cd;             // [nby] row block / [nbx] column block needs closing
    int wpT;
};


__global__ void __launch_bounds__(256) reach_kernel(int W, int H, Bits bp, ReachT rt, int edge_rows /* < 0: first call, everything is new; else only
                                                    that many rows at each end of the window received bits since the last call */,
                                                    const int32_t *skip_if_done, int32_t *flags /* [0..2] change flags, [3] alternations */, int32_t *changed, int32_t *err) {
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int warp = gtid >> 5, nwarps = nth >> 5;
    const int nbx = bp.wp, nby = rt.wpT;           // 64 x 64 blocks per row / per column of the row-major planes
    __shared__ u64 s_in[2][256][5], s_out[2][256][5];
    const int ntx = (nbx + 3) >> 2, nty = (nby + 3) >> 2;   // 256 x 256 tiles of the row-major planes
    if (skip_if_done && *((volatile const int32_t *)skip_if_done)) return;   // uniform over the grid, before any barrier
    if (edge_rows < 0) {
        for (int tile = blockIdx.x; tile < ntx * nty; tile += gridDim.x)
            transpose_tile2(bp.aN, bp.aS, H, bp.wp, rt.aNt, rt.aSt, W, rt.wpT, tile % ntx, tile / ntx, false, false, s_in, s_out);
    } else {
        // resumed call: the transposed arrows and planes of the last call are still valid; only the edge rows must be
        // re-closed and their blocks re-transposed
        const int lo_blocks = (min(edge_rows, H) + 63) >> 6, hi_first = max(H - edge_rows, 0) >> 6;
        for (int i = gtid; i < nby * nbx; i += nth) {
            const int by = i / nbx;
            if (by < lo_blocks || by >= hi_first) { rt.bdR[i] = 1; rt.rd[by] = 1; }
        }
        __threadfence();
        grid.sync();
    }
    // does tile (tx, ty) of the row-major block grid hold a dirty block?  (uniform over the CTA); clears the flags
    auto tile_dirty = [&](uint8_t *bd, int tx, int ty) {
        bool any = false;
        for (int by = ty * 4; by < min(ty * 4 + 4, nby); by++)
            for (int bx = tx * 4; bx < min(tx * 4 + 4, nbx); bx++) any |= *((volatile uint8_t *)(bd + (size_t)by * nbx + bx)) != 0;
        return any;
    };
    auto tile_clear = [&](uint8_t *bd, int tx, int ty) {
        if (threadIdx.x < 16) {
            const int by = ty * 4 + (threadIdx.x >> 2), bx = tx * 4 + (threadIdx.x & 3);
            if (by < nby && bx < nbx) bd[(size_t)by * nbx + bx] = 0;
        }
    };
    for (int it = 0;; it++) {
        int32_t *flag = flags + it % 3;
        const bool all = it == 0 && edge_rows < 0;
        bool ch = false;
        // ---- 1. close the rows whose block row received new bits
        for (int y = warp; y < H; y += nwarps)
            if (all || *((volatile uint8_t *)(rt.rd + (y >> 6))))
                ch |= line_closure(bp.fw, bp.bw, bp.aE, bp.aW, (size_t)y * bp.wp, bp.wp, W, lane, rt.bdR + (size_t)(y >> 6) * nbx, 1);
        __threadfence();
        grid.sync();
        // ---- 2. tiles with changed blocks -> transposed planes
        for (int i = gtid; i < nby; i += nth) rt.rd[i] = 0;
        for (int tile = blockIdx.x; tile < ntx * nty; tile += gridDim.x) {
            const int tx = tile % ntx, ty = tile / ntx;
            if (!all && !tile_dirty(rt.bdR, tx, ty)) continue;
            __syncthreads();   // every thread has read the flags
            tile_clear(rt.bdR, tx, ty);
            const bool tch = transpose_tile2(bp.fw, bp.bw, H, bp.wp, rt.fwT, rt.bwT, W, rt.wpT, tx, ty, true, it == 0 && edge_rows >= 0, s_in, s_out);
            if (tch && threadIdx.x < 4 && tx * 4 + (int)threadIdx.x < nbx) rt.cd[tx * 4 + threadIdx.x] = 1;
        }
        __threadfence();
        grid.sync();
        // ---- 3. close the columns (lines of the transposed planes) that cross a re-transposed tile
        for (int x = warp; x < W; x += nwarps)
            if (all || *((volatile uint8_t *)(rt.cd + (x >> 6))))
                ch |= line_closure(rt.fwT, rt.bwT, rt.aNt, rt.aSt, (size_t)x * rt.wpT, rt.wpT, H, lane, rt.bdT + (x >> 6), nbx);
        if (__syncthreads_or(ch) && threadIdx.x == 0) *flag = 1;
        __threadfence();
        grid.sync();
        // ---- 4. tiles with changed blocks -> row-major planes (tile (tx, ty) of the row-major grid = tile (ty, tx) of the transposed one)
        for (int i = gtid; i < nbx; i += nth) rt.cd[i] = 0;
        for (int tile = blockIdx.x; tile < ntx * nty; tile += gridDim.x) {
            const int tx = tile % ntx, ty = tile / ntx;
            if (!tile_dirty(rt.bdT, tx, ty)) continue;
            __syncthreads();
            tile_clear(rt.bdT, tx, ty);
            const bool tch = transpose_tile2(rt.fwT, rt.bwT, W, rt.wpT, bp.fw, bp.bw, H, bp.wp, ty, tx, true, false, s_in, s_out);
            if (tch && threadIdx.x < 4 && ty * 4 + (int)threadIdx.x < nby) rt.rd[ty * 4 + threadIdx.x] = 1;
        }
        __threadfence();
        grid.sync();
        const int any = *((volatile int32_t *)flag);
        if (blockIdx.x == 0 && threadIdx.x == 0) { flags[(it + 2) % 3] = 0; flags[3] = it + 1; if (any && changed) *changed = 1; }
        if (!any) break;
        if (it > 100000) { if (blockIdx.x == 0 && threadIdx.x == 0) *err = 12; break; }
    }
}

// ---------------------------------------------------------------- exact fallback searches
// An exact search that comes up empty after touching a shard cut may have missed a path through rows this window does not
// hold: that is reported (flag 13), never returned as "unreachable".
__device__ __forceinline__ bool on_cut(const LightsCtx &L, int y) { return (L.cut_lo && y == 0) || (L.cut_hi && y == L.H - 1); }
// does the result of the candidate in row y reach a row this shard owns?  (its lights lie within 2 rows, its scan cells
// within traffic_light_range + 1); candidates deeper in the halo are recomputed by their owner and overwritten
__device__ __forceinline__ bool matters(const LightsCtx &L, int y) {
    const int m = L.tl_range + 3;
    return y >= L.cut_lo - m && y < L.H - L.cut_hi + m;
}

__device__ bool bfs_forward(const LightsCtx &L, int from, int to, bool strict) {
    int q[BFS_CAP];
    int head = 0, tail = 0;
    bool cut = false;
    q[tail++] = from;
    while (head < tail) {
        const int c = q[head++];
        if (c == to) return true;
        const uint32_t d = L.D[c];
        const int x = c % L.W, y = c / L.W;
        cut |= on_cut(L, y);
        for (int i = 0; i < dl_len(d); i++) {
            const int k = dl_get(d, i), nx = x + dx_of(k), ny = y + dy_of(k);
            if (!L.has(nx, ny)) continue;
            const int j = L.at(nx, ny);
            bool seen = false;
            for (int u = 0; u < tail; u++) seen |= (q[u] == j);
            if (seen) continue;
            if (j != to && L.D[j] == 0) continue;   // arrow-less cells are dead ends of the BFS
            if (tail == BFS_CAP) { if (strict) *L.err = 10; return false; }
            q[tail++] = j;
        }
    }
    if (cut && strict) *L.err = 13;
    return false;
}

__device__ bool bfs_backward(const LightsCtx &L, int from, int to, bool strict) {   // does `from` reach `to`? search predecessors of `to`
    int q[BFS_CAP];
    int head = 0, tail = 0;
    bool cut = false;
    q[tail++] = to;
    while (head < tail) {
        const int c = q[head++];
        if (c == from) return true;
        const int x = c % L.W, y = c / L.W;
        cut |= on_cut(L, y);
        for (int k = 0; k < 4; k++) {   // predecessor p = c - dir(k) with arrow k
            const int nx = x - dx_of(k), ny = y - dy_of(k);
            if (!L.has(nx, ny)) continue;
            const int j = L.at(nx, ny);
            if (!dl_has(L.D[j], k)) continue;
            bool seen = false;
            for (int u = 0; u < tail; u++) seen |= (q[u] == j);
            if (seen) continue;
            if (tail == BFS_CAP) { if (strict) *L.err = 11; return false; }
            q[tail++] = j;
        }
    }
    if (cut && strict) *L.err = 13;
    return false;
}

// cell.py:201-227 `a.leads_to(b)`
__device__ __forceinline__ bool leads_to(const LightsCtx &L, int a, int b, bool strict) {
    if (a == b) return true;
    const int ax = a % L.W, ay = a / L.W, bx = b % L.W, by = b / L.W;
    const bool a_bw = L.bit(L.b.bw, ax, ay), b_fw = L.bit(L.b.fw, bx, by);
    if (a_bw && b_fw) return true;
    if (!a_bw) return bfs_forward(L, a, b, strict);
    return bfs_backward(L, a, b, strict);
}

// ---------------------------------------------------------------- 4. per ControlledRoad evaluation
__device__ __forceinline__ bool before_cm(int sx, int sy, int cx, int cy) { return sx < cx || (sx == cx && sy < cy); }   // column-major order

// cell type as the reference sees it when it visits (cx,cy)
__device__ __forceinline__ int type_at_time(const LightsCtx &L, int sx, int sy, int cx, int cy) {
    if (before_cm(sx, sy, cx, cy) && L.bit(L.b.cr, sx, sy)) return T_CR;
    return L.T[L.at(sx, sy)];
}

// ---- `leads_to` without reachability planes ---------------------------------------------------------------------------
// Almost every query of the reverse march is answered by the lane itself (the cell has an arrow onto the cell one step
// closer, which leads to the road).  What is left (0.6 % of the queries of a synthetic city: lane-change arrows make the
// march cross to the opposite carriageway, ring corners) asks for a real path a ~> c between two cells at most
// traffic_light_range + 1 apart on one line, c = a + k * p.  A path that is FOUND is an exact "True" whatever the rest of
// the city looks like, so the search runs in stages of growing cost and only what no stage can settle (and every
// "False") falls through to the reachability planes of the whole window:
//   1. Z witness (one thread): a runs straight along a lateral direction s, crosses to the line of c through k arrows p,
//      and runs back against s into c:  a -s-> a_j -p-> c_j -(-s)-> c.  On a two-way road this is the U-turn at the next
//      crossing; it settles all but a handful of queries per city.
//   2. window closure (one warp): the set of cells that reach c inside a 128 x 128-cell window around c, as the fixed
//      point of word fills on the arrow bit-planes (rows in registers, neighbours by shuffle).
constexpr int Z_REACH = 96;      // cells a Z witness may run along the lateral direction
constexpr int WIN_CELLS = 128;   // side of the closure window (2 words x 128 rows, 4 rows per lane)

__device__ bool z_witness(const LightsCtx &L, int ax, int ay, int cx, int cy, int p, int k) {
    for (int side = 0; side < 2; side++) {
        const int s = side ? ((p + 3) & 3) : right_of(p), o = opp_of(s);
        int x1 = ax, y1 = ay, x2 = cx, y2 = cy;   // a_{j-1}, c_{j-1}
        for (int j = 1; j <= Z_REACH; j++) {
            if (!dl_has(L.D[L.at(x1, y1)], s)) break;                 // a's run ends here
            x1 += dx_of(s); y1 += dy_of(s); x2 += dx_of(s); y2 += dy_of(s);
            if (!L.has(x1, y1) || !L.has(x2, y2)) break;
            if (!dl_has(L.D[L.at(x2, y2)], o)) break;                 // c_j must step towards c
            bool link = true;                                          // k arrows p from a_j land on c_j
            for (int m = 0, lx = x1, ly = y1; m < k && link; m++, lx += dx_of(p), ly += dy_of(p)) link = dl_has(L.D[L.at(lx, ly)], p);
            if (link) return true;
        }
    }
    return false;
}

// Result record of one ControlledRoad: bits 0..39 up to 8 light cells as 5-bit offsets (dy+2)*5+(dx+2),
// bits 40..59 reverse-scan length per entry of the ordered direction list (5 bits each), bits 60..63 the
// number of light cells.
// `lt(nb, c, dir, k)` answers nb.leads_to(c) for the cell k steps behind c against arrow `dir` when the lane itself does
// not: true / false; a functor that cannot decide remembers that and says false (the caller re-evaluates the candidate).
template <class LT>
__device__ u64 lights_eval(const LightsCtx &L, int cx, int cy, LT &lt) {
    const int c = L.at(cx, cy);
    const int t = L.T[c];
    const uint32_t rd = L.D[c];
    u64 rec = 0;
    int nacc = 0;
    auto push = [&](int ax, int ay) { rec |= (u64)((ay - cy + 2) * 5 + (ax - cx + 2)) << (5 * nacc); nacc++; };
    int vx[4], vy[4], nv = 0;
    for (int r = 0; r < dl_len(rd); r++) {   // cells to the right of every arrow, de-duplicated (:1465-1474)
        const int k = right_of(dl_get(rd, r)), bx = cx + dx_of(k), by = cy + dy_of(k);
        bool dup = false;
        for (int u = 0; u < nv; u++) dup |= (vx[u] == bx && vy[u] == by);
        if (!dup) { vx[nv] = bx; vy[nv] = by; nv++; }
    }
    for (int u = 0; u < nv; u++) {
        if (!L.has(vx[u], vy[u])) continue;
        // everything this neighbour can be asked is fetched together (independent loads: one memory round trip, not four)
        const int vc = L.at(vx[u], vy[u]);
        const int fx = 2 * vx[u] - cx, fy = 2 * vy[u] - cy;
        const bool f_in = L.has(fx, fy);
        const bool conv = before_cm(vx[u], vy[u], cx, cy) && L.bit(L.b.cr, vx[u], vy[u]);
        const int vt = L.T[vc], ft = f_in ? (int)L.T[L.at(fx, fy)] : -1;
        const uint32_t vd = L.D[vc];
        const int st = conv ? (int)T_CR : vt;   // type_at_time
        if (st == T_CR || st == t) {
            if (!(vd & rd & 0xf)) continue;   // shares no arrow (:1483)
            if (ft == T_SIDEWALK) push(fx, fy);
        }
        if (vt == T_SIDEWALK) push(vx[u], vy[u]);
    }
    if (nacc == 0) return 0;
    rec |= (u64)nacc << 60;
    int depth = 0;   // budget shared by all directions (:1528-1548)
    for (int i = 0; i < dl_len(rd); i++) {
        const int fd = dl_get(rd, i), k = opp_of(fd);
        int bx = cx + dx_of(k), by = cy + dy_of(k), cnt = 0;
        // type, arrows and candidate bit of a cell travel together, and the NEXT cell's are requested before this cell's are
        // looked at: two steps of the march are in flight at any time (the march is a chain of dependent round trips otherwise;
        // fetching all twelve cells up front was tried and lost to its own traffic)
        struct Cell { bool in, conv; int t; uint32_t d; };
        auto fetch = [&](int x, int y) {
            Cell q{L.has(x, y), false, -1, 0u};
            if (q.in) { const int i = L.at(x, y); q.conv = before_cm(x, y, cx, cy) && L.bit(L.b.cr, x, y); q.t = L.T[i]; q.d = L.D[i]; }
            return q;
        };
        Cell cur = fetch(bx, by);
        while (depth <= L.tl_range) {
            if (!cur.in) break;
            const Cell nxt = depth < L.tl_range ? fetch(bx + dx_of(k), by + dy_of(k)) : Cell{false, false, -1, 0u};
            if ((cur.conv ? (int)T_CR : cur.t) != t) break;   // type_at_time
            // the cell one step closer to c leads to c (that is why the march got here), so a cell with an arrow onto it does too;
            // only the others (lane-change arrows, opposite lanes) need a search
            if (!dl_has(cur.d, fd) && !lt(L.at(bx, by), c, fd, cnt + 1)) break;
            cnt++;
            bx += dx_of(k); by += dy_of(k); depth++;
            cur = nxt;
        }
        rec |= (u64)cnt << (40 + 5 * i);
    }
    return rec;
}

struct LtPlanes {   // the reachability planes of the window + exact bounded searches (leads_to above)
    const LightsCtx &L; bool strict;
    __device__ bool operator()(int nb, int c, int, int) const { return leads_to(L, nb, c, strict); }
};

struct LtWitness {   // stage 1
    const LightsCtx &L; bool undecided;
    __device__ bool operator()(int nb, int c, int fd, int k) {
        if (!(L.stages & 1) && z_witness(L, nb % L.W, nb / L.W, c % L.W, c / L.W, fd, k)) return true;
        undecided = true;
        return false;
    }
};

struct LtWindow {   // stage 2: bit (x, y) of `win` (rows y0.., words wx0, wx0 + 1) = that cell reaches c inside the window
    const LightsCtx &L; const u64 (*win)[2]; int wx0, y0; bool undecided;
    __device__ bool operator()(int nb, int, int, int) {
        const int x = nb % L.W - wx0 * 64, y = nb / L.W - y0;
        if (x >= 0 && x < WIN_CELLS && y >= 0 && y < WIN_CELLS && ((win[y][x >> 6] >> (x & 63)) & 1ull)) return true;
        undecided = true;   // not reached inside the window: the planes decide (a path may leave the window)
        return false;
    }
};

// cell.py:229-239
__device__ __forceinline__ bool directly_leads_to(const LightsCtx &L, int from, int to) {
    const uint32_t d = L.D[from];
    const int x = from % L.W, y = from / L.W;
    for (int i = 0; i < dl_len(d); i++) {
        const int k = dl_get(d, i), nx = x + dx_of(k), ny = y + dy_of(k);
        if (L.has(nx, ny) && L.at(nx, ny) == to) return true;
    }
    return false;
}

// _scan_for_traffic_flow_forward (city_model.py:1550-1584) for the ControlledRoad (cx, cy), currently scanning from
// "road" (rx, ry): emit(cell) for every assigned outgoing cell, in the reference's order.  Recursion depth is bounded
// by traffic_light_range + 1.  The reference sees (cx, cy) itself and every earlier-converted candidate as
// ControlledRoad (type_at_time), everything else with its original type.
template <class F>
__device__ void forward_scan(const LightsCtx &L, int cx, int cy, int rx, int ry, uint32_t sdirs, int orig, int scan_depth, F &emit) {
    const int road = L.at(rx, ry);
    for (int i = 0; i < dl_len(sdirs); i++) {
        const int rd = dl_get(sdirs, i);
        int bx = rx + dx_of(rd), by = ry + dy_of(rd), depth = scan_depth;
        while (depth <= L.tl_range) {
            if (!L.has(bx, by)) break;
            const int c = L.at(bx, by);
            const int t = (bx == cx && by == cy) ? T_CR : type_at_time(L, bx, by, cx, cy);
            if (t == T_INTER) {
                if (L.fwd_mode == 1) { emit(c); depth++; }
                else if (L.fwd_mode == 2) emit(c);
            } else if (t == orig) {
                if (directly_leads_to(L, c, road)) forward_scan(L, cx, cy, bx, by, sdirs, orig, depth + 1, emit);
                else if (dl_has(L.D[c], rd)) { emit(c); depth++; }
            } else {
                break;
            }
            bx += dx_of(rd); by += dy_of(rd);
        }
    }
}

__device__ __forceinline__ int rec_nacc(u64 r) { return (int)(r >> 60); }
__device__ __forceinline__ int rec_acc(u64 r, int u, int c, int W) {
    const int code = (int)(r >> (5 * u)) & 31;
    return c + (code / 5 - 2) * W + (code % 5 - 2);
}
__device__ __forceinline__ int rec_cnt(u64 r, int i) { return (int)(r >> (40 + 5 * i)) & 31; }
__device__ __forceinline__ int rec_nsc(u64 r) { return rec_cnt(r, 0) + rec_cnt(r, 1) + rec_cnt(r, 2) + rec_cnt(r, 3); }

__device__ __forceinline__ void or_byte(uint8_t *p, uint32_t bits) {
    uint32_t *w = (uint32_t *)((uintptr_t)p & ~(uintptr_t)3);
    atomicOr(w, bits << (8 * ((uintptr_t)p & 3)));
}

// controlled_road.light = tl (:1517) and nb.light = tl for the cells of the reverse march (:1542; lost again if nb itself is
// converted later -- a fresh CellAgent -- so candidates are skipped): the aux updates a finished record implies.  The cell
// TYPE of the candidate changes later (cr_type_kernel): other candidates still read the original types.
__device__ __forceinline__ void apply_aux(const LightsCtx &L, int c, u64 r, uint8_t *A) {
    const int na = rec_nacc(r);
    if (na) {
        // The march cells are MARKED in a bit-plane (L2-resident, and the cells of a march along a row share one or two words: one or
        // two atomics instead of one per cell); aux_light_merge_kernel ORs the marks into the aux plane once all records are there.
        // (As byte atomics on the aux plane itself this was the evaluation's top stall: 60 M read-modify-writes of DRAM sectors.)
        const uint32_t rd = L.D[c];
        const int cx = c % L.W, cy = c / L.W;
        for (int d = 0; d < dl_len(rd); d++) {
            const int k = opp_of(dl_get(rd, d)), cnt = rec_cnt(r, d);
            if (!cnt) continue;
            if (dy_of(k) == 0) {
                const int x0 = dx_of(k) > 0 ? cx + 1 : cx - cnt, x1 = x0 + cnt - 1;   // cnt <= 31: at most two words
                u64 *row = L.b.lm + (size_t)cy * L.b.wp;
                const int w0 = x0 >> 6, w1 = x1 >> 6;
                const u64 m0 = ~0ull << (x0 & 63), m1 = ~0ull >> (63 - (x1 & 63));
                if (w0 == w1) atomicOr(row + w0, m0 & m1);
                else { atomicOr(row + w0, m0); atomicOr(row + w1, m1); }
            } else {
                const u64 bit = 1ull << (cx & 63);
                for (int s = 1; s <= cnt; s++) atomicOr(L.b.lm + (size_t)(cy + s * dy_of(k)) * L.b.wp + (cx >> 6), bit);
            }
        }
    }
    A[c] = (uint8_t)((A[c] & (AUX_RING | AUX_EVER)) | (na ? AUX_LIGHT : 0) | L.T[c]);
}

// marks of the reverse marches -> AUX_LIGHT, except on candidates (a converted cell is a fresh CellAgent: it only carries the light of
// its own record, written by apply_aux).  One thread per word of the plane: it owns the 64 aux bytes it may touch.
__global__ void __launch_bounds__(256) aux_light_merge_kernel(int W, long long nw, int wp, const u64 *__restrict__ lm, const u64 *__restrict__ cr,
                                                              uint8_t *__restrict__ A) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nw) return;
    const u64 m = lm[i] & ~cr[i];
    if (!m) return;
    const uint32_t y = (uint32_t)i / (uint32_t)wp, wx = (uint32_t)i - y * (uint32_t)wp;   // word indices fit 32 bits
    uint8_t *row = A + (size_t)y * W + (size_t)wx * 64;
    if ((W & 15) == 0) {
        auto spread = [](uint32_t b) { return (((b | (b << 7) | (b << 14) | (b << 21)) & 0x01010101u) * (uint32_t)AUX_LIGHT); };   // 4 bits -> AUX_LIGHT in 4 bytes
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t mm = (uint32_t)(m >> (16 * q)) & 0xffffu;
            if (!mm) continue;
            uint4 v = *reinterpret_cast<uint4 *>(row + 16 * q);
            v.x |= spread(mm & 15u); v.y |= spread((mm >> 4) & 15u); v.z |= spread((mm >> 8) & 15u); v.w |= spread(mm >> 12);
            *reinterpret_cast<uint4 *>(row + 16 * q) = v;
        }
    } else {
        u64 rest = m;
        while (rest) { const int b = __ffsll((long long)rest) - 1; rest &= rest - 1; row[b] |= AUX_LIGHT; }
    }
}

__device__ __forceinline__ void mark_lights(const LightsCtx &L, int c, u64 r) {
    for (int u = 0; u < rec_nacc(r); u++) {
        const int a = rec_acc(r, u, c, L.W);
        const int ax = a % L.W, ay = a / L.W;
        atomicOr(L.b.tl + (size_t)ay * L.b.wp + (ax >> 6), 1ull << (ax & 63));
    }
}

// MODE 0: every candidate, `leads_to` from Z witnesses; candidates with an undecided query go to the `pend` list.
// MODE 1: every candidate, from the reachability planes (the staged entry points: the caller has closed them).
// MODE 2: the candidates of the list `sel` (what stages 1 and 2 left over), from the reachability planes.
template <int MODE>
__global__ void __launch_bounds__(128) lights_eval_kernel(LightsCtx L, const int32_t *__restrict__ n_cr, const int32_t *__restrict__ cr_cell,
                                                          u64 *__restrict__ rec, uint8_t *A, const int32_t *__restrict__ sel,
                                                          int32_t *n_pend, int32_t *pend, int cap_pend) {
    const int n = MODE == 2 ? min(*n_cr, cap_pend) : *n_cr;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int i = MODE == 2 ? sel[j] : j;
        const int c = cr_cell[i];
        if (MODE == 0) {
            LtWitness lt{L, false};
            const u64 r = lights_eval(L, c % L.W, c / L.W, lt);
            rec[i] = r;
            mark_lights(L, c, r);   // the light cells never depend on `leads_to`
            if (!lt.undecided) { apply_aux(L, c, r, A); continue; }
            if (!matters(L, c / L.W)) { apply_aux(L, c, r, A); continue; }   // deep in a neighbour's rows: its owner computes it, nothing here reads it
            const int k = atomicAdd(n_pend, 1);
            if (k < cap_pend) pend[k] = i; else *L.err = 28;
        } else {
            LtPlanes lt{L, matters(L, c / L.W)};
            const u64 r = lights_eval(L, c % L.W, c / L.W, lt);
            rec[i] = r;
            if (MODE == 1) mark_lights(L, c, r);
            apply_aux(L, c, r, A);
        }
    }
}

// Stage 2: one warp per candidate stage 1 left undecided.  B = cells that reach c inside the window = least fixed point of
//   B |= aE & (B >> 1)  |  aW & (B << 1)  (row fills, one subtraction per word)   |  aN & B[y + 1]  |  aS & B[y - 1].
// Lane l keeps rows 4l .. 4l + 3 (two words each) of B and of the four arrow planes in registers; the rows of the
// neighbouring lanes arrive by shuffle.  A converged window goes to shared memory and lane 0 evaluates the candidate.
__global__ void __launch_bounds__(128) lights_window_kernel(LightsCtx L, const int32_t *__restrict__ n_pend, const int32_t *__restrict__ pend,
                                                            const int32_t *__restrict__ cr_cell, u64 *__restrict__ rec, uint8_t *A,
                                                            int32_t *n_pend2, int32_t *pend2, int cap_pend) {
    constexpr uint32_t FULL = 0xffffffffu;
    __shared__ u64 s_win[4][WIN_CELLS][2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int n = min(*n_pend, cap_pend);
    for (int e = warp; e < n; e += nwarps) {
        const int i = pend[e];
        const int c = cr_cell[i], cx = c % L.W, cy = c / L.W;
        bool undecided = true;
        u64 r = 0;
        if (!(L.stages & 2)) {
            const int wp = L.b.wp;
            int wx0 = (cx >> 6) - ((cx & 63) < 32 ? 1 : 0);
            wx0 = max(0, min(wx0, wp - 2));
            const int y0 = max(0, min(cy - WIN_CELLS / 2, L.H - WIN_CELLS));
            u64 aN[4][2], aE[4][2], aS[4][2], aW[4][2], b[4][2];
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int w = 0; w < 2; w++) {
                    const int y = y0 + lane * 4 + q, wx = wx0 + w;
                    const bool in = y < L.H && wx < wp;
                    const size_t o = (size_t)y * wp + wx;
                    aN[q][w] = in ? L.b.aN[o] : 0ull; aE[q][w] = in ? L.b.aE[o] : 0ull;
                    aS[q][w] = in ? L.b.aS[o] : 0ull; aW[q][w] = in ? L.b.aW[o] : 0ull;
                    b[q][w] = (y == cy && wx == (cx >> 6)) ? 1ull << (cx & 63) : 0ull;
                }
            // an arrow out of the window must not pull anything in: the top row's N arrows and the bottom row's S arrows see
            // zeros through the shuffles below, the E / W arrows at the word ends see no carry
            for (int iter = 0; iter < 1024; iter++) {
                bool ch = false;
#pragma unroll
                for (int q = 0; q < 4; q++) {   // rows: towards higher x through W arrows, towards lower x through E arrows
                    u64 w0 = fill_up(b[q][0], aW[q][0]);
                    u64 w1 = fill_up(b[q][1] | ((w0 >> 63) & aW[q][1] & 1ull), aW[q][1]);
                    w1 = fill_down(w1, aE[q][1]);
                    w0 = fill_down(w0 | (((w1 & 1ull) << 63) & aE[q][0]), aE[q][0]);
                    ch |= (w0 != b[q][0]) | (w1 != b[q][1]);
                    b[q][0] = w0; b[q][1] = w1;
                }
#pragma unroll
                for (int w = 0; w < 2; w++) {   // columns: the row above through N arrows (downwards), the row below through S arrows
                    u64 up = __shfl_down_sync(FULL, b[0][w], 1);
                    if (lane == 31) up = 0ull;
#pragma unroll
                    for (int q = 3; q >= 0; q--) {
                        const u64 v = b[q][w] | (aN[q][w] & (q == 3 ? up : b[q + 1][w]));
                        ch |= v != b[q][w];
                        b[q][w] = v;
                    }
                    u64 dn = __shfl_up_sync(FULL, b[3][w], 1);
                    if (lane == 0) dn = 0ull;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const u64 v = b[q][w] | (aS[q][w] & (q == 0 ? dn : b[q - 1][w]));
                        ch |= v != b[q][w];
                        b[q][w] = v;
                    }
                }
                if (!__any_sync(FULL, ch)) break;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) { s_win[wid][lane * 4 + q][0] = b[q][0]; s_win[wid][lane * 4 + q][1] = b[q][1]; }
            __syncwarp();
            if (lane == 0) {
                LtWindow lt{L, s_win[wid], wx0, y0, false};
                r = lights_eval(L, cx, cy, lt);
                undecided = lt.undecided;
            }
            __syncwarp();
        }
        if (lane == 0) {
            if (!undecided) { rec[i] = r; apply_aux(L, c, r, A); }
            else { const int k = atomicAdd(n_pend2, 1); if (k < cap_pend) pend2[k] = i; else *L.err = 28; }
        }
    }
}

// 5. link tables, gather form: ONE thread per light scans the 5 x 5 cells around it (a ControlledRoad lists only
// lights within two cells, see lights_eval) in ascending cell order and takes every candidate whose record names this
// light.  No atomics, so the tables come out in a canonical order: controlled cells ascending, incoming lane cells
// grouped by their controlled cell in scan order.  FILL == false: counts -> ctrl_off / inc_off (scanned afterwards).
template <bool FILL>
__global__ void __launch_bounds__(128) light_links_kernel(LightsCtx L, const int32_t *__restrict__ n_lights, const int32_t *__restrict__ light_cell,
                                                          const int32_t *__restrict__ cr_prefix, const u64 *__restrict__ rec, int32_t *ctrl_off,
                                                          int32_t *inc_off, int32_t *__restrict__ ctrl_cell, int32_t *__restrict__ inc_cell, int cap_ctrl,
                                                          int cap_inc, int32_t *out_off, int32_t *__restrict__ out_cell, int cap_out) {
    constexpr int CH = 4;   // candidates looked up together: their prefix words, then their records, travel as independent loads
    const int n = min(*n_lights, L.cap_lights);
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < n; l += gridDim.x * blockDim.x) {
        const int a = light_cell[l], ax = a % L.W, ay = a / L.W;
        int nc = 0, ni = 0, no = 0;
        int pc = FILL ? ctrl_off[l] : 0, pi = FILL ? inc_off[l] : 0, po = (FILL && L.fwd) ? out_off[l] : 0;
        auto rec_of = [&](int c) {
            const int x = c % L.W, y = c / L.W;
            const size_t wi = (size_t)y * L.b.wp + (x >> 6);
            return cr_prefix[wi] + __popcll(L.b.cr[wi] & ((1ull << (x & 63)) - 1ull));
        };
        auto process = [&](int c, u64 q) {
            bool mine = false;
            for (int u = 0; u < rec_nacc(q); u++) mine |= rec_acc(q, u, c, L.W) == a;
            if (!mine) return;
            if (L.fwd) {   // _scan_for_traffic_flow_forward, once per (light, controlled road) like the reference
                auto emit = [&](int cell) { if (FILL) { if (po < cap_out) out_cell[po] = cell; else *L.err = 27; po++; } else no++; };
                forward_scan(L, c % L.W, c / L.W, c % L.W, c / L.W, L.D[c], (int)L.T[c], 0, emit);
            }
            if (!FILL) { nc++; ni += rec_nsc(q); return; }
            if (pc < cap_ctrl) ctrl_cell[pc] = c; else *L.err = 20;
            pc++;
            const uint32_t rd = L.D[c];
            for (int d = 0; d < dl_len(rd); d++) {
                const int kk = opp_of(dl_get(rd, d)), step = dy_of(kk) * L.W + dx_of(kk);
                for (int s2 = 1; s2 <= rec_cnt(q, d); s2++, pi++) { if (pi < cap_inc) inc_cell[pi] = c + s2 * step; else *L.err = 21; }
            }
        };
        // the candidate bits of the 5 x 5 cells around the light, row by row (= ascending cell order), five independent loads
        uint32_t all = 0;   // bit 5 * r + k = candidate at (ax - 2 + k, ay - 2 + r)
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const int y = ay - 2 + r;
            const uint32_t row = (y >= 0 && y < L.H) ? (uint32_t)extract_bits(L.b.cr + (size_t)y * L.b.wp, L.b.wp, ax - 2, 5) : 0u;
            all |= row << (5 * r);
        }
        // the first CH candidates are looked up TOGETHER (their prefix words, then their records, as independent loads: the chain
        // bit -> prefix -> record used to be walked candidate by candidate); static indices only, so everything stays in registers
        auto cell_of = [&](int bit) { return (ay - 2 + bit / 5) * L.W + ax - 2 + bit % 5; };
        int cc[CH], m = 0;
        uint32_t rest = all;
#pragma unroll
        for (int j = 0; j < CH; j++) {
            if (rest) { cc[j] = cell_of(__ffs(rest) - 1); rest &= rest - 1; m = j + 1; }
        }
        int idx[CH];
        u64 rr[CH];
#pragma unroll
        for (int j = 0; j < CH; j++) if (j < m) idx[j] = rec_of(cc[j]);
#pragma unroll
        for (int j = 0; j < CH; j++) if (j < m) rr[j] = rec[idx[j]];
#pragma unroll
        for (int j = 0; j < CH; j++) if (j < m) process(cc[j], rr[j]);
        while (rest) {   // (a light with more than CH candidates around it: the rest one by one, in order)
            const int c = cell_of(__ffs(rest) - 1);
            rest &= rest - 1;
            process(c, rec[rec_of(c)]);
        }
        if (!FILL) { ctrl_off[l] = nc; inc_off[l] = ni; if (L.fwd) out_off[l] = no; }
    }
}

// currently_scanned_road.light = traffic_light for the forward-scan cells (:1562,1566,1578); a candidate that is converted
// later gets a fresh CellAgent, so only cells that are no candidates keep the bit.  Runs before cr_apply (types still original).
__global__ void __launch_bounds__(128) fwd_mark_kernel(LightsCtx L, const int32_t *__restrict__ n_cr, const int32_t *__restrict__ cr_cell,
                                                       const u64 *__restrict__ rec, uint8_t *A) {
    const int n = *n_cr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (!rec_nacc(rec[i])) continue;
        const int c = cr_cell[i], cx = c % L.W, cy = c / L.W;
        auto emit = [&](int cell) { if (!L.bit(L.b.cr, cell % L.W, cell / L.W)) or_byte(A + cell, AUX_LIGHT); };
        forward_scan(L, cx, cy, cx, cy, L.D[c], (int)L.T[c], 0, emit);
    }
}

// per candidate: convert the cell (place_cell(..., "ControlledRoad") keeps the arrows and remembers the original type,
// :1455-1459; the aux side of it was written with the record, apply_aux)
__global__ void __launch_bounds__(256) cr_type_kernel(const int32_t *__restrict__ n_cr, const int32_t *__restrict__ cr_cell, uint8_t *T, int32_t *B) {
    const int n = *n_cr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = cr_cell[i];
        if (T[c] == T_BE) B[c] = 0;
        T[c] = T_CR;
    }
}

// Sidewalk -> TrafficLight (:1506-1509)
__global__ void __launch_bounds__(256) tl_apply_kernel(const int32_t *__restrict__ n_lights, const int32_t *__restrict__ light_cell, int cap,
                                                       uint8_t *T, uint16_t *D, uint8_t *A) {
    const int n = min(*n_lights, cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = light_cell[i];
        T[c] = T_TL; D[c] = 0; A[c] &= (AUX_RING | AUX_EVER);
    }
}

__global__ void close_offsets_kernel(const int32_t *n_lights, int32_t *ctrl_off, int32_t *inc_off, int32_t *out_off, const int32_t *totals,
                                     int cap_lights, int cap_ctrl, int cap_inc, int cap_out, int32_t *err) {
    const int n = *n_lights;
    if (n > cap_lights) { *err = 22; return; }
    ctrl_off[n] = totals[0];
    inc_off[n] = totals[1];
    if (totals[0] > cap_ctrl) *err = 20;
    if (totals[1] > cap_inc) *err = 21;
    if (out_off) { out_off[n] = totals[2]; if (totals[2] > cap_out) *err = 27; }
}

}  // namespace tsim

using namespace tsim;

struct LightsWs {   // fixed part of the workspace, shared by the three stages
    int32_t *scal;   // [0..1] pivot candidates, [4..6] link totals, [8..11] reach flags, [12] n_cr, [13] / [14] candidates left undecided by stage 1 / 2, [16..19] reach control
    Bits bp;
    ReachT rt;
    int32_t *cr_prefix, *tl_prefix, *scan_tmp, *cr_cell;
    u64 *rec;
    int32_t *pend, *pend2;   // candidates whose `leads_to` queries stage 1 (Z witnesses) / stage 2 (window closures) left undecided
    int W, H, wp, cap_cr, cap_pend;
    long long n, nw;
    size_t end;      // first free byte after the fixed part
};

static tsim_status lights_ws(const tsim_cfg *cfg, void *workspace, size_t ws_bytes, LightsWs &L) {
    L.W = cfg->width; L.H = cfg->win_rows;   // window-local rows throughout
    L.n = (long long)L.W * L.H;
    L.wp = div_up(L.W, 64);
    L.nw = (long long)L.wp * L.H;
    L.cap_cr = (int)(L.n / 4 + 1024);
    L.cap_pend = (int)(L.n / 64 + 4096);
    char *w = (char *)workspace;
    size_t o = 0;
    auto take = [&](size_t bytes) { char *q = w + o; o += (bytes + 255) & ~(size_t)255; return q; };
    L.scal = (int32_t *)take(64 * 4);
    u64 **planes[] = {&L.bp.aN, &L.bp.aE, &L.bp.aS, &L.bp.aW, &L.bp.I, &L.bp.R, &L.bp.fw, &L.bp.bw, &L.bp.cr, &L.bp.tl, &L.bp.lm};
    for (u64 **pl : planes) *pl = (u64 *)take((size_t)L.nw * 8);
    L.bp.wp = L.wp;
    L.rt.wpT = div_up(L.H, 64);
    u64 **tplanes[] = {&L.rt.aNt, &L.rt.aSt, &L.rt.fwT, &L.rt.bwT};
    for (u64 **pl : tplanes) *pl = (u64 *)take((size_t)L.rt.wpT * L.W * 8);
    L.rt.bdR = (uint8_t *)take((size_t)L.rt.wpT * L.wp);
    L.rt.bdT = (uint8_t *)take((size_t)L.rt.wpT * L.wp);
    L.rt.rd = (uint8_t *)take((size_t)L.rt.wpT);
    L.rt.cd = (uint8_t *)take((size_t)L.wp);
    L.cr_prefix = (int32_t *)take((size_t)L.nw * 4);
    L.tl_prefix = (int32_t *)take((size_t)L.nw * 4);
    L.scan_tmp = (int32_t *)take(scan_tmp_bytes(L.nw));
    L.cr_cell = (int32_t *)take((size_t)L.cap_cr * 4);
    L.rec = (u64 *)take((size_t)L.cap_cr * 8);
    L.pend = (int32_t *)take((size_t)L.cap_pend * 4);
    L.pend2 = (int32_t *)take((size_t)L.cap_pend * 4);
    L.end = o;
    if (!workspace || o > ws_bytes) { set_error("the lights pass needs %zu workspace bytes, got %zu", o, ws_bytes); return TSIM_ERR_WORKSPACE; }
    return TSIM_OK;
}

static tsim_status lights_check(const tsim_cfg *cfg) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if (cfg->traffic_light_range < 0 || cfg->traffic_light_range > MAX_TL_RANGE) {
        set_error("traffic_light_range %d outside 0..%d", cfg->traffic_light_range, MAX_TL_RANGE);
        return TSIM_ERR_UNSUPPORTED;
    }
    return TSIM_OK;
}

// stage 1: bit-planes, pivot candidates, compact candidate list
extern "C" tsim_status tsim_lights_prepare(const tsim_cfg *cfg, const tsim_planes *p, int32_t *pivot_out, int32_t *err_flag, void *workspace,
                                           size_t ws_bytes, void *stream) {
    tsim_status st = lights_check(cfg);
    if (st != TSIM_OK) return st;
    if (!p || !p->cell_type || !p->dirs || !err_flag) { set_error("tsim_lights_prepare: bad arguments"); return TSIM_ERR_CONFIG; }
    LightsWs L;
    if ((st = lights_ws(cfg, workspace, ws_bytes, L)) != TSIM_OK) return st;
    cudaStream_t cs = (cudaStream_t)stream;
    // the pivot is the first intersection at or after the middle row of the WINDOW: well inside it, so that (almost) every
    // road of the window can reach it and be reached from it without leaving the window
    const int mid_row = L.H / 2;
    TSIM_CUDA(cudaMemsetAsync(L.scal, 0, 64 * 4, cs));
    init_pivot_kernel<<<1, 1, 0, cs>>>(L.scal);
    TSIM_LAUNCH_CHECK();
    static int bits_mul = -1;   // arrow-bit extraction of lights_bits_kernel: 0 shift-and-mask per dirs word, 1 byte gather + multiplication
    if (bits_mul < 0) { const char *e = getenv("TSIM_LIGHTS_BITS"); bits_mul = (e && !strcmp(e, "shift")) ? 0 : 1; }   // measured in bench.py: lights pass 1.955 (shift) / 1.933 ms (mul)
    const dim3 bits_grid(div_up(L.wp * 4, 256), L.H < 65535 ? L.H : 65535, div_up(L.H, 65535));
    if (bits_mul) lights_bits_kernel<true><<<bits_grid, 256, 0, cs>>>(L.W, L.H, p->cell_type, p->dirs, L.bp,
                                                                                                                   (long long)mid_row * L.W, L.scal);
    else lights_bits_kernel<false><<<bits_grid, 256, 0, cs>>>(L.W, L.H, p->cell_type, p->dirs, L.bp,
                                                                                                                   (long long)mid_row * L.W, L.scal);
    TSIM_LAUNCH_CHECK();
    cr_bits_kernel<<<div_up(L.nw, 256), 256, 0, cs>>>(L.H, L.bp, L.cr_prefix);
    TSIM_LAUNCH_CHECK();
    if ((st = exclusive_scan_i32(L.cr_prefix, L.nw, L.scan_tmp, L.scal + 12, cs)) != TSIM_OK) return st;
    bit_list_kernel<<<div_up(L.nw, 256), 256, 0, cs>>>(L.W, L.H, L.wp, L.bp.cr, L.cr_prefix, L.cr_cell, L.cap_cr, err_flag, 23);
    TSIM_LAUNCH_CHECK();
    if (pivot_out) TSIM_CUDA(cudaMemcpyAsync(pivot_out, L.scal, 8, cudaMemcpyDeviceToDevice, cs));
    return TSIM_OK;
}

// stage 2a: seed the reachability planes.  pivot == NULL: this window's own candidate (single device).
extern "C" tsim_status tsim_lights_seed(const tsim_cfg *cfg, const int32_t *pivot, void *workspace, size_t ws_bytes, void *stream) {
    tsim_status st = lights_check(cfg);
    if (st != TSIM_OK) return st;
    LightsWs L;
    if ((st = lights_ws(cfg, workspace, ws_bytes, L)) != TSIM_OK) return st;
    reach_seed_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(L.W, pivot ? pivot : L.scal, pivot == nullptr, L.bp);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}

// stage 2b: closure of the planes inside this window; *changed (device, optional) is set to 1 if a bit was added
template <int NCH>
static tsim_status launch_lines(u64 *f, u64 *b, const u64 *up, const u64 *dn, int nlines, int wp, int nbits, const uint8_t *line_dirty, bool all,
                                uint8_t *bd, int bd_line_stride, int bd_word_stride, int32_t *ctl, cudaStream_t cs) {
    reach_lines_kernel<NCH><<<div_up(nlines, 8), 256, 0, cs>>>(f, b, up, dn, nlines, wp, nbits, line_dirty, all, bd, bd_line_stride, bd_word_stride, ctl);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}

static tsim_status reach_lines(u64 *f, u64 *b, const u64 *up, const u64 *dn, int nlines, int wp, int nbits, const uint8_t *line_dirty, bool all,
                               uint8_t *bd, int bd_line_stride, int bd_word_stride, int32_t *ctl, cudaStream_t cs) {
    // lines of up to 256 words live in registers (1, 2, 4 or 8 words per lane); longer ones are closed from global memory
    if (wp <= 32) return launch_lines<1>(f, b, up, dn, nlines, wp, nbits, line_dirty, all, bd, bd_line_stride, bd_word_stride, ctl, cs);
    if (wp <= 64) return launch_lines<2>(f, b, up, dn, nlines, wp, nbits, line_dirty, all, bd, bd_line_stride, bd_word_stride, ctl, cs);
    if (wp <= 128) return launch_lines<4>(f, b, up, dn, nlines, wp, nbits, line_dirty, all, bd, bd_line_stride, bd_word_stride, ctl, cs);
    if (wp <= 256) return launch_lines<8>(f, b, up, dn, nlines, wp, nbits, line_dirty, all, bd, bd_line_stride, bd_word_stride, ctl, cs);
    return launch_lines<0>(f, b, up, dn, nlines, wp, nbits, line_dirty, all, bd, bd_line_stride, bd_word_stride, ctl, cs);
}

constexpr int REACH_ALTERNATIONS = 10;   // enqueued as ordinary launches; anything beyond is finished by the cooperative kernel

// TSIM_REACH_ALTERNATIONS=<n> overrides the number of enqueued alternations (tests use 0 / 1 to exercise the fallback kernel)
static int reach_alternations() {
    const char *e = getenv("TSIM_REACH_ALTERNATIONS");
    if (!e || !*e) return REACH_ALTERNATIONS;
    const int v = atoi(e);
    return v < 0 ? 0 : (v > 64 ? 64 : v);
}

// stage 2b: closure of the planes inside this window; *changed (device, optional) is set to 1 if a bit was added.
// Each phase (row closures, re-transposition of the changed tiles, column closures, transposition back) is its own
// launch with its own grid; a phase kernel returns at once when the closure is already complete, so a fixed number of
// alternations is enqueued without any host round trip.  A city that needs more turns than that is finished by the
// persistent cooperative kernel (same result, one launch, grid-wide barriers), which otherwise exits immediately.
// `gate` (device scalar, optional): the closure is skipped -- every phase kernel returns at its first instruction -- when *gate == 0
static tsim_status lights_reach_impl(const tsim_cfg *cfg, int32_t edge_rows, int32_t *changed, int32_t *err_flag, void *workspace, size_t ws_bytes,
                                     void *stream, const int32_t *gate) {
    tsim_status st = lights_check(cfg);
    if (st != TSIM_OK) return st;
    if (!err_flag) { set_error("tsim_lights_reach: NULL err_flag"); return TSIM_ERR_CONFIG; }
    LightsWs L;
    if ((st = lights_ws(cfg, workspace, ws_bytes, L)) != TSIM_OK) return st;
    cudaStream_t cs = (cudaStream_t)stream;
    const int W = L.W, H = L.H, wp = L.wp, wpT = L.rt.wpT, nbx = wp, nby = wpT;
    const int ntiles = div_up(nbx, 8) * div_up(nby, 8);
    const size_t tsmem = (size_t)2 * 512 * 9 * 8;   // 72 KB of dynamic shared memory per transposing CTA
    TSIM_CUDA(cudaFuncSetAttribute(reach_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
    const Bits &bp = L.bp;
    const ReachT &rt = L.rt;
    int32_t *ctl = L.scal + 16;      // [0] changed, [1] done, [2] alternations
    int32_t *flags = L.scal + 8;     // cooperative kernel's own flags
    TSIM_CUDA(cudaMemsetAsync(flags, 0, 16, cs));
    reach_gate_kernel<<<1, 1, 0, cs>>>(ctl, gate);
    TSIM_LAUNCH_CHECK();
    TSIM_CUDA(cudaMemsetAsync(rt.bdR, 0, (size_t)((char *)rt.cd - (char *)rt.bdR) + wp, cs));   // the four flag arrays are contiguous
    const bool first = edge_rows < 0;
    // A gated closure (tsim_layout_lights: only cities with a query the local searches cannot settle need it) is left to the one
    // persistent kernel below: a city that does not need it pays two launches, not fifty that return at once.
    const int n_alt = gate ? 0 : reach_alternations();
    if (n_alt == 0) {
        // nothing enqueued: the cooperative kernel transposes the arrows itself
    } else if (first) {   // transposed arrow planes (kept in the workspace for resumed calls)
        reach_transpose_kernel<<<ntiles, 256, tsmem, cs>>>(bp.aN, bp.aS, H, wp, rt.aNt, rt.aSt, W, wpT, rt.bdR, nbx, nby, true, true, false, rt.cd, rt.cd, 0, ctl);
        TSIM_LAUNCH_CHECK();
        TSIM_CUDA(cudaMemsetAsync(rt.cd, 0, wp, cs));   // (the arrow transposition marked its lines; they are not reachability changes)
    } else {
        reach_mark_edges_kernel<<<div_up(nbx * nby, 256) < 1184 ? div_up(nbx * nby, 256) : 1184, 256, 0, cs>>>(H, edge_rows, nbx, nby, rt.bdR, rt.rd);
        TSIM_LAUNCH_CHECK();
    }
    for (int a = 0; a < n_alt; a++) {
        const bool all = first && a == 0;
        // rows: lines of the row-major planes; a changed word (y, w) marks block (y >> 6, w)
        if ((st = reach_lines(bp.fw, bp.bw, bp.aE, bp.aW, H, wp, W, rt.rd, all, rt.bdR, nbx, 1, ctl, cs)) != TSIM_OK) return st;
        reach_transpose_kernel<<<ntiles, 256, tsmem, cs>>>(bp.fw, bp.bw, H, wp, rt.fwT, rt.bwT, W, wpT, rt.bdR, nbx, nby, true, all, !first && a == 0, rt.cd,
                                                       rt.rd, nby, ctl);
        TSIM_LAUNCH_CHECK();
        // columns: lines of the transposed planes; a changed word (x, wy) marks block (wy, x >> 6)
        if ((st = reach_lines(rt.fwT, rt.bwT, rt.aNt, rt.aSt, W, wpT, H, rt.cd, all, rt.bdT, 1, nbx, ctl, cs)) != TSIM_OK) return st;
        reach_transpose_kernel<<<ntiles, 256, tsmem, cs>>>(rt.fwT, rt.bwT, W, wpT, bp.fw, bp.bw, H, wp, rt.bdT, nbx, nby, false, false, false, rt.rd, rt.cd,
                                                       nbx, ctl);
        TSIM_LAUNCH_CHECK();
        reach_ctl_kernel<<<1, 1, 0, cs>>>(ctl, changed);
        TSIM_LAUNCH_CHECK();
    }
    // not converged after the enqueued alternations: the cooperative kernel finishes the closure (exits at once otherwise)
    int dev = 0, sms = 0, per_sm = 0;
    TSIM_CUDA(cudaGetDevice(&dev));
    TSIM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TSIM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reach_kernel, 256, 0));
    int grid = sms * (per_sm < 1 ? 1 : per_sm);
    const int want = div_up(H > W ? H : W, 8);   // one warp per row / per column
    if (grid > want) grid = want;
    int full = -1;
    const int32_t *skip = ctl + 1;
    LightsWs *Lp = &L;
    void *args[] = {&Lp->W, &Lp->H, &Lp->bp, &Lp->rt, &full, &skip, &flags, &changed, &err_flag};
    TSIM_COOP_LAUNCH(reach_kernel, dim3(grid), dim3(256), args, cs);
    return TSIM_OK;
}

extern "C" tsim_status tsim_lights_reach(const tsim_cfg *cfg, int32_t edge_rows, int32_t *changed, int32_t *err_flag, void *workspace, size_t ws_bytes,
                                         void *stream) {
    return lights_reach_impl(cfg, edge_rows, changed, err_flag, workspace, ws_bytes, stream, nullptr);
}

// byte offsets of the two reachability planes inside the workspace ([win_rows][words_per_row] uint64, bit x & 63 of
// word x >> 6 = cell x), so that shards can exchange (OR) their halo rows between tsim_lights_reach calls
extern "C" tsim_status tsim_lights_reach_planes(const tsim_cfg *cfg, size_t ws_bytes, size_t *fw_off, size_t *bw_off, int32_t *words_per_row) {
    tsim_status st = lights_check(cfg);
    if (st != TSIM_OK) return st;
    if (!fw_off || !bw_off || !words_per_row) { set_error("tsim_lights_reach_planes: NULL output"); return TSIM_ERR_CONFIG; }
    LightsWs L;
    char dummy;
    if ((st = lights_ws(cfg, &dummy, ws_bytes, L)) != TSIM_OK) return st;
    *fw_off = (size_t)((char *)L.bp.fw - &dummy); *bw_off = (size_t)((char *)L.bp.bw - &dummy); *words_per_row = L.wp;
    return TSIM_OK;
}

// stage 3: evaluate the candidates, number the lights, build the link tables, convert the cells
// TSIM_LIGHTS_STAGES=<mask> switches search stages off (tests: every stage must give the same city): 1 no Z witnesses, 2 no window closures
static int lights_stages_off() {
    const char *e = getenv("TSIM_LIGHTS_STAGES");
    return (e && *e) ? (atoi(e) & 3) : 0;
}

// lazy == false: the caller has closed the reachability planes (tsim_lights_seed / tsim_lights_reach) and every `leads_to` is read
// from them.  lazy == true (tsim_layout_lights): the queries are settled by the staged searches first and the planes are only
// closed -- inside this call, around this window's own pivot -- when a query is left over.
// what: 1 = evaluate the candidates (records, light bit-plane, aux updates), 2 = number the lights, build the link tables, convert
// the cells; 3 = both.
static tsim_status lights_finish_impl(const tsim_cfg *cfg, const tsim_planes *p, const tsim_light_links *lk, int32_t *err_flag, void *workspace,
                                      size_t ws_bytes, void *stream, bool lazy, int what = 3) {
    tsim_status st = lights_check(cfg);
    if (st != TSIM_OK) return st;
    if (!p || !p->cell_type || !p->dirs || !p->aux || !p->block_id || !lk || !err_flag || !lk->n_lights || !lk->light_cell || !lk->ctrl_off ||
        !lk->ctrl_cell || !lk->inc_off || !lk->inc_cell) {
        set_error("tsim_lights_finish: bad arguments");
        return TSIM_ERR_CONFIG;
    }
    LightsWs ws;
    if ((st = lights_ws(cfg, workspace, ws_bytes, ws)) != TSIM_OK) return st;
    char *w = (char *)workspace;
    size_t o = ws.end;
    auto take = [&](size_t bytes) { char *q = w + o; o += (bytes + 255) & ~(size_t)255; return q; };
    int32_t *scan_tmp = (int32_t *)take(scan_tmp_bytes(ws.nw > lk->cap_lights ? ws.nw : lk->cap_lights));
    if (o > ws_bytes) { set_error("tsim_lights_finish needs %zu workspace bytes, got %zu", o, ws_bytes); return TSIM_ERR_WORKSPACE; }
    cudaStream_t cs = (cudaStream_t)stream;
    const int W = ws.W, H = ws.H, wp = ws.wp, cap_cr = ws.cap_cr;
    const long long nw = ws.nw;
    const Bits &bp = ws.bp;
    int32_t *scal = ws.scal, *n_cr = ws.scal + 12, *cr_cell = ws.cr_cell, *tl_prefix = ws.tl_prefix;
    u64 *rec = ws.rec;
    // 4. evaluate every candidate once (launch covers the capacity; threads beyond *n_cr exit)
    const int list_grid = div_up(cap_cr, 128) < 148 * 16 ? div_up(cap_cr, 128) : 148 * 16;   // grid-stride over the compact list
    const int fwd = cfg->forward_traffic_light_range ? 1 : 0;
    if (fwd && (!lk->out_off || !lk->out_cell || lk->cap_out < 1)) { set_error("forward_traffic_light_range needs the outgoing link table"); return TSIM_ERR_CONFIG; }
    LightsCtx L{W, H, cfg->traffic_light_range, lk->cap_lights, fwd, cfg->forward_intersections_mode, lights_stages_off(), cfg->win_y0 > 0 ? cfg->win_halo : 0,
                cfg->win_y0 + cfg->win_rows < cfg->height ? cfg->win_halo : 0, p->cell_type, p->dirs, bp, err_flag};
    if (!(what & 1)) {
        // records are there already (tsim_lights_eval)
    } else if (!lazy) {
        lights_eval_kernel<1><<<list_grid, 128, 0, cs>>>(L, n_cr, cr_cell, rec, p->aux, nullptr, nullptr, nullptr, 0);
        TSIM_LAUNCH_CHECK();
    } else {
        int32_t *n_pend = scal + 13, *n_pend2 = scal + 14;
        lights_eval_kernel<0><<<list_grid, 128, 0, cs>>>(L, n_cr, cr_cell, rec, p->aux, nullptr, n_pend, ws.pend, ws.cap_pend);
        TSIM_LAUNCH_CHECK();
        lights_window_kernel<<<148 * 4, 128, 0, cs>>>(L, n_pend, ws.pend, cr_cell, rec, p->aux, n_pend2, ws.pend2, ws.cap_pend);
        TSIM_LAUNCH_CHECK();
        // whatever is still undecided (and every "False") needs the planes: closed now, or skipped when the list is empty
        if ((st = tsim_lights_seed(cfg, nullptr, workspace, ws_bytes, stream)) != TSIM_OK) return st;
        if ((st = lights_reach_impl(cfg, -1, nullptr, err_flag, workspace, ws_bytes, stream, n_pend2)) != TSIM_OK) return st;
        const int pgrid = div_up(ws.cap_pend, 128) < 148 * 4 ? div_up(ws.cap_pend, 128) : 148 * 4;
        lights_eval_kernel<2><<<pgrid, 128, 0, cs>>>(L, n_pend2, cr_cell, rec, p->aux, ws.pend2, nullptr, nullptr, ws.cap_pend);
        TSIM_LAUNCH_CHECK();
    }
    if (what & 1) {   // every record is final: the marks of their reverse marches go into the aux plane
        aux_light_merge_kernel<<<div_up(nw, 256), 256, 0, cs>>>(W, nw, wp, bp.lm, bp.cr, p->aux);
        TSIM_LAUNCH_CHECK();
    }
    if (!(what & 2)) return TSIM_OK;
    // 5. lights in ascending cell order
    bit_count_kernel<<<div_up(nw, 256), 256, 0, cs>>>(nw, bp.tl, tl_prefix);
    TSIM_LAUNCH_CHECK();
    if ((st = exclusive_scan_i32(tl_prefix, nw, scan_tmp, lk->n_lights, cs)) != TSIM_OK) return st;
    bit_list_kernel<<<div_up(nw, 256), 256, 0, cs>>>(W, H, wp, bp.tl, tl_prefix, lk->light_cell, lk->cap_lights, err_flag, 22);
    TSIM_LAUNCH_CHECK();
    // count -> offsets -> fill (gather form, one thread per light)
    const int lgrid = div_up(lk->cap_lights, 128) < 148 * 16 ? div_up(lk->cap_lights, 128) : 148 * 16;
    light_links_kernel<false><<<lgrid, 128, 0, cs>>>(L, lk->n_lights, lk->light_cell, ws.cr_prefix, rec, lk->ctrl_off, lk->inc_off, nullptr, nullptr, 0, 0,
                                                     lk->out_off, nullptr, 0);
    TSIM_LAUNCH_CHECK();
    if (fwd && (st = exclusive_scan_i32(lk->out_off, lk->cap_lights, scan_tmp, scal + 6, cs, lk->n_lights)) != TSIM_OK) return st;
    if ((st = exclusive_scan_i32(lk->ctrl_off, lk->cap_lights, scan_tmp, scal + 4, cs, lk->n_lights)) != TSIM_OK) return st;
    if ((st = exclusive_scan_i32(lk->inc_off, lk->cap_lights, scan_tmp, scal + 5, cs, lk->n_lights)) != TSIM_OK) return st;
    close_offsets_kernel<<<1, 1, 0, cs>>>(lk->n_lights, lk->ctrl_off, lk->inc_off, fwd ? lk->out_off : nullptr, scal + 4, lk->cap_lights, lk->cap_ctrl,
                                          lk->cap_inc, lk->cap_out, err_flag);
    TSIM_LAUNCH_CHECK();
    light_links_kernel<true><<<lgrid, 128, 0, cs>>>(L, lk->n_lights, lk->light_cell, ws.cr_prefix, rec, lk->ctrl_off, lk->inc_off, lk->ctrl_cell,
                                                    lk->inc_cell, lk->cap_ctrl, lk->cap_inc, lk->out_off, lk->out_cell, lk->cap_out);
    TSIM_LAUNCH_CHECK();
    if (fwd) {
        fwd_mark_kernel<<<list_grid, 128, 0, cs>>>(L, n_cr, cr_cell, rec, p->aux);
        TSIM_LAUNCH_CHECK();
    }
    cr_type_kernel<<<(div_up(cap_cr, 256) < 148 * 8 ? div_up(cap_cr, 256) : 148 * 8), 256, 0, cs>>>(n_cr, cr_cell, p->cell_type, p->block_id);
    TSIM_LAUNCH_CHECK();
    tl_apply_kernel<<<(div_up(lk->cap_lights, 256) < 148 * 8 ? div_up(lk->cap_lights, 256) : 148 * 8), 256, 0, cs>>>(lk->n_lights, lk->light_cell, lk->cap_lights, p->cell_type, p->dirs, p->aux);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}

extern "C" tsim_status tsim_lights_finish(const tsim_cfg *cfg, const tsim_planes *p, const tsim_light_links *lk, int32_t *err_flag, void *workspace,
                                          size_t ws_bytes, void *stream) {
    return lights_finish_impl(cfg, p, lk, err_flag, workspace, ws_bytes, stream, false);
}

extern "C" tsim_status tsim_lights_eval(const tsim_cfg *cfg, const tsim_planes *p, const tsim_light_links *lk, int32_t *err_flag, void *workspace,
                                        size_t ws_bytes, void *stream) {
    return lights_finish_impl(cfg, p, lk, err_flag, workspace, ws_bytes, stream, true, 1);
}

extern "C" tsim_status tsim_lights_links(const tsim_cfg *cfg, const tsim_planes *p, const tsim_light_links *lk, int32_t *err_flag, void *workspace,
                                         size_t ws_bytes, void *stream) {
    return lights_finish_impl(cfg, p, lk, err_flag, workspace, ws_bytes, stream, true, 2);
}

extern "C" tsim_status tsim_layout_lights(const tsim_cfg *cfg, const tsim_planes *p, const tsim_light_links *lk, int32_t *err_flag,
                                          void *workspace, size_t ws_bytes, void *stream) {
    tsim_status st;
    if ((st = tsim_lights_prepare(cfg, p, nullptr, err_flag, workspace, ws_bytes, stream)) != TSIM_OK) return st;
    return lights_finish_impl(cfg, p, lk, err_flag, workspace, ws_bytes, stream, true);
}
