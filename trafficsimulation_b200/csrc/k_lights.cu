// k_lights.cu -- _add_traffic_lights (city_model.py:1422-1499), _assign_traffic_light (:1501-1520),
// _scan_for_traffic_flow_reverse (:1528-1548) and CellAgent.leads_to (cell.py:201-227).
//
// The reference sweeps the grid COLUMN-major and mutates it while it goes; three things depend on
// that order and are reproduced with explicit predicates instead of a serial sweep:
//   * a neighbour already converted to ControlledRoad (it precedes the cell in column-major order and
//     itself points into an Intersection) no longer has its original type -- `type_at_time`;
//   * Sidewalk -> TrafficLight conversions never influence a test (both are accepted), so they commute;
//   * `nb.light` of a cell that is converted later is reset by the conversion (a fresh CellAgent).
// `leads_to` is a BFS over the whole directed arrow graph.  It is answered in O(1) for almost every
// query from two reachability planes computed once: FW = reachable from a pivot intersection, BW = can
// reach the pivot; BW(a) && FW(b) implies a -> pivot -> b.  The remaining queries have a source that
// cannot reach the pivot (a sink region, e.g. highway exit lanes) or a target the pivot cannot reach (a
// source region); their forward / backward closures are tiny and are searched exactly, with a bounded
// visited list that raises TSIM_ERR_CAPACITY rather than guessing.
//
// Kernel sequence (all on the caller's stream):
//   pivot -> reach (cooperative, tile wavefront) -> mark ControlledRoad candidates -> mark lights ->
//   compact lights (scan) -> count links -> scan -> fill links + has-light bits -> apply types.
#include <cooperative_groups.h>
#include "scan.cuh"

namespace cg = cooperative_groups;

namespace tsim {

constexpr int F_CR = 1, F_TL = 2;   // flag plane bits
constexpr int BFS_CAP = 160;        // visited cells of an exact fallback search

struct LightsCtx {
    int W, H, tl_range, cap_lights;
    const uint8_t *T; const uint16_t *D;
    const unsigned long long *fw, *bw; int wp;   // reach bit-planes: FW = reachable from the pivot, BW = reaches it
    const uint8_t *F;   // F_CR / F_TL
    int32_t *err;
    __device__ __forceinline__ bool has(int x, int y) const { return x >= 0 && x < W && y >= 0 && y < H; }
    __device__ __forceinline__ int at(int x, int y) const { return y * W + x; }
};

// ---------------------------------------------------------------- pivot
__global__ void __launch_bounds__(256) pivot_kernel(long long n, long long mid, const uint8_t *__restrict__ T, const uint16_t *__restrict__ D, int32_t *piv /* [2] */) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (T[i] == T_INTER && D[i] != 0) {
        if (i >= mid) atomicMin(piv + 0, (int)i);
        atomicMin(piv + 1, (int)i);
    }
}

__global__ void init_pivot_kernel(int32_t *piv) { piv[0] = 0x7fffffff; piv[1] = 0x7fffffff; }

// ---------------------------------------------------------------- arrow bit-planes
// One bit per cell and direction: AR[d][y][wx] bit (x & 63) = cell (x,y) has arrow d.  64 cells per
// word turn a lane march of 64 steps into a 6-step Kogge-Stone fill, and 32 rows per warp turn a
// column march into a 5-step shuffle scan.
struct BitPlanes {
    unsigned long long *ar[4];   // arrows N,E,S,W
    unsigned long long *fw, *bw; // reachable from the pivot / reaches the pivot
    int wp;                      // words per row
};

__global__ void __launch_bounds__(256) pack_arrows_kernel(int W, int H, const uint16_t *__restrict__ D, BitPlanes bp) {
    // one warp per 64 cells of a row: two ballots per direction
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (long long)bp.wp * H) return;
    const int y = (int)(gw / bp.wp), wx = (int)(gw % bp.wp), x0 = wx * 64;
    const uint32_t d0 = (x0 + lane < W) ? D[(size_t)y * W + x0 + lane] : 0u;
    const uint32_t d1 = (x0 + 32 + lane < W) ? D[(size_t)y * W + x0 + 32 + lane] : 0u;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const unsigned long long lo = __ballot_sync(0xffffffffu, (d0 >> k) & 1u), hi = __ballot_sync(0xffffffffu, (d1 >> k) & 1u);
        if (lane == 0) bp.ar[k][(size_t)y * bp.wp + wx] = lo | (hi << 32);
    }
    if (lane == 0) { bp.fw[(size_t)y * bp.wp + wx] = 0ull; bp.bw[(size_t)y * bp.wp + wx] = 0ull; }
}

constexpr int RTH = 32;   // rows per warp tile (tile = 64 x 32 cells)

__global__ void reach_seed_kernel(int W, int32_t *piv, BitPlanes bp, uint8_t *dirty) {
    int p = piv[0] != 0x7fffffff ? piv[0] : piv[1];
    if (p == 0x7fffffff) { piv[0] = -1; return; }   // no intersection at all: every query goes to the exact search
    piv[0] = p;
    const int x = p % W, y = p / W;
    bp.fw[(size_t)y * bp.wp + (x >> 6)] = 1ull << (x & 63);
    bp.bw[(size_t)y * bp.wp + (x >> 6)] = 1ull << (x & 63);
    dirty[(y / RTH) * bp.wp + (x >> 6)] = 1;
}

// occluded fills inside a 64-bit row word (Kogge-Stone): `p` = cells that may be entered from the
// lower (up-fill) / higher (down-fill) neighbour
__device__ __forceinline__ unsigned long long fill_up(unsigned long long f, unsigned long long p) {
    f |= p & (f << 1);  p &= p << 1;
    f |= p & (f << 2);  p &= p << 2;
    f |= p & (f << 4);  p &= p << 4;
    f |= p & (f << 8);  p &= p << 8;
    f |= p & (f << 16); p &= p << 16;
    f |= p & (f << 32);
    return f;
}
__device__ __forceinline__ unsigned long long fill_down(unsigned long long f, unsigned long long p) {
    f |= p & (f >> 1);  p &= p >> 1;
    f |= p & (f >> 2);  p &= p >> 2;
    f |= p & (f >> 4);  p &= p >> 4;
    f |= p & (f >> 8);  p &= p >> 8;
    f |= p & (f >> 16); p &= p >> 16;
    f |= p & (f >> 32);
    return f;
}
// the same fills across the 32 rows of a warp tile (lane = row): row y may be entered from row y-1
// (lane_up) / y+1 (lane_down) at the columns of `p`
__device__ __forceinline__ unsigned long long lanes_up(unsigned long long f, unsigned long long p, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long ff = __shfl_up_sync(0xffffffffu, f, d), pp = __shfl_up_sync(0xffffffffu, p, d);
        if (lane < d) { ff = 0ull; pp = 0ull; }
        f |= p & ff;
        p &= pp;
    }
    return f;
}
__device__ __forceinline__ unsigned long long lanes_down(unsigned long long f, unsigned long long p, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long ff = __shfl_down_sync(0xffffffffu, f, d), pp = __shfl_down_sync(0xffffffffu, p, d);
        if (lane + d > 31) { ff = 0ull; pp = 0ull; }
        f |= p & ff;
        p &= pp;
    }
    return f;
}

// ---------------------------------------------------------------- reach (FW/BW from the pivot)
// Persistent cooperative kernel, one WARP per 64x32 tile.  A tile is (re)processed only when a
// neighbouring tile changed one of its border cells; inside a tile the closure is a handful of
// bit-parallel fills instead of a cell-by-cell march.
__global__ void __launch_bounds__(256) reach_kernel(int W, int H, BitPlanes bp, uint8_t *dirty0, uint8_t *dirty1, int32_t *counter /* [0..1] wave counters, [2] waves */) {
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int wp = bp.wp, tilesy = (H + RTH - 1) / RTH, ntiles = wp * tilesy;
    int wave = 0;
    for (;; wave++) {
        uint8_t *cur = (wave & 1) ? dirty1 : dirty0, *nxt = (wave & 1) ? dirty0 : dirty1;
        int32_t *cnt = counter + (wave & 1);
        for (int tile = warp; tile < ntiles; tile += nwarps) {
            int flag = 0;
            if (lane == 0) { flag = __ldcg(cur + tile); if (flag) cur[tile] = 0; }
            flag = __shfl_sync(0xffffffffu, flag, 0);
            if (!flag) continue;
            const int wx = tile % wp, ty = tile / wp, y = ty * RTH + lane;
            const bool in = y < H;
            const size_t o = (size_t)y * wp + wx;
            auto ld = [&](const unsigned long long *pl, bool ok, size_t off) { return ok ? __ldcg(pl + off) : 0ull; };
            const unsigned long long aN = ld(bp.ar[DN], in, o), aE = ld(bp.ar[DE], in, o), aS = ld(bp.ar[DS], in, o), aW = ld(bp.ar[DW], in, o);
            unsigned long long f = ld(bp.fw, in, o), b = ld(bp.bw, in, o);
            const unsigned long long f0 = f, b0 = b;
            // halos: west / east words of my row, rows just below / above the tile
            const bool hw = in && wx > 0, he = in && wx + 1 < wp;
            const unsigned long long fW = ld(bp.fw, hw, o - 1), bW = ld(bp.bw, hw, o - 1), eW = ld(bp.ar[DE], hw, o - 1);
            const unsigned long long fE = ld(bp.fw, he, o + 1), bE = ld(bp.bw, he, o + 1), wE = ld(bp.ar[DW], he, o + 1);
            const int yb = ty * RTH - 1, yt = ty * RTH + RTH;
            const bool hb = yb >= 0, ht = yt < H;
            const size_t ob = (size_t)yb * wp + wx, ot = (size_t)yt * wp + wx;
            // (loaded by every lane: same address, one transaction)
            const unsigned long long fB = ld(bp.fw, hb, ob), bB = ld(bp.bw, hb, ob), nB = ld(bp.ar[DN], hb, ob);
            const unsigned long long fT = ld(bp.fw, ht, ot), bT = ld(bp.bw, ht, ot), sT = ld(bp.ar[DS], ht, ot);
            // inflow that never changes while the tile iterates
            const unsigned long long f_in = (((fW & eW) >> 63) & 1ull) | ((((fE & wE) & 1ull)) << 63) |
                                            (lane == 0 ? (fB & nB) : 0ull) | (lane == 31 ? (fT & sT) : 0ull);
            const unsigned long long b_in = (aW & ((bW >> 63) & 1ull)) | (aE & ((bE & 1ull) << 63)) |
                                            (lane == 0 ? (aS & bB) : 0ull) | (lane == 31 ? (aN & bT) : 0ull);
            f |= f_in; b |= b_in;
            for (int it = 0; it < 4096; it++) {
                const unsigned long long pf = f, pb = b;
                // FW moves WITH the arrows
                f = fill_up(f, aE << 1);                 // east:  x entered from x-1 if (x-1) has E
                f = fill_down(f, aW >> 1);               // west
                {   // north: row y entered from y-1 where row y-1 has N
                    unsigned long long pn = __shfl_up_sync(0xffffffffu, aN, 1); if (lane == 0) pn = 0ull;
                    f = lanes_up(f, pn, lane);
                    unsigned long long ps = __shfl_down_sync(0xffffffffu, aS, 1); if (lane == 31) ps = 0ull;
                    f = lanes_down(f, ps, lane);
                }
                // BW moves AGAINST the arrows: a cell with arrow d inherits from its d-neighbour
                b = fill_down(b, aE);                    // cell x takes from x+1 if x has E
                b = fill_up(b, aW);                      // cell x takes from x-1 if x has W
                b = lanes_down(b, aN, lane);             // row y takes from y+1 where row y has N
                b = lanes_up(b, aS, lane);               // row y takes from y-1 where row y has S
                if (!__any_sync(0xffffffffu, f != pf || b != pb)) break;
            }
            const bool chf = f != f0, chb = b != b0;
            if (in && chf) bp.fw[o] = f;
            if (in && chb) bp.bw[o] = b;
            // wake the neighbours whose halo I changed
            const unsigned long long ch = (f ^ f0) | (b ^ b0);
            const bool wW = __any_sync(0xffffffffu, in && (ch & 1ull)) && wx > 0;
            const bool wEe = __any_sync(0xffffffffu, in && (ch >> 63)) && wx + 1 < wp;
            const bool wB = __shfl_sync(0xffffffffu, ch != 0ull, 0) && ty > 0;
            const int last = min(RTH, H - ty * RTH) - 1;
            const bool wT = __shfl_sync(0xffffffffu, ch != 0ull, last) && ty + 1 < tilesy;
            if (lane == 0) {
                int woke = 0;
                if (wW) { nxt[tile - 1] = 1; woke = 1; }
                if (wEe) { nxt[tile + 1] = 1; woke = 1; }
                if (wB) { nxt[tile - wp] = 1; woke = 1; }
                if (wT) { nxt[tile + wp] = 1; woke = 1; }
                if (woke) atomicAdd(cnt, 1);
            }
        }
        __threadfence();
        grid.sync();
        const int any = *((volatile int32_t *)cnt);
        if (blockIdx.x == 0 && threadIdx.x == 0) { counter[(wave + 1) & 1] = 0; counter[2] = wave + 1; }
        if (!any) break;
        grid.sync();
    }
}

// ---------------------------------------------------------------- exact fallback searches
__device__ bool bfs_forward(const LightsCtx &L, int from, int to) {
    int q[BFS_CAP];
    int head = 0, tail = 0;
    q[tail++] = from;
    while (head < tail) {
        const int c = q[head++];
        if (c == to) return true;
        const uint32_t d = L.D[c];
        const int x = c % L.W, y = c / L.W;
        for (int i = 0; i < dl_len(d); i++) {
            const int k = dl_get(d, i), nx = x + dx_of(k), ny = y + dy_of(k);
            if (!L.has(nx, ny)) continue;
            const int j = L.at(nx, ny);
            bool seen = false;
            for (int u = 0; u < tail; u++) seen |= (q[u] == j);
            if (seen) continue;
            if (j != to && L.D[j] == 0) continue;   // arrow-less cells are dead ends of the BFS
            if (tail == BFS_CAP) { *L.err = 10; return false; }
            q[tail++] = j;
        }
    }
    return false;
}

__device__ bool bfs_backward(const LightsCtx &L, int from, int to) {   // does `from` reach `to`? search predecessors of `to`
    int q[BFS_CAP];
    int head = 0, tail = 0;
    q[tail++] = to;
    while (head < tail) {
        const int c = q[head++];
        if (c == from) return true;
        const int x = c % L.W, y = c / L.W;
        for (int k = 0; k < 4; k++) {   // predecessor p = c - dir(k) with arrow k
            const int nx = x - dx_of(k), ny = y - dy_of(k);
            if (!L.has(nx, ny)) continue;
            const int j = L.at(nx, ny);
            if (!dl_has(L.D[j], k)) continue;
            bool seen = false;
            for (int u = 0; u < tail; u++) seen |= (q[u] == j);
            if (seen) continue;
            if (tail == BFS_CAP) { *L.err = 11; return false; }
            q[tail++] = j;
        }
    }
    return false;
}

// cell.py:201-227 `a.leads_to(b)`
__device__ __forceinline__ bool leads_to(const LightsCtx &L, int a, int b) {
    if (a == b) return true;
    const int ax = a % L.W, ay = a / L.W, bx = b % L.W, by = b / L.W;
    const bool a_bw = (L.bw[(size_t)ay * L.wp + (ax >> 6)] >> (ax & 63)) & 1ull;
    const bool b_fw = (L.fw[(size_t)by * L.wp + (bx >> 6)] >> (bx & 63)) & 1ull;
    if (a_bw && b_fw) return true;
    if (!a_bw) return bfs_forward(L, a, b);
    return bfs_backward(L, a, b);
}

// ---------------------------------------------------------------- per ControlledRoad evaluation
__device__ __forceinline__ bool before_cm(int sx, int sy, int cx, int cy) { return sx < cx || (sx == cx && sy < cy); }   // column-major order

// cell type as the reference sees it when it visits (cx,cy)
__device__ __forceinline__ int type_at_time(const LightsCtx &L, int sx, int sy, int cx, int cy) {
    const int s = L.at(sx, sy);
    if ((L.F[s] & F_CR) && before_cm(sx, sy, cx, cy)) return T_CR;
    return L.T[s];
}

// is (x,y) converted to ControlledRoad?  (:1439-1452)
__device__ __forceinline__ bool is_controlled(int W, int H, const uint8_t *T, const uint16_t *D, int x, int y) {
    const int i = y * W + x;
    if (!in_set(SET_ROAD_NO_INT, T[i])) return false;
    const uint32_t d = D[i];
    for (int q = 0; q < dl_len(d); q++) {
        const int k = dl_get(d, q), nx = x + dx_of(k), ny = y + dy_of(k);
        if (nx >= 0 && nx < W && ny >= 0 && ny < H && T[ny * W + nx] == T_INTER) return true;
    }
    return false;
}

struct Eval {
    int acc[8], nacc;    // Sidewalk cells that become / are this road's lights
    int sc[12], nsc;     // reverse-scan cells (assigned incoming lane cells), in scan order
};

__device__ void lights_eval(const LightsCtx &L, int cx, int cy, Eval &e) {
    const int c = L.at(cx, cy);
    const int t = L.T[c];
    const uint32_t rd = L.D[c];
    e.nacc = 0; e.nsc = 0;
    int vx[4], vy[4], nv = 0;
    for (int r = 0; r < dl_len(rd); r++) {   // cells to the right of every arrow, de-duplicated (:1465-1474)
        const int k = right_of(dl_get(rd, r)), bx = cx + dx_of(k), by = cy + dy_of(k);
        bool dup = false;
        for (int u = 0; u < nv; u++) dup |= (vx[u] == bx && vy[u] == by);
        if (!dup) { vx[nv] = bx; vy[nv] = by; nv++; }
    }
    for (int u = 0; u < nv; u++) {
        if (!L.has(vx[u], vy[u])) continue;
        const int st = type_at_time(L, vx[u], vy[u], cx, cy);
        if (st == T_CR || st == t) {
            if (!(L.D[L.at(vx[u], vy[u])] & rd & 0xf)) continue;   // shares no arrow (:1483)
            const int fx = 2 * vx[u] - cx, fy = 2 * vy[u] - cy;
            if (L.has(fx, fy) && L.T[L.at(fx, fy)] == T_SIDEWALK) e.acc[e.nacc++] = L.at(fx, fy);
        }
        if (L.T[L.at(vx[u], vy[u])] == T_SIDEWALK) e.acc[e.nacc++] = L.at(vx[u], vy[u]);
    }
    if (e.nacc == 0) return;
    int depth = 0;   // budget shared by all directions (:1528-1548)
    for (int i = 0; i < dl_len(rd); i++) {
        const int k = opp_of(dl_get(rd, i));
        int bx = cx + dx_of(k), by = cy + dy_of(k);
        while (depth <= L.tl_range) {
            if (!L.has(bx, by)) break;
            if (type_at_time(L, bx, by, cx, cy) != t) break;
            const int nb = L.at(bx, by);
            if (!leads_to(L, nb, c)) break;
            e.sc[e.nsc++] = nb;
            bx += dx_of(k); by += dy_of(k); depth++;
        }
    }
}

__device__ __forceinline__ void or_byte(uint8_t *p, uint32_t bits) {
    uint32_t *w = (uint32_t *)((uintptr_t)p & ~(uintptr_t)3);
    atomicOr(w, bits << (8 * ((uintptr_t)p & 3)));
}

__global__ void __launch_bounds__(256) mark_cr_kernel(int W, int H, const uint8_t *__restrict__ T, const uint16_t *__restrict__ D, uint8_t *F) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)W * H) return;
    uint8_t f = 0;
    if (in_set(SET_ROAD_NO_INT, T[i]) && is_controlled(W, H, T, D, (int)(i % W), (int)(i / W))) f = F_CR;
    F[i] = f;
}

// mode 0: mark lights; 1: count links per light; 2: fill links and set has-light bits
template <int MODE>
__global__ void __launch_bounds__(128) lights_pass_kernel(LightsCtx L, uint8_t *F, uint8_t *A, const int32_t *__restrict__ lid,
                                                          int32_t *ctrl_cnt, int32_t *inc_cnt, const int32_t *__restrict__ ctrl_off,
                                                          const int32_t *__restrict__ inc_off, int32_t *ctrl_cell, int32_t *inc_cell,
                                                          int cap_ctrl, int cap_inc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)L.W * L.H) return;
    if (!(L.F[i] & F_CR)) return;
    Eval e;
    lights_eval(L, (int)(i % L.W), (int)(i / L.W), e);
    if (e.nacc == 0) return;
    if (MODE == 0) {
        for (int u = 0; u < e.nacc; u++) or_byte(F + e.acc[u], F_TL);
    } else if (MODE == 1) {
        for (int u = 0; u < e.nacc; u++) {
            const int l = lid[e.acc[u]];
            if (l >= L.cap_lights) continue;
            atomicAdd(ctrl_cnt + l, 1);
            if (e.nsc) atomicAdd(inc_cnt + l, e.nsc);
        }
    } else {
        for (int u = 0; u < e.nacc; u++) {
            const int l = lid[e.acc[u]];
            if (l >= L.cap_lights) continue;
            const int pc = ctrl_off[l] + atomicAdd(ctrl_cnt + l, 1);
            if (pc < cap_ctrl) ctrl_cell[pc] = (int32_t)i; else *L.err = 20;
            if (e.nsc) {
                const int pi = inc_off[l] + atomicAdd(inc_cnt + l, e.nsc);
                for (int s = 0; s < e.nsc; s++) { if (pi + s < cap_inc) inc_cell[pi + s] = e.sc[s]; else *L.err = 21; }
            }
        }
        or_byte(A + i, AUX_LIGHT);                                   // controlled_road.light = tl (:1517)
        for (int s = 0; s < e.nsc; s++)                              // nb.light = tl (:1542), lost again if nb is converted later
            if (!(L.F[e.sc[s]] & F_CR)) or_byte(A + e.sc[s], AUX_LIGHT);
    }
}

// lights in ascending cell order: count flags per tile / rank
__global__ void __launch_bounds__(256) flag_count_kernel(long long n, const uint8_t *__restrict__ F, int bit, int32_t *tile_count) {
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * SCAN_TILE;
    int c = 0;
#pragma unroll
    for (int k = 0; k < SCAN_TILE / 256; k++) {
        const long long i = base + k * 256 + threadIdx.x;
        if (i < n && (F[i] & bit)) c++;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_count[blockIdx.x] = s_cnt;
}

__global__ void __launch_bounds__(256) flag_rank_kernel(long long n, const uint8_t *__restrict__ F, int bit, const int32_t *__restrict__ tile_off,
                                                        int32_t *__restrict__ lid, int32_t *__restrict__ cells, int cap, int32_t *err) {
    __shared__ int s_warp[8];
    const long long base = (long long)blockIdx.x * SCAN_TILE;
    int running = tile_off[blockIdx.x];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = 0; k < SCAN_TILE / 256; k++) {
        const long long i = base + k * 256 + threadIdx.x;
        const bool f = i < n && (F[i] & bit);
        const uint32_t m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[w] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { const int c = s_warp[q]; if (q < w) before += c; total += c; }
        if (f) {
            const int id = running + before + __popc(m & ((1u << lane) - 1u));
            lid[i] = id;
            if (id < cap) cells[id] = (int32_t)i; else *err = 22;
        }
        running += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) lights_apply_kernel(long long n, uint8_t *T, uint16_t *D, uint8_t *A, int32_t *B, const uint8_t *__restrict__ F) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t f = F[i];
    if (f & F_CR) {   // place_cell(..., "ControlledRoad") keeps the arrows, remembers the original type (:1455-1459)
        const int t = T[i];
        T[i] = T_CR;
        A[i] = (uint8_t)((A[i] & (AUX_RING | AUX_EVER | AUX_LIGHT)) | t);
        if (t == T_BE) B[i] = 0;
    } else if (f & F_TL) {   // :1506-1509
        T[i] = T_TL; D[i] = 0; A[i] &= (AUX_RING | AUX_EVER);
    }
}

__global__ void close_offsets_kernel(const int32_t *n_lights, int32_t *ctrl_off, int32_t *inc_off, const int32_t *totals, int cap_lights,
                                     int cap_ctrl, int cap_inc, int32_t *err) {
    const int n = *n_lights;
    if (n > cap_lights) { *err = 22; return; }
    ctrl_off[n] = totals[0];
    inc_off[n] = totals[1];
    if (totals[0] > cap_ctrl) *err = 20;
    if (totals[1] > cap_inc) *err = 21;
}

}  // namespace tsim

using namespace tsim;

extern "C" tsim_status tsim_layout_lights(const tsim_cfg *cfg, const tsim_planes *p, const tsim_light_links *lk, int32_t *err_flag,
                                          void *workspace, size_t ws_bytes, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if (!p || !p->cell_type || !p->dirs || !p->aux || !p->block_id || !lk || !err_flag || !lk->n_lights || !lk->light_cell || !lk->ctrl_off ||
        !lk->ctrl_cell || !lk->inc_off || !lk->inc_cell) {
        set_error("tsim_layout_lights: bad arguments");
        return TSIM_ERR_CONFIG;
    }
    if (cfg->forward_traffic_light_range) { set_error("forward_traffic_light_range is not implemented on the GPU path"); return TSIM_ERR_UNSUPPORTED; }
    if (cfg->halo != 0 || cfg->rows != cfg->height) { set_error("tsim_layout_lights: run on the gathered grid"); return TSIM_ERR_UNSUPPORTED; }
    cudaStream_t cs = (cudaStream_t)stream;
    const int W = cfg->width, H = cfg->height;
    const long long n = (long long)W * H;
    const int ntiles = div_up(n, SCAN_TILE);
    const int rtiles = div_up(W, 64) * div_up(H, RTH);
    // workspace layout
    char *w = (char *)workspace;
    size_t o = 0;
    auto take = [&](size_t bytes) { char *q = w + o; o += (bytes + 255) & ~(size_t)255; return q; };
    int32_t *scal = (int32_t *)take(64 * 4);   // [0..1] pivot, [4..5] link totals, [8..9] wave counters, [10] waves
    const int wp = div_up(W, 64);
    BitPlanes bp;
    for (int k = 0; k < 4; k++) bp.ar[k] = (unsigned long long *)take((size_t)wp * H * 8);
    bp.fw = (unsigned long long *)take((size_t)wp * H * 8);
    bp.bw = (unsigned long long *)take((size_t)wp * H * 8);
    bp.wp = wp;
    uint8_t *F = (uint8_t *)take(n);
    int32_t *lid = (int32_t *)take(n * 4);
    int32_t *tile_cnt = (int32_t *)take((size_t)ntiles * 4);
    uint8_t *dirty0 = (uint8_t *)take(rtiles), *dirty1 = (uint8_t *)take(rtiles);
    int32_t *cnt_ctrl = (int32_t *)take((size_t)lk->cap_lights * 4), *cnt_inc = (int32_t *)take((size_t)lk->cap_lights * 4);
    int32_t *scan_tmp = (int32_t *)take((size_t)(div_up(lk->cap_lights, SCAN_TILE) + 1) * 4);
    if (!workspace || o > ws_bytes) { set_error("tsim_layout_lights needs %zu workspace bytes, got %zu", o, ws_bytes); return TSIM_ERR_WORKSPACE; }

    TSIM_CUDA(cudaMemsetAsync(scal, 0, 64 * 4, cs));
    init_pivot_kernel<<<1, 1, 0, cs>>>(scal);
    TSIM_LAUNCH_CHECK();
    TSIM_CUDA(cudaMemsetAsync(dirty0, 0, rtiles, cs));
    TSIM_CUDA(cudaMemsetAsync(dirty1, 0, rtiles, cs));
    pivot_kernel<<<div_up(n, 256), 256, 0, cs>>>(n, (long long)(H / 2) * W, p->cell_type, p->dirs, scal);
    TSIM_LAUNCH_CHECK();
    pack_arrows_kernel<<<div_up((long long)wp * H * 32, 256), 256, 0, cs>>>(W, H, p->dirs, bp);
    TSIM_LAUNCH_CHECK();
    reach_seed_kernel<<<1, 1, 0, cs>>>(W, scal, bp, dirty0);
    TSIM_LAUNCH_CHECK();
    {
        int dev = 0, sms = 0, per_sm = 0;
        TSIM_CUDA(cudaGetDevice(&dev));
        TSIM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        TSIM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reach_kernel, 256, 0));
        int grid = sms * (per_sm < 1 ? 1 : per_sm);
        if (grid > div_up(rtiles, 8)) grid = div_up(rtiles, 8);
        int Wv = W, Hv = H;
        int32_t *counter = scal + 8;
        void *args[] = {&Wv, &Hv, &bp, &dirty0, &dirty1, &counter};
        TSIM_CUDA(cudaLaunchCooperativeKernel((const void *)reach_kernel, dim3(grid), dim3(256), args, 0, cs));
    }
    mark_cr_kernel<<<div_up(n, 256), 256, 0, cs>>>(W, H, p->cell_type, p->dirs, F);
    TSIM_LAUNCH_CHECK();
    LightsCtx L{W, H, cfg->traffic_light_range, lk->cap_lights, p->cell_type, p->dirs, bp.fw, bp.bw, wp, F, err_flag};
    lights_pass_kernel<0><<<div_up(n, 128), 128, 0, cs>>>(L, F, p->aux, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0);
    TSIM_LAUNCH_CHECK();
    // compact the lights in ascending cell order
    flag_count_kernel<<<ntiles, 256, 0, cs>>>(n, F, F_TL, tile_cnt);
    TSIM_LAUNCH_CHECK();
    scan_tiles_kernel<<<1, 1024, 0, cs>>>(ntiles, tile_cnt, lk->n_lights);
    TSIM_LAUNCH_CHECK();
    flag_rank_kernel<<<ntiles, 256, 0, cs>>>(n, F, F_TL, tile_cnt, lid, lk->light_cell, lk->cap_lights, err_flag);
    TSIM_LAUNCH_CHECK();
    // count -> offsets -> fill
    TSIM_CUDA(cudaMemsetAsync(cnt_ctrl, 0, (size_t)lk->cap_lights * 4, cs));
    TSIM_CUDA(cudaMemsetAsync(cnt_inc, 0, (size_t)lk->cap_lights * 4, cs));
    lights_pass_kernel<1><<<div_up(n, 128), 128, 0, cs>>>(L, F, p->aux, lid, cnt_ctrl, cnt_inc, nullptr, nullptr, nullptr, nullptr, 0, 0);
    TSIM_LAUNCH_CHECK();
    TSIM_CUDA(cudaMemcpyAsync(lk->ctrl_off, cnt_ctrl, (size_t)lk->cap_lights * 4, cudaMemcpyDeviceToDevice, cs));
    TSIM_CUDA(cudaMemcpyAsync(lk->inc_off, cnt_inc, (size_t)lk->cap_lights * 4, cudaMemcpyDeviceToDevice, cs));
    if ((st = exclusive_scan_i32(lk->ctrl_off, lk->cap_lights, scan_tmp, scal + 4, cs)) != TSIM_OK) return st;
    if ((st = exclusive_scan_i32(lk->inc_off, lk->cap_lights, scan_tmp, scal + 5, cs)) != TSIM_OK) return st;
    close_offsets_kernel<<<1, 1, 0, cs>>>(lk->n_lights, lk->ctrl_off, lk->inc_off, scal + 4, lk->cap_lights, lk->cap_ctrl, lk->cap_inc, err_flag);
    TSIM_LAUNCH_CHECK();
    TSIM_CUDA(cudaMemsetAsync(cnt_ctrl, 0, (size_t)lk->cap_lights * 4, cs));
    TSIM_CUDA(cudaMemsetAsync(cnt_inc, 0, (size_t)lk->cap_lights * 4, cs));
    lights_pass_kernel<2><<<div_up(n, 128), 128, 0, cs>>>(L, F, p->aux, lid, cnt_ctrl, cnt_inc, lk->ctrl_off, lk->inc_off, lk->ctrl_cell,
                                                          lk->inc_cell, lk->cap_ctrl, lk->cap_inc);
    TSIM_LAUNCH_CHECK();
    lights_apply_kernel<<<div_up(n, 256), 256, 0, cs>>>(n, p->cell_type, p->dirs, p->aux, p->block_id, F);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}
