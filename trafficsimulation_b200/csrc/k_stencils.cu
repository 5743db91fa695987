// k_stencils.cu -- the 1-halo stencil passes over the packed planes:
//   tsim_layout_dead_ends   (city_model.py:811-840)
//   tsim_layout_upgrade_r2  (city_model.py:842-879, 211-306)
//   tsim_layout_fix_dirs    (city_model.py:969-1012 then :1035-1070)
//   tsim_maps               (city_model.py:2151-2199)
//
// All of them are SPARSE: only road cells (~30 % of a city) can change, and a road cell changes
// only by looking at its 4 neighbours.  Each thread streams 16 cells of the type plane with one
// 128-bit load, rejects strips without candidate types with a couple of ALU ops, and touches the
// neighbours (L1/L2 hits: they are the lines its own warp or the adjacent row's warp just loaded)
// only for candidates.  DRAM traffic is therefore the 1 B/cell type plane plus the few changed
// sectors -- the algorithmic bytes of SURVEY.md §8(d).
#include <cooperative_groups.h>
#include "cells_stencil.cuh"
#include "bitplane.cuh"

namespace cg = cooperative_groups;

namespace tsim {

struct Shard {   // row window of a shard allocation (global row numbers)
    int W, H, y0, nrows, ylo, yhi;   // allocation holds rows [y0, y0+nrows) = [ylo, yhi); rows outside do not exist for the call
    __host__ explicit Shard(const tsim_cfg &c) {
        W = c.width; H = c.height; y0 = c.win_y0; nrows = c.win_rows; ylo = y0; yhi = y0 + nrows;
    }
};

__device__ __forceinline__ PlaneView make_view(const Shard &s, const uint8_t *T, const uint16_t *D, const uint8_t *A) {
    PlaneView v; v.T = T; v.D = D; v.A = A; v.W = s.W; v.H = s.H; v.y0 = s.y0; v.nrows = s.nrows;
    return v;
}

// Generic sparse sweep: calls f(x, y) for every owned cell whose type is in `set`.
template <class F>
__device__ __forceinline__ void sparse_sweep(const Shard &s, const uint8_t *T, uint32_t set, F f, long long tid, long long nthreads) {
    if ((s.W & 15) == 0) {
        const uint32_t sw = (uint32_t)s.W >> 4;
        const uint32_t nstrips = sw * (uint32_t)(s.yhi - s.ylo);
        for (uint32_t i = (uint32_t)tid; i < nstrips; i += (uint32_t)nthreads) {
            const uint32_t ry = i / sw;
            const int y = s.ylo + (int)ry, x0 = (int)(i - ry * sw) << 4;
            const uint4 q = __ldcg(reinterpret_cast<const uint4 *>(T + (size_t)(y - s.y0) * s.W + x0));
            uint32_t m = strip_set_mask(q, set);
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                f(x0 + b, y);
            }
        }
    } else {
        const long long n = (long long)s.W * (s.yhi - s.ylo);
        for (long long i = tid; i < n; i += nthreads) {
            const int y = s.ylo + (int)(i / s.W), x = (int)(i % s.W);
            const int t = T[(size_t)(y - s.y0) * s.W + x];
            if ((set >> t) & 1u) f(x, y);
        }
    }
}

// type sets as exclusive byte ranges (enum order in common.cuh: R1 9, R2 10, R3 11, Intersection 12, HighwayEntrance 13,
// HighwayExit 14, BlockEntrance 19, Sidewalk 7)
constexpr TypeRanges RNG_ROAD_LIKE{T_R1 - 1, T_HWY_OUT + 1, T_BE - 1, T_BE + 1};
constexpr TypeRanges RNG_REMOVABLE{T_R2 - 1, T_INTER + 1, 0, 0};
constexpr TypeRanges RNG_R2{T_R2 - 1, T_R2 + 1, 0, 0};
constexpr TypeRanges RNG_SIDEWALK{T_SIDEWALK - 1, T_SIDEWALK + 1, 0, 0};
constexpr TypeRanges RNG_BE{T_BE - 1, T_BE + 1, 0, 0};
constexpr TypeRanges RNG_INTER{T_INTER - 1, T_INTER + 1, 0, 0};
static_assert(SET_ROAD_LIKE == (M(T_R1) | M(T_R2) | M(T_R3) | M(T_INTER) | M(T_HWY_IN) | M(T_HWY_OUT) | M(T_BE)), "RNG_ROAD_LIKE");
static_assert(SET_REMOVABLE == (M(T_R2) | M(T_R3) | M(T_INTER)), "RNG_REMOVABLE");

// The 4-neighbourhood of a 16-cell strip held in registers: the strip itself, the strips above and below it and the
// two cells left and right of it.  The neighbour taps of a candidate cell then cost no memory instruction at all.
struct StripView {
    uint32_t r[3][4];   // rows yc-1, yc, yc+1 (16 type bytes each; 0xff.. when the row does not exist)
    int left, right;    // cells (x0-1, yc) and (x0+16, yc), -1 outside the grid
    int x0, yc;
    const uint16_t *D; int W, y0;
    __device__ __forceinline__ static int byte_of(const uint32_t (&w)[4], int k) {
        const uint32_t v = k < 8 ? (k < 4 ? w[0] : w[1]) : (k < 12 ? w[2] : w[3]);
        return (int)((v >> (8 * (k & 3))) & 0xffu);
    }
    __device__ __forceinline__ int t(int x, int y) const {   // (x, y) in the 4-neighbourhood of a strip cell
        const int k = x - x0;
        if (y == yc) return k < 0 ? left : (k > 15 ? right : byte_of(r[1], k));
        const int v = byte_of(r[y < yc ? 0 : 2], k);
        return v == 0xff ? -1 : v;
    }
    __device__ __forceinline__ uint32_t d(int x, int y) const { return D[(size_t)(y - y0) * W + x]; }
    __device__ __forceinline__ size_t at(int x, int y) const { return (size_t)(y - y0) * W + x; }
    // 16-bit masks "the neighbour of strip cell k in that direction has a type of `rg`"
    struct Nbr { uint32_t c, up, dn, lf, rt; };
    __device__ __forceinline__ Nbr neighbours(const TypeRanges rg, bool in_left, bool in_right) const {
        Nbr n;
        n.c = strip_range_mask(r[1], rg); n.up = strip_range_mask(r[2], rg); n.dn = strip_range_mask(r[0], rg);
        n.lf = ((n.c << 1) | (uint32_t)in_left) & 0xffffu;
        n.rt = (n.c >> 1) | ((uint32_t)in_right << 15);
        return n;
    }
    __device__ __forceinline__ bool edge_in(int t, const TypeRanges rg) const {
        return t >= 0 && (((uint32_t)t > rg.lo1 && (uint32_t)t < rg.hi1) || ((uint32_t)t > rg.lo2 && (uint32_t)t < rg.hi2));
    }
};
__device__ __forceinline__ uint32_t at_least_two(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return (a & b) | (c & d) | ((a | b) & (c | d));
}

// sparse sweep with the neighbourhood in registers: f(x, y, view) for every cell whose type is in `set` (W % 16 == 0)
// `refine(view, mask)` narrows the candidate mask with word-level logic on the neighbour masks before the per-cell work
template <class R, class F>
__device__ __forceinline__ void sparse_sweep3(const Shard &s, const uint8_t *T, const uint16_t *D, const TypeRanges set, R refine, F f, long long tid,
                                              long long nthreads) {
    // strip index arithmetic in 32 bits (a window holds < 2^31 cells, so < 2^27 strips): the 64-bit division this loop used to do
    // per strip is a ~100-instruction software routine
    const uint32_t sw = (uint32_t)s.W >> 4;
    const uint32_t nstrips = sw * (uint32_t)(s.yhi - s.ylo);
    for (uint32_t i = (uint32_t)tid; i < nstrips; i += (uint32_t)nthreads) {
        const uint32_t ry = i / sw;
        const int y = s.ylo + (int)ry, x0 = (int)(i - ry * sw) << 4;
        const uint8_t *row = T + (size_t)(y - s.y0) * s.W + x0;
        const uint4 q = *reinterpret_cast<const uint4 *>(row);
        const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
        uint32_t m = strip_range_mask(qw, set);
        if (!m) continue;
        StripView v;
        const uint4 none = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        const uint4 qd = y > s.y0 ? *reinterpret_cast<const uint4 *>(row - s.W) : none;
        const uint4 qu = y + 1 < s.y0 + s.nrows ? *reinterpret_cast<const uint4 *>(row + s.W) : none;
        v.r[0][0] = qd.x; v.r[0][1] = qd.y; v.r[0][2] = qd.z; v.r[0][3] = qd.w;
        v.r[1][0] = q.x; v.r[1][1] = q.y; v.r[1][2] = q.z; v.r[1][3] = q.w;
        v.r[2][0] = qu.x; v.r[2][1] = qu.y; v.r[2][2] = qu.z; v.r[2][3] = qu.w;
        v.left = x0 > 0 ? (int)row[-1] : -1;
        v.right = x0 + 16 < s.W ? (int)row[16] : -1;
        v.x0 = x0; v.yc = y; v.D = D; v.W = s.W; v.y0 = s.y0;
        m = refine(v, m);
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            f(x0 + b, y, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// dead ends: persistent cooperative kernel, sweeps until a whole sweep changes nothing.
// Removal is monotone (a removed cell never comes back and only lowers its neighbours' counts), so
// the fixed point is unique and any asynchronous order reaches it; stale reads can only delay a
// removal, never cause a wrong one.  A thread that removes a cell follows the stub it just exposed.
// ------------------------------------------------------------------------------------------------
struct CoherentView {   // L2-coherent reads (bypass L1) for planes that other SMs modify in this kernel
    uint8_t *T; int W, H, y0, y1;
    // A row outside the window but inside the grid is UNKNOWN here: it counts as a (non-removable) road, so a
    // cell next to a shard cut is never removed on missing information -- at worst its removal waits for the
    // halo exchange (removal is monotone, so waiting never changes the fixed point).
    __device__ __forceinline__ int t(int x, int y) const {
        if (x < 0 || x >= W || y < 0 || y >= H) return -1;
        if (y < y0 || y >= y1) return T_R1;
        return (int)__ldcg(T + (size_t)(y - y0) * W + x);
    }
};

__global__ void __launch_bounds__(256) dead_ends_kernel(Shard s, uint8_t *T, uint16_t *D, uint8_t *A, int32_t *flags /* [0]=changed, [1]=sweeps */) {
    cg::grid_group grid = cg::this_grid();
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (long long)gridDim.x * blockDim.x;
    CoherentView v{T, s.W, s.H, s.y0, s.y0 + s.nrows};
    int sweeps = 0;
    for (;;) {
        sweeps++;
        bool changed = false;
        auto remove_from = [&](int x, int y) {
            int cx = x, cy = y;
            for (int guard = 0; guard < (1 << 20); guard++) {
                if (!dead_end_cell(v, cx, cy)) break;
                const size_t i = (size_t)(cy - s.y0) * s.W + cx;
                T[i] = T_SIDEWALK; D[i] = 0; A[i] &= (AUX_RING | AUX_EVER);   // place_cell(..., "Sidewalk")
                changed = true;
                // follow the stub: the single remaining road-like neighbour, if it is removable and owned
                int nx = -1, ny = -1;
                const int ox[4] = {1, -1, 0, 0}, oy[4] = {0, 0, 1, -1};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int t = v.t(cx + ox[k], cy + oy[k]);
                    if (t >= 0 && in_set(SET_REMOVABLE, t)) { nx = cx + ox[k]; ny = cy + oy[k]; }
                }
                if (nx < 0 || ny < s.ylo || ny >= s.yhi) break;
                cx = nx; cy = ny;
            }
        };
        if (sweeps == 1 && (s.W & 15) == 0) {
            // First sweep: candidates are found from register strips (plain cached loads).  A stale type can only show
            // MORE road neighbours than there are now, i.e. miss a dead end that another thread is creating right now --
            // and that thread follows the stub it exposes with coherent loads.  If this sweep removes nothing, nothing
            // was written at all and the result is exact; otherwise the coherent sweeps below run to the fixed point.
            sparse_sweep3(s, T, D, RNG_REMOVABLE, [&](const StripView &sv, uint32_t m) {
                StripView::Nbr n = sv.neighbours(RNG_ROAD_LIKE, sv.edge_in(sv.left, RNG_ROAD_LIKE), sv.edge_in(sv.right, RNG_ROAD_LIKE));
                // rows beyond a shard cut are unknown = road (see CoherentView)
                if (sv.yc + 1 >= s.y0 + s.nrows && sv.yc + 1 < s.H) n.up = 0xffffu;
                if (sv.yc - 1 < s.y0 && sv.yc - 1 >= 0) n.dn = 0xffffu;
                return m & ~at_least_two(n.up, n.dn, n.lf, n.rt);
            }, [&](int x, int y, const StripView &) { remove_from(x, y); }, tid, nthreads);
        } else {
            sparse_sweep(s, T, SET_REMOVABLE, remove_from, tid, nthreads);
        }
        if (changed) flags[0] = 1;
        grid.sync();
        const int any = *((volatile int32_t *)flags);
        grid.sync();
        if (tid == 0) { flags[0] = 0; flags[1] = sweeps; }
        if (!any) break;
        grid.sync();
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upgrade_r2_kernel(tsim_cfg c, Shard s, uint8_t *T, uint16_t *D, uint8_t *A,
                                                         const uint32_t *__restrict__ rowt, const uint32_t *__restrict__ colt, int32_t *err) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (long long)gridDim.x * blockDim.x;
    const PlaneView pv = make_view(s, T, D, A);
    auto cell = [&](int x, int y, const auto &v) {
        int t_new; uint32_t d_new;
        const int r = upgrade_r2_cell(c, v, __ldg(rowt + y), __ldg(colt + x), x, y, t_new, d_new);
        if (r == 0) return;
        if (r == 3) { *err = 1; return; }
        const size_t i = pv.at(x, y);
        // in place is safe: the pass reads Sidewalk-ness and sub-block-road-ness of neighbours, which it
        // never changes (an R2 neighbour matters only when subblock_road_type == R2, where the cell's own
        // type already decides).
        T[i] = (uint8_t)t_new; D[i] = (uint16_t)d_new;
        const uint8_t a = A[i] & (AUX_RING | AUX_EVER);
        A[i] = r == 1 ? (uint8_t)(a | AUX_EVER) : (uint8_t)(a & ~AUX_EVER);
    };
    if ((s.W & 15) == 0)
        sparse_sweep3(s, T, D, RNG_R2, [&](const StripView &sv, uint32_t m) {   // only R2 cells with >= 2 Sidewalk neighbours can change (:850-856)
            const StripView::Nbr n = sv.neighbours(RNG_SIDEWALK, sv.left == T_SIDEWALK, sv.right == T_SIDEWALK);
            return m & at_least_two(n.up, n.dn, n.lf, n.rt);
        }, cell, tid, nthreads);
    else sparse_sweep(s, T, M(T_R2), [&](int x, int y) { cell(x, y, pv); }, tid, nthreads);
}

__global__ void __launch_bounds__(256) validate_dirs_kernel(Shard s, const uint8_t *T, uint16_t *D) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (long long)gridDim.x * blockDim.x;
    const PlaneView pv = make_view(s, T, D, nullptr);
    // in place is safe: only Intersection lists are written, and an Intersection neighbour's list is never read
    auto cell = [&](int x, int y, const auto &v) {
        const size_t i = pv.at(x, y);
        const uint32_t od = D[i], nd = validate_dirs_cell(v, x, y, od);
        if (nd != od) D[i] = (uint16_t)nd;
    };
    if ((s.W & 15) == 0) sparse_sweep3(s, T, D, RNG_INTER, [](const StripView &, uint32_t m) { return m; }, cell, tid, nthreads);
    else sparse_sweep(s, T, M(T_INTER), [&](int x, int y) { cell(x, y, pv); }, tid, nthreads);
}

__global__ void __launch_bounds__(256) entrance_dirs_kernel(Shard s, const uint8_t *T, uint16_t *D) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (long long)gridDim.x * blockDim.x;
    const PlaneView pv = make_view(s, T, D, nullptr);
    // in place is safe: a cell's new list depends on its own old list and on neighbour TYPES only
    if ((s.W & 15) == 0) {
        // register strips; the dirs word is only touched when an entrance is involved (the cell is one, or has one next to it)
        sparse_sweep3(s, T, D, RNG_ROAD_LIKE, [&](const StripView &sv, uint32_t m) {
            const StripView::Nbr n = sv.neighbours(RNG_BE, sv.left == T_BE, sv.right == T_BE);
            return m & (n.c | n.up | n.dn | n.lf | n.rt);
        }, [&](int x, int y, const StripView &v) {
            const int t = v.t(x, y);
            const size_t i = pv.at(x, y);
            const uint32_t od = D[i], nd = entrance_dirs_cell(v, x, y, t, od);
            if (nd != od) D[i] = (uint16_t)nd;
        }, tid, nthreads);
    } else {
        sparse_sweep(s, T, SET_ROAD_LIKE, [&](int x, int y) {
            const size_t i = pv.at(x, y);
            const uint32_t od = D[i], nd = entrance_dirs_cell(pv, x, y, (int)T[i], od);
            if (nd != od) D[i] = (uint16_t)nd;
        }, tid, nthreads);
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) maps_kernel(long long n, const uint8_t *__restrict__ T, const uint16_t *__restrict__ D,
                                                   const uint8_t *__restrict__ A, uint8_t *__restrict__ o_road, uint8_t *__restrict__ o_type,
                                                   uint8_t *__restrict__ o_int, uint8_t *__restrict__ o_dirs) {
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (i0 >= n) return;
    if (VEC == 16) {
        const uint4 tq = __ldg(reinterpret_cast<const uint4 *>(T + i0)), aq = __ldg(reinterpret_cast<const uint4 *>(A + i0));
        const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(D + i0)), d1 = __ldg(reinterpret_cast<const uint4 *>(D + i0 + 8));
        const uint32_t tw[4] = {tq.x, tq.y, tq.z, tq.w}, aw[4] = {aq.x, aq.y, aq.z, aq.w};
        const uint32_t dw[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        uint32_t r[4] = {0, 0, 0, 0}, ty[4] = {0, 0, 0, 0}, in[4] = {0, 0, 0, 0}, al[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int t = (tw[k >> 2] >> (8 * (k & 3))) & 0xff;
            const uint32_t a = (aw[k >> 2] >> (8 * (k & 3))) & 0xff, d = (dw[k >> 1] >> (16 * (k & 1))) & 0xffff;
            uint8_t q0, q1, q2, q3;
            maps_cell(t, d, a, q0, q1, q2, q3);
            r[k >> 2] |= (uint32_t)q0 << (8 * (k & 3)); ty[k >> 2] |= (uint32_t)q1 << (8 * (k & 3));
            in[k >> 2] |= (uint32_t)q2 << (8 * (k & 3)); al[k >> 2] |= (uint32_t)q3 << (8 * (k & 3));
        }
        *reinterpret_cast<uint4 *>(o_road + i0) = make_uint4(r[0], r[1], r[2], r[3]);
        *reinterpret_cast<uint4 *>(o_type + i0) = make_uint4(ty[0], ty[1], ty[2], ty[3]);
        *reinterpret_cast<uint4 *>(o_int + i0) = make_uint4(in[0], in[1], in[2], in[3]);
        *reinterpret_cast<uint4 *>(o_dirs + i0) = make_uint4(al[0], al[1], al[2], al[3]);
    } else {
        maps_cell(T[i0], D[i0], A[i0], o_road[i0], o_type[i0], o_int[i0], o_dirs[i0]);
    }
}

static int coop_grid(const void *kernel, int threads) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

static int sweep_grid(const Shard &s) {
    const long long strips = ((s.W & 15) == 0) ? (long long)(s.W >> 4) * (s.yhi - s.ylo) : (long long)s.W * (s.yhi - s.ylo);
    long long blocks = (strips + 255) / 256;
    const long long cap = 148LL * 32;   // persistent-ish: 32 CTAs per SM worth of work per launch, grid-stride beyond
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace tsim

using namespace tsim;

static tsim_status check_planes(const tsim_planes *p, const char *who) {
    if (!p || !p->cell_type || !p->dirs || !p->aux || !p->block_id) { set_error("%s: NULL plane", who); return TSIM_ERR_CONFIG; }
    if (((uintptr_t)p->cell_type | (uintptr_t)p->dirs | (uintptr_t)p->aux | (uintptr_t)p->block_id) & 15) {
        set_error("%s: planes must be 16-byte aligned", who);
        return TSIM_ERR_CONFIG;
    }
    return TSIM_OK;
}

extern "C" tsim_status tsim_layout_dead_ends(const tsim_cfg *cfg, const tsim_planes *p, int32_t *sweeps, void *workspace,
                                             size_t ws_bytes, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if ((st = check_planes(p, "tsim_layout_dead_ends")) != TSIM_OK) return st;
    if (!workspace || ws_bytes < 64) { set_error("tsim_layout_dead_ends: workspace too small"); return TSIM_ERR_WORKSPACE; }
    cudaStream_t cs = (cudaStream_t)stream;
    Shard s(*cfg);
    int32_t *flags = (int32_t *)workspace;
    TSIM_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(int32_t), cs));
    int grid = coop_grid((const void *)dead_ends_kernel, 256);
    uint8_t *T = p->cell_type; uint16_t *D = p->dirs; uint8_t *A = p->aux;
    void *args[] = {&s, &T, &D, &A, &flags};
    TSIM_COOP_LAUNCH(dead_ends_kernel, dim3(grid), dim3(256), args, cs);
    if (sweeps) TSIM_CUDA(cudaMemcpyAsync(sweeps, flags + 1, sizeof(int32_t), cudaMemcpyDeviceToDevice, cs));
    return TSIM_OK;
}

extern "C" tsim_status tsim_layout_upgrade_r2(const tsim_cfg *cfg, const tsim_planes *p, const tsim_lines *lines, int32_t *err_flag,
                                              void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if ((st = check_planes(p, "tsim_layout_upgrade_r2")) != TSIM_OK) return st;
    if (!lines || !lines->row || !lines->col) { set_error("tsim_layout_upgrade_r2: NULL line table"); return TSIM_ERR_CONFIG; }
    if (!err_flag) { set_error("tsim_layout_upgrade_r2: NULL err_flag"); return TSIM_ERR_CONFIG; }
    cudaStream_t cs = (cudaStream_t)stream;
    Shard s(*cfg);
    upgrade_r2_kernel<<<sweep_grid(s), 256, 0, cs>>>(*cfg, s, p->cell_type, p->dirs, p->aux, lines->row, lines->col, err_flag);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}

extern "C" tsim_status tsim_layout_fix_dirs(const tsim_cfg *cfg, const tsim_planes *p, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if ((st = check_planes(p, "tsim_layout_fix_dirs")) != TSIM_OK) return st;
    cudaStream_t cs = (cudaStream_t)stream;
    Shard s(*cfg);
    validate_dirs_kernel<<<sweep_grid(s), 256, 0, cs>>>(s, p->cell_type, p->dirs);
    TSIM_LAUNCH_CHECK();
    entrance_dirs_kernel<<<sweep_grid(s), 256, 0, cs>>>(s, p->cell_type, p->dirs);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}

extern "C" tsim_status tsim_maps(const tsim_cfg *cfg, const tsim_planes *p, uint8_t *is_road, uint8_t *road_type, uint8_t *intersection,
                                 uint8_t *allowed_dirs, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if ((st = check_planes(p, "tsim_maps")) != TSIM_OK) return st;
    if (!is_road || !road_type || !intersection || !allowed_dirs) { set_error("tsim_maps: NULL output"); return TSIM_ERR_CONFIG; }
    cudaStream_t cs = (cudaStream_t)stream;
    // maps are produced for the whole window; outputs are [win_rows][W]
    const long long n = (long long)cfg->width * cfg->win_rows;
    const size_t off = 0;
    const bool aligned = (n % 16 == 0) && (off % 16 == 0) &&
                         !(((uintptr_t)is_road | (uintptr_t)road_type | (uintptr_t)intersection | (uintptr_t)allowed_dirs) & 15);
    if (aligned)
        maps_kernel<16><<<div_up(n, 16 * 256), 256, 0, cs>>>(n, p->cell_type + off, p->dirs + off, p->aux + off, is_road, road_type, intersection, allowed_dirs);
    else
        maps_kernel<1><<<div_up(n, 256), 256, 0, cs>>>(n, p->cell_type + off, p->dirs + off, p->aux + off, is_road, road_type, intersection, allowed_dirs);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}
