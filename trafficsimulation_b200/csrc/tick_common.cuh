// tick_common.cuh -- what the two tick kernels share: constants of the reference (config.py), the light-group controller
// (intersection_light_group.py:396-494) with its staged stop_map writes, and the generation-tagged claim words.
#pragma once
#include "common.cuh"

namespace tsim {

constexpr int MIN_GREEN = 5, MAX_GREEN = 30, GAP_TICKS = 3, GREEN_DURATION = 20;   // config.py:354-359
constexpr int AWARENESS = 10, MALFUNCTION_TICKS = 400, COLLISION_TICKS = 600, STUCK_THRESHOLD = 30, RAIN_REDUCTION = 2;   // config.py:279,321,326,306,263
constexpr int NO_CLAIM = 0x7fffffff;
constexpr int MAX_SPEED = 5;          // random.randint(1, 5), vehicle_base.py:112: a vehicle plans at most 5 cells
constexpr int GEN_PER_TICK = 1024;   // claim generations per tick: sweeps 1..1022, spawner 1023
typedef unsigned long long u64;
enum { S_TICK = 0, S_ERR = 1, S_ITERS = 2, S_UPDATES = 3, S_FLAG0 = 4, S_FLAG1 = 5, S_UPD_HI = 6 /* 64-bit: [6..7] */, S_FLAG2 = 8, S_XERR = 9 /* shard exchange */,
       S_NLIST = 12, S_LIST_OK = 13 /* live_idx: entries, valid for the next tick */, S_NCAND = 14 /* sideswipe candidates of the tick */, S_NCONT = 15 /* vehicles in the claim fixed point of the tick */ };

// Live-list kernel: what a vehicle needs to know about the cells ahead lives in BIT PLANES over 8 x 8-cell tiles (one 64-bit word
// per tile and plane: bit (y & 7) * 8 + (x & 7) of word (y >> 3) * tiles_x + (x >> 3)).  A straight run of five cells lies in at
// most two tiles, so a look-ahead is two loads per plane and a mark / an occupancy change one or two atomics whatever the
// direction of travel; the hot planes of an 8192 x 8192 city are 8 MB each and stay in L2.
enum { PL_OCC = 0,    // occupancy_map
       PL_STOP = 1,   // stop_map (committed)
       PL_T1 = 2,     // the cell lies on the planned cells of a vehicle this tick ...
       PL_T2 = 3,     // ... of more than one, or a light group staged a stop_map write on it: whoever plans to enter it must look closer
       PL_STG = 4,    // a light group staged a stop_map write this tick (value in `stopw`)
       PL_WANT = 5,   // a sideswipe candidate asks which vehicle stands here (this tick's phase A only)
       N_PLANES = 6 };

constexpr int OUTSIDE = -2;          // a tape cell that lies outside this shard's window (host-side translation, tsim.h)

struct TickArgs {
    int W, H, n_ticks, algo;   // H = rows of the window; every cell index below is LOCAL to the window
    int own_lo, own_hi;        // local cell range of the rows this shard owns (vehicle_updates counts only those)
    tsim_light_tables lt;
    tsim_tick_tapes tp;
    tsim_tick_state st;
    int sort_every;                           // ... every so many ticks
    int tile_sx, tile_sy, tiles_x, n_tiles;   // live-list kernel, sorted append: tile = (y >> tile_sy) * tiles_x + (x >> tile_sx); n_tiles == 0: plain append
    // live-list kernel: the bit planes (tsim_tick_state.probe), words per plane, tiles per row
    unsigned long long *bits;
    unsigned long long w_magic;   // ceil(2^64 / W)
    long long n_tw;
    int occ_tiles_x;
    // live-list kernel, light groups (tsim_tick_state.group_ws, built by tsim_tick_init): per group, its incoming lanes and its
    // cluster as (tile, mask) pairs over the occupancy plane
    const unsigned long long *occ;   // = bits (plane PL_OCC)
    const unsigned long long *gq_mask;
    const int32_t *gq_tile, *gq_cnt;
    int gq_base_ew, gq_base_cl;
    // ... and the cells its lights control, flattened: group -> (cell, role) with role 0 = a light of g_all, 1 = of g_ns, 2 = of g_ew
    const int32_t *gc_off, *gc_cell;
    const uint8_t *gc_role;
};

__device__ __forceinline__ void cell_wb(const TickArgs &a, int c, int &word, int &bit) {
    const int y = (int)__umul64hi((unsigned long long)(unsigned)c, a.w_magic), x = c - y * a.W;   // c / W by a multiplication (exact: c * W < 2^64)
    word = (y >> 3) * a.occ_tiles_x + (x >> 3);
    bit = (y & 7) * 8 + (x & 7);
}
__device__ __forceinline__ unsigned long long *bit_plane(const TickArgs &a, int plane) { return a.bits + (size_t)plane * a.n_tw; }
__device__ __forceinline__ int bit_get(const TickArgs &a, int plane, int c) {
    int w, b;
    cell_wb(a, c, w, b);
    return (int)((__ldcg(bit_plane(a, plane) + w) >> b) & 1ull);
}
__device__ __forceinline__ void bit_set(const TickArgs &a, int plane, int c) { int w, b; cell_wb(a, c, w, b); atomicOr(bit_plane(a, plane) + w, 1ull << b); }
__device__ __forceinline__ void bit_clear(const TickArgs &a, int plane, int c) { int w, b; cell_wb(a, c, w, b); atomicAnd(bit_plane(a, plane) + w, ~(1ull << b)); }
// vehicles on the cells of list `kind` (0 N-S lanes, 1 W-E lanes, 2 cluster) of group g; a cell listed twice counts twice (its second
// mention sits in an entry of its own)
__device__ __forceinline__ int occ_count(const TickArgs &a, int kind, int base, int g) {
    const int n = a.gq_cnt[kind * a.lt.n_groups + g];
    int q = 0;
    for (int i0 = 0; i0 < n; i0 += 4) {   // four entries at a time: the tile loads, then the occupancy loads, travel together
        int t[4];
        unsigned long long m[4], o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { t[j] = i0 + j < n ? a.gq_tile[base + i0 + j] : -1; m[j] = i0 + j < n ? a.gq_mask[base + i0 + j] : 0ull; }
#pragma unroll
        for (int j = 0; j < 4; j++) o[j] = t[j] >= 0 ? __ldcg(a.occ + t[j]) : 0ull;
#pragma unroll
        for (int j = 0; j < 4; j++) q += __popcll(o[j] & m[j]);
    }
    return q;
}
// both lane lists of a group at once (their loads overlap)
__device__ __forceinline__ void occ_count_lanes(const TickArgs &a, int g, int &ns_q, int &ew_q) {
    const int ng = a.lt.n_groups, n0 = a.gq_cnt[g], n1 = a.gq_cnt[ng + g], b0 = a.lt.g_nsin_off[g], b1 = a.gq_base_ew + a.lt.g_ewin_off[g];
    int q0 = 0, q1 = 0;
    for (int i0 = 0; i0 < max(n0, n1); i0 += 3) {
        int t[6];
        unsigned long long m[6], o[6];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            t[j] = i0 + j < n0 ? a.gq_tile[b0 + i0 + j] : -1; m[j] = i0 + j < n0 ? a.gq_mask[b0 + i0 + j] : 0ull;
            t[3 + j] = i0 + j < n1 ? a.gq_tile[b1 + i0 + j] : -1; m[3 + j] = i0 + j < n1 ? a.gq_mask[b1 + i0 + j] : 0ull;
        }
#pragma unroll
        for (int j = 0; j < 6; j++) o[j] = t[j] >= 0 ? __ldcg(a.occ + t[j]) : 0ull;
#pragma unroll
        for (int j = 0; j < 3; j++) { q0 += __popcll(o[j] & m[j]); q1 += __popcll(o[3 + j] & m[3 + j]); }
    }
    ns_q = q0; ew_q = q1;
}

template <class F>
__device__ __forceinline__ void for_light_cells(const tsim_light_tables &lt, const int32_t *off, const int32_t *lights, int g, F f) {
    for (int k = off[g]; k < off[g + 1]; k++) {
        const int l = lights[k];
        for (int q = lt.tl_off[l]; q < lt.tl_off[l + 1]; q++) f(lt.tl_cells[q]);
    }
}

// light group, controller (intersection_light_group.py:396-494): returns what the group does to its lights this tick -- 0 nothing,
// 1 all red (a phase change is pending but the intersection is occupied), 2 + p phase p goes (0 N-S, 1 W-E) -- and writes its state
template <bool PROBE>
__device__ __forceinline__ int group_controller(const TickArgs &a, int g) {
    const tsim_light_tables &lt = a.lt;
    const tsim_tick_state &s = a.st;
    int cur = s.g_cur[g], pend = s.g_pend[g];
    if (pend < 0) {
        if (a.algo == 0) {   // run_queue_actuated :463-494
            const int qt = ++s.g_qt[g];
            int ns_q = 0, ew_q = 0;
            if (PROBE) {
                occ_count_lanes(a, g, ns_q, ew_q);
            } else {
                for (int k = lt.g_nsin_off[g]; k < lt.g_nsin_off[g + 1]; k++) ns_q += (int)s.occupancy[lt.g_nsin[k]];
                for (int k = lt.g_ewin_off[g]; k < lt.g_ewin_off[g + 1]; k++) ew_q += (int)s.occupancy[lt.g_ewin[k]];
            }
            const int cur_q = cur == 0 ? ns_q : ew_q, opp_q = cur == 0 ? ew_q : ns_q;
            if (qt == 1) { s.g_last[g] = cur_q; s.g_gap[g] = 0; }
            if (cur_q > s.g_last[g]) { s.g_last[g] = cur_q; s.g_gap[g] = 0; } else s.g_gap[g]++;
            if (qt >= MIN_GREEN && (s.g_gap[g] >= GAP_TICKS || qt >= MAX_GREEN || (opp_q > cur_q && cur_q == 0))) {
                const int next = 1 - cur;
                if (next != cur && next != pend) pend = next;   // apply_phase :386-393
                s.g_qt[g] = 0;
            }
        } else if (a.algo == 2) {   // run_pressure_control :448-461 (compute_max_pressure numba_utilities.py:74-85): every tick without a
            // pending phase, the phase of the larger in-minus-out pressure.  The lists hold the cells the reference reads (tsim.h)
            auto occ_at = [&](int c) { return PROBE ? bit_get(a, PL_OCC, c) : (int)s.occupancy[c]; };
            int ns_p = 0, ew_p = 0;
            for (int k = lt.g_nsin_off[g]; k < lt.g_nsin_off[g + 1]; k++) ns_p += occ_at(lt.g_nsin[k]);
            for (int k = lt.g_nsout_off[g]; k < lt.g_nsout_off[g + 1]; k++) ns_p -= occ_at(lt.g_nsout[k]);
            for (int k = lt.g_ewin_off[g]; k < lt.g_ewin_off[g + 1]; k++) ew_p += occ_at(lt.g_ewin[k]);
            for (int k = lt.g_ewout_off[g]; k < lt.g_ewout_off[g + 1]; k++) ew_p -= occ_at(lt.g_ewout[k]);
            const int ph = ns_p > ew_p ? 0 : 1;
            if (ph != cur && ph != pend) pend = ph;   // apply_phase :386-393
        } else if (a.algo == 3) {   // run_neighbor_green_wave :522-546: the phase was settled for all groups by green_wave_prepass
            const int ph = s.g_wave[g];
            if (ph != cur && ph != pend) pend = ph;
        } else {             // run_fixed_time :427-441
            const int ft = ++s.g_ft_timer[g];
            if (ft == 1) { const int ph = s.g_ft_phase[g]; if (ph != cur && ph != pend) pend = ph; }
            if (ft >= GREEN_DURATION) { s.g_ft_phase[g] = 1 - s.g_ft_phase[g]; s.g_ft_timer[g] = 0; }
        }
    }
    int plan = 0;
    if (pend >= 0) {         // _execute_phase_change :348-384
        bool occupied = false;
        if (PROBE) occupied = occ_count(a, 2, a.gq_base_cl + lt.g_cl_off[g], g) != 0;
        else for (int k = lt.g_cl_off[g]; k < lt.g_cl_off[g + 1]; k++) occupied |= s.occupancy[lt.g_cl[k]] != 0;
        if (occupied) plan = 1;
        else { plan = 2 + pend; cur = pend; pend = -1; }
    }
    s.g_cur[g] = cur; s.g_pend[g] = pend; s.g_plan[g] = plan;
    return plan;
}

// NEIGHBOR_GREEN_WAVE (intersection_light_group.py:522-546): a group follows the phase its neighbours hold -- and a neighbour that steps
// EARLIER in the activation order (lower index) has already run its own step of this tick, commit included (:348-384).  That is a
// recurrence along the activation order; it is triangular, so it has one solution, and Jacobi sweeps over all groups reach it: after
// k sweeps every group whose chain of lower-ranked neighbours is shorter than k is final.  A sweep that changes nothing has read only
// final values, so the phases it computed (g_wave[g]) are the sequential ones.  Runs before phase 1 of a tick, on the tick-start
// occupancy; scratch: g_wave[0..ng) phase asked for (-1: a phase is pending, the controller does not run), [ng..2ng) phase held
// after this tick's step, [2ng..3ng) bit 0 = ns queue longer, bit 1 = cluster occupied, [3ng..3ng+3) sweep flags in rotation.
template <bool PROBE, class Grid>
__device__ void green_wave_prepass(const TickArgs &a, Grid &grid, int tid, int nth) {
    const tsim_light_tables &lt = a.lt;
    const tsim_tick_state &s = a.st;
    const int ng = lt.n_groups;
    int32_t *ph = s.g_wave, *c1 = s.g_wave + ng, *qq = s.g_wave + 2 * ng, *flags = s.g_wave + 3 * ng;
    for (int g = tid; g < ng; g += nth) {
        int ns_q = 0, ew_q = 0;
        bool occupied = false;
        if (PROBE) {
            occ_count_lanes(a, g, ns_q, ew_q);
            occupied = occ_count(a, 2, a.gq_base_cl + lt.g_cl_off[g], g) != 0;
        } else {
            for (int k = lt.g_nsin_off[g]; k < lt.g_nsin_off[g + 1]; k++) ns_q += (int)s.occupancy[lt.g_nsin[k]];
            for (int k = lt.g_ewin_off[g]; k < lt.g_ewin_off[g + 1]; k++) ew_q += (int)s.occupancy[lt.g_ewin[k]];
            for (int k = lt.g_cl_off[g]; k < lt.g_cl_off[g + 1]; k++) occupied |= s.occupancy[lt.g_cl[k]] != 0;
        }
        qq[g] = (ns_q > ew_q ? 1 : 0) | (occupied ? 2 : 0);
        c1[g] = s.g_cur[g];
    }
    if (tid == 0) { flags[0] = 0; flags[1] = 0; flags[2] = 0; }
    grid.sync();
    for (int iter = 0; iter <= ng + 1; iter++) {
        int32_t *flag = flags + iter % 3;
        bool changed = false;
        for (int g = tid; g < ng; g += nth) {
            const int cur0 = s.g_cur[g], pend0 = s.g_pend[g], q = qq[g];
            int p = -1, pend = pend0;
            if (pend0 < 0) {
                bool favor_ns = false, favor_ew = false;
#pragma unroll
                for (int d = 0; d < 4; d++) {
                    const int nb = lt.g_nbr[4 * g + d];
                    if (nb < 0) continue;
                    const int c = nb < g ? __ldcg(c1 + nb) : s.g_cur[nb];   // stepped before me this tick / steps after me
                    if (d < 2 && c == 0) favor_ns = true;
                    if (d >= 2 && c == 1) favor_ew = true;
                }
                p = (favor_ns && !favor_ew) ? 0 : (favor_ew && !favor_ns) ? 1 : ((q & 1) ? 0 : 1);
                if (p != cur0) pend = p;   // apply_phase :386-393 (nothing is pending here)
            }
            const int held = (pend >= 0 && !(q & 2)) ? pend : cur0;   // _execute_phase_change :348-384
            ph[g] = p;
            if (held != __ldcg(c1 + g)) { c1[g] = held; changed = true; }
        }
        if (changed) *flag = 1;
        if (tid == 0) flags[(iter + 1) % 3] = 0;   // written in the next sweep; last read two barriers ago
        grid.sync();
        if (!*((volatile int32_t *)flag)) break;
    }
}

// light group: controller + decision; stop_map writes are staged in `stopw` (atomicMax, priority =
// group index, then order inside the group) so that concurrent groups reproduce the sequential result
template <bool PROBE>
__device__ void group_decide(const TickArgs &a, int g) {
    const tsim_light_tables &lt = a.lt;
    const tsim_tick_state &s = a.st;
    const int plan = group_controller<PROBE>(a, g);
    if (plan == 0) return;
    const int base = (g + 1) * 4;
    if (plan == 1) {
        for_light_cells(lt, lt.g_all_off, lt.g_all, g, [&](int c) { atomicMax(s.stopw + c, base + 1); });
    } else {
        const bool ns_go = plan == 2;
        for_light_cells(lt, ns_go ? lt.g_ns_off : lt.g_ew_off, ns_go ? lt.g_ns : lt.g_ew, g, [&](int c) { atomicMax(s.stopw + c, base + 0); });
        for_light_cells(lt, ns_go ? lt.g_ew_off : lt.g_ns_off, ns_go ? lt.g_ew : lt.g_ns, g, [&](int c) { atomicMax(s.stopw + c, base + 3); });
    }
}

template <bool PROBE>
__device__ void group_apply(const TickArgs &a, int g) {
    const tsim_light_tables &lt = a.lt;
    const tsim_tick_state &s = a.st;
    const int plan = s.g_plan[g];
    if (plan == 0) return;
    auto commit = [&](int c) {
        const int w = *((volatile int32_t *)(s.stopw + c));
        if ((w >> 2) == g + 1) {
            s.stopw[c] = 0;
            s.stop_map[c] = (uint8_t)(w & 1);
        }
    };
    if (plan == 1) {
        for_light_cells(lt, lt.g_all_off, lt.g_all, g, commit);
    } else {
        for_light_cells(lt, lt.g_ns_off, lt.g_ns, g, commit);
        for_light_cells(lt, lt.g_ew_off, lt.g_ew, g, commit);
    }
}

// stop_map as the vehicles of this tick see it: the staged write of a light group that acted this tick, else the map
__device__ __forceinline__ int stop_now(const tsim_tick_state &s, int c) {
    const int w = __ldcg(s.stopw + c);
    return w ? (w & 1) : (int)s.stop_map[c];
}

// claim of the sweep `gen` on cell c: rank of the lowest-ranked vehicle that ends there, NO_CLAIM if none
__device__ __forceinline__ int claim_rank(const u64 *plane, int c, uint32_t gen) {
    const u64 k = __ldcg(plane + c);
    return (uint32_t)(k >> 32) == gen ? (int)(0xffffffffu - (uint32_t)k) : NO_CLAIM;
}
__device__ __forceinline__ void claim_cell(u64 *plane, int c, uint32_t gen, int rank) {
    atomicMax(plane + c, ((u64)gen << 32) | (u64)(0xffffffffu - (uint32_t)rank));
}

}  // namespace tsim
