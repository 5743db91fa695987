// k_tick2.cu -- the vehicle tick around the LIVE vehicles (tsim_tick_state.probe / recs / plans set).
//
// Same tick as k_tick.cu (same reference lines, same claim fixed point: see its header); what changes is what a tick touches:
//   * a compacted LIVE LIST instead of the attempt array.  A vehicle is a 48-byte record that moves to a new slot every
//     tick: survivors of the move phase and the spawns of the tick are appended (one atomic per warp) to the other half of
//     `recs`, so that every phase reads its vehicles as consecutive 48-byte records and a fleet that has mostly arrived (or
//     has mostly not spawned yet) costs what its live vehicles cost;
//   * BIT PLANES over 8 x 8-cell tiles instead of per-cell maps (tick_common.cuh): occupancy, stop, "planned by one vehicle",
//     "planned by several / stop write staged", staged-stop.  A look-ahead of five cells is two 64-bit loads per plane whatever
//     the direction of travel, the planes of an 8192 x 8192 city are 8 MB each and stay in L2, and -- what bounds a large fleet
//     is the rate of L2 atomics (~90 G/s measured), not bytes -- a vehicle makes one atomic per TILE it marks, unmarks or
//     changes the occupancy of, not one per cell;
//   * the claim fixed point (k_tick.cu) only for CONTESTED vehicles: phase A marks every cell a vehicle may end on; a vehicle
//     none of whose cells carries a second mark cannot be blocked by anybody and blocks nobody: it moves as planned.  The others
//     (1-2 % of a dense fleet) go on a list, and only that list is swept;
//   * a per-tick PLAN record (32 bytes: the <= 5 planned cells, activation rank, max_steps, stop bits): the sweeps of the
//     fixed point read it instead of the vehicle state, the tapes and the route arena;
//   * the spawner's claims (lowest attempt index per origin cell) are made during the move phase in the claim plane the
//     final sweep did not use, so a tick has one grid-wide barrier less: decide | sweep x n | move | spawn.
//   * (fleets of >= 200 k vehicles, tsim_tick_state.sort_keys / tile_ws set) the survivors are appended TILE BY TILE: a counting
//     sort by the 64 x 64-cell tile of the new position (one atomic per vehicle for its rank inside the tile, a scan of the tile
//     counts by one CTA, one more pass that moves the 48-byte records), so that the vehicles a warp handles are neighbours on the
//     grid and their probes hit the same cache lines: the fleet is memory-bound on these 32-byte sector gathers (ncu: 2.4 KB of
//     DRAM traffic per vehicle and tick without the sort), which a cell-sorted order turns into cache hits.  One barrier more.
// The public maps (occupancy / stop_map / stuck_map) and the vehicle SoA of tsim_tick_state are written on demand by
// tsim_tick_export: a tick does not touch them (a scattered byte write costs a 32-byte sector read and written back).  What
// stuck_map says about a cell is a property of the vehicle standing on it (move_vehicle city_model.py:1956-1958 sets the cell a
// vehicle ends on, every cell it leaves or crosses is cleared): one bit of its record.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include "tick_common.cuh"

namespace cg = cooperative_groups;

namespace tsim {

enum { S_NLIVE0 = 10, S_NLIVE1 = 11 };   // scalars: entries of the two halves of the live list

struct __align__(16) VRec {   // TSIM_TICK_VREC_BYTES
    long long path_off;   // next cell of the route in ev_cells
    int32_t v;            // spawn-attempt index (row of the tapes)
    int32_t pos;          // cell
    int32_t path_len, steps, stranded, target;
    int16_t stuck_ticks;
    int8_t base_speed, cur_speed;
    int8_t is_stuck, prev_valid, malfunction, direction;
    int8_t collision;     // is_in_collision (sideswipe, vehicle_base.py:534-541); shares `stranded` with the malfunction
    int8_t prev_cur;      // current_speed / stranded flags as they were BEFORE this tick's phase A: what a vehicle earlier in the
    int8_t prev_flags;    //   list sees of one later in the list (bit 0 malfunction, bit 1 collision)
    int8_t mark;          // stuck_map of the cell the vehicle stands on (set by the move that brought it there)
    int32_t pad;
};
static_assert(sizeof(VRec) == TSIM_TICK_VREC_BYTES, "VRec size");

struct __align__(16) VPlan {   // TSIM_TICK_PLAN_BYTES
    int32_t cell[MAX_SPEED];   // the cells the vehicle may enter this tick (first max_steps entries)
    int32_t rank;
    uint8_t m, k, stop, early;   // max_steps; steps granted (phase A: = m, then by the last sweep); stop_now bits of the cells, bit 7 = contested; early exit
    int32_t target;              // an arriving vehicle claims nothing
};
static_assert(sizeof(VPlan) == TSIM_TICK_PLAN_BYTES, "VPlan size");

// the planned cells of a vehicle as (tile word, bit mask) pairs; e[i] = pair of cell i, b[i] = its bit.  Everything is indexed
// by unrolled loops so that it stays in registers.
struct PathBits {
    int w[MAX_SPEED];
    u64 m[MAX_SPEED];
    int e[MAX_SPEED], b[MAX_SPEED];
    int n;
};
__device__ __forceinline__ void path_bits(const TickArgs &a, const int32_t (&cell)[MAX_SPEED], int cnt, PathBits &p) {
    p.n = 0;
#pragma unroll
    for (int i = 0; i < MAX_SPEED; i++) { p.w[i] = -1; p.m[i] = 0; p.e[i] = 0; p.b[i] = 0; }
#pragma unroll
    for (int i = 0; i < MAX_SPEED; i++) {
        if (i >= cnt) continue;
        int w, b;
        cell_wb(a, cell[i], w, b);
        int e = p.n;
#pragma unroll
        for (int q = 0; q < i; q++) if (q < p.n && p.w[q] == w) e = q;
#pragma unroll
        for (int q = 0; q <= i; q++) if (q == e) { p.w[q] = w; p.m[q] |= 1ull << b; }
        if (e == p.n) p.n++;
        p.e[i] = e; p.b[i] = b;
    }
}
// bit of cell i in the words `v` loaded for the pairs of p
__device__ __forceinline__ int path_bit(const PathBits &p, const u64 (&v)[MAX_SPEED], int i) {
    u64 w = 0;
#pragma unroll
    for (int q = 0; q < MAX_SPEED; q++) if (p.e[i] == q) w = v[q];
    return (int)((w >> p.b[i]) & 1ull);
}
__device__ __forceinline__ void path_load(const TickArgs &a, int plane, const PathBits &p, u64 (&v)[MAX_SPEED]) {
    const u64 *pl = bit_plane(a, plane);
#pragma unroll
    for (int q = 0; q < MAX_SPEED; q++) v[q] = q < p.n ? __ldcg(pl + p.w[q]) : 0ull;
}

// stop_map of cell c as the vehicles of this tick see it (the staged write of a light group that acted this tick wins)
__device__ __forceinline__ int stop_seen(const TickArgs &a, int c) {
    int w, b;
    cell_wb(a, c, w, b);
    const u64 stg = __ldcg(bit_plane(a, PL_STG) + w), stp = __ldcg(bit_plane(a, PL_STOP) + w);
    return ((stg >> b) & 1ull) ? (__ldcg(a.st.stopw + c) & 1) : (int)((stp >> b) & 1ull);
}

// phase A of one vehicle (vehicle_base.py:616-663 on the tick-start snapshot): updates the record, fills the plan
__device__ __forceinline__ void decide2(const TickArgs &a, VRec &r, VPlan &pl, int t, int slot) {
    const tsim_tick_state &s = a.st;
    const tsim_tick_tapes &tp = a.tp;
    const int v = r.v;
    const size_t tv = (size_t)t * tp.n_vehicles + v;
    // everything this vehicle may need from the tapes travels together (independent loads)
    const int stamp = s.ev_stamp[v];
    const uint8_t malf = tp.malfunction[tv];
    pl.rank = tp.rank[tv];
    pl.m = 0; pl.k = 0; pl.stop = 0; pl.early = 0; pl.target = r.target;
#pragma unroll
    for (int i = 0; i < MAX_SPEED; i++) pl.cell[i] = -1;
    if (stamp == t) { r.path_off = s.ev_poff[v]; r.path_len = s.ev_plen[v]; }   // the re-plan the reference made this tick (replayed)
    r.prev_cur = r.cur_speed;
    r.prev_flags = (int8_t)((r.malfunction ? 1 : 0) | (r.collision ? 2 : 0));
    if (r.malfunction || r.collision) {   // _tick_stranded :552-565
        if (--r.stranded <= 0) { r.malfunction = 0; r.collision = 0; r.stranded = 0; }
        if (r.malfunction || r.collision) { r.base_speed = 0; r.cur_speed = 0; pl.early = 1; return; }
    }
    if (malf & 1) {   // _check_malfunction :608-610
        r.malfunction = 1; r.collision = 0; r.stranded = MALFUNCTION_TICKS; r.base_speed = 0; r.cur_speed = 0; pl.early = 1;
        return;
    }
    const int pos = r.pos;
    // _check_sideswipe_collision :567-605.  The draw only matters where the tape says it fires (bit 1); whether it is made at all
    // depends on the vehicle next to this one AS IT IS WHEN THIS VEHICLE'S TURN COMES in list order, so a vehicle with a firing
    // draw, a direction and an occupied cell to its left or right is only REGISTERED here (and asks who stands there); the
    // candidates of a tick are settled one after the other once every vehicle has decided (sideswipe_fixup)
    if ((malf & 2) && r.direction >= 0) {
        const int d = r.direction, x = pos % a.W, y = pos / a.W;
        bool any = false;
#pragma unroll
        for (int side = 0; side < 2; side++) {
            const int l = side ? right_of(d) : ((d + 3) & 3), nx = x + dx_of(l), ny = y + dy_of(l);
            if (nx < 0 || nx >= a.W || ny < 0 || ny >= a.H) continue;
            const int c = ny * a.W + nx;
            if (bit_get(a, PL_OCC, c)) { bit_set(a, PL_WANT, c); any = true; }
        }
        if (any) s.sort_keys[atomicAdd(s.scalars + S_NCAND, 1)] = slot;
    }
    if (bit_get(a, PL_STOP, pos)) { r.base_speed = 0; r.cur_speed = 0; pl.early = 1; return; }   // :639-643
    int base = r.base_speed;
    if (base == 0) { base = tp.speed[tv]; r.base_speed = (int8_t)base; }   // :94-112
    int sp = base;
    if (tp.rain_map && tp.rain_map[pos] == 1) sp = max(1, sp - RAIN_REDUCTION);
    r.cur_speed = (int8_t)sp;
    if (sp > MAX_SPEED) s.scalars[S_ERR] = 33;   // tape contract: speeds are random.randint(1, 5) (vehicle_base.py:112)
    // _scan_ahead_for_obstacles :422-452 and _determine_max_steps :719-731: cells at or beyond min(speed, len) cannot lower max_steps
    const int len = r.path_len;
    const int32_t *path = tp.ev_cells + r.path_off;
    int ms = min(min(sp, len), MAX_SPEED);
    int32_t cell[MAX_SPEED];
#pragma unroll
    for (int i = 0; i < MAX_SPEED; i++) cell[i] = i < ms ? path[i] : -1;
#pragma unroll
    for (int i = MAX_SPEED - 1; i >= 0; i--) if (i < ms && cell[i] < 0) ms = i;   // a cell outside the window blocks, like a vehicle
    PathBits pb;
    path_bits(a, cell, ms, pb);
    {
        u64 occ[MAX_SPEED], stp[MAX_SPEED];
        path_load(a, PL_OCC, pb, occ);
        path_load(a, PL_STOP, pb, stp);
#pragma unroll
        for (int q = 0; q < MAX_SPEED; q++) occ[q] |= stp[q];
#pragma unroll
        for (int i = MAX_SPEED - 1; i >= 0; i--) if (i < ms && path_bit(pb, occ, i)) ms = i;
    }
#pragma unroll
    for (int i = 0; i < MAX_SPEED; i++) pl.cell[i] = i < ms ? cell[i] : -1;
    pl.m = (uint8_t)ms;
    pl.k = (uint8_t)ms;   // what an uncontested vehicle does: none of its cells is a stop cell (max_steps ends before the first one)
    // Every cell this vehicle could end on is marked: once = nobody else plans to come here, twice = somebody does.  Only vehicles
    // with a twice-marked cell take part in the claim fixed point (the marks come off again in the move phase).  One atomic per tile.
    {
        u64 mk[MAX_SPEED], was[MAX_SPEED];
#pragma unroll
        for (int q = 0; q < MAX_SPEED; q++) {   // the cells this side of max_steps
            mk[q] = 0;
#pragma unroll
            for (int i = 0; i < MAX_SPEED; i++) if (i < ms && pb.e[i] == q) mk[q] |= 1ull << pb.b[i];
        }
        u64 *t1 = bit_plane(a, PL_T1), *t2 = bit_plane(a, PL_T2);
#pragma unroll
        for (int q = 0; q < MAX_SPEED; q++) was[q] = mk[q] ? atomicOr(t1 + pb.w[q], mk[q]) : 0ull;   // all first (independent round trips)
#pragma unroll
        for (int q = 0; q < MAX_SPEED; q++) if (was[q] & mk[q]) atomicOr(t2 + pb.w[q], was[q] & mk[q]);
    }
    if (ms <= 0) {
        r.base_speed = 0;
        if (pos == r.target) s.scalars[S_ERR] = 30;   // tape contract: a live vehicle is never at its target in phase A
        pl.early = 1;
    }
}

// The candidates of a tick, ONE thread, in list order (= ascending spawn-attempt index, the order run_parallel_decide visits
// the vehicles with one worker): the vehicle to the left, then the one to the right; the first that is moving -- as far as its
// last step_decide says: this tick's if it comes earlier in the list, last tick's if later -- the opposite way takes the draw,
// which fires (that is why the vehicle is a candidate): _set_collision for both (vehicle_base.py:534-541, 600 ticks).  The
// partner, if it comes LATER in the list, meets its own step_decide already stranded (599 left, early exit); if it came earlier
// its plan for this tick stands.  Rare by construction (the reference's chance is 1e-9 per draw); a few loads per candidate.
__device__ void sideswipe_fixup(const TickArgs &a, VRec *rc, VPlan *plans, const u64 *who, uint32_t gen0, int n_cand) {
    const tsim_tick_state &s = a.st;
    int32_t *cand = s.sort_keys;
    for (int i = 1; i < n_cand; i++) {   // insertion sort by vehicle index
        const int c = cand[i], cv = rc[c].v;
        int j = i - 1;
        while (j >= 0 && rc[cand[j]].v > cv) { cand[j + 1] = cand[j]; j--; }
        cand[j + 1] = c;
    }
    for (int q = 0; q < n_cand; q++) {
        const int i = cand[q];
        VRec &R = rc[i];
        const int d = R.direction, x = R.pos % a.W, y = R.pos / a.W;
        bool settled = R.collision || R.malfunction;   // hit by an earlier candidate of this tick: its own turn finds it stranded
        for (int side = 0; side < 2; side++) {
            const int l = side ? right_of(d) : ((d + 3) & 3), nx = x + dx_of(l), ny = y + dy_of(l);
            if (nx < 0 || nx >= a.W || ny < 0 || ny >= a.H) continue;
            const int c = ny * a.W + nx;
            if (settled || !bit_get(a, PL_OCC, c)) continue;
            const u64 w = who[c];
            if ((uint32_t)(w >> 32) != gen0) { s.scalars[S_ERR] = 35; continue; }   // an occupied cell without a live vehicle on it
            const int j = (int)(uint32_t)w;
            VRec &U = rc[j];
            const bool earlier = U.v < R.v;
            const int u_cur = earlier ? U.cur_speed : U.prev_cur;
            const bool u_stranded = earlier ? (U.malfunction || U.collision) : (U.prev_flags != 0);
            if (u_cur <= 0 || U.is_stuck || u_stranded) continue;
            if (U.direction != opp_of(d)) continue;
            R.collision = 1; R.malfunction = 0; R.stranded = COLLISION_TICKS; R.base_speed = 0; R.cur_speed = 0;
            plans[i].k = 0; plans[i].early = 1;   // m stays: the marks on its planned cells still come off in the move phase
            U.collision = 1; U.malfunction = 0; U.base_speed = 0; U.cur_speed = 0; U.prev_cur = 0; U.prev_flags = 2;
            if (earlier) U.stranded = COLLISION_TICKS;                                   // decided before the hit: it still makes this tick's move
            else { U.stranded = COLLISION_TICKS - 1; plans[j].k = 0; plans[j].early = 1; }   // _tick_stranded at its own turn
            settled = true;
        }
    }
    for (int q = 0; q < n_cand; q++) {   // the questions are answered: take the marks off again (a cell may have been asked about twice)
        const VRec &R = rc[cand[q]];
        const int d = R.direction, x = R.pos % a.W, y = R.pos / a.W;
        for (int side = 0; side < 2; side++) {
            const int l = side ? right_of(d) : ((d + 3) & 3), nx = x + dx_of(l), ny = y + dy_of(l);
            if (nx < 0 || nx >= a.W || ny < 0 || ny >= a.H) continue;
            if (bit_get(a, PL_WANT, ny * a.W + nx)) bit_clear(a, PL_WANT, ny * a.W + nx);
        }
    }
    s.scalars[S_NCAND] = 0;
}

// The light cells of those of a warp's 32 groups (lane L: group g0 + L, plan != 0: it acts this tick) as ONE list the lanes stride
// over, four items per lane at a time so that their loads travel together: f(groups[4], plans[4], indices into the flat cell list
// [4], -1 = none).  All 32 lanes call.  A group that acts writes a few dozen cells, and idle intersections all switch in the same
// tick (every MIN_GREEN ticks): taking the groups, or the cells, one after the other is a long chain of memory round trips that
// the whole grid waits for at the next barrier.
constexpr int GI = 4;
template <class F>
__device__ __forceinline__ void warp_group_cells(const TickArgs &a, int g0, int ng, int lane, int plan, F f) {
    constexpr uint32_t FULL = 0xffffffffu;
    int b = 0, n = 0;
    if (g0 + lane < ng) { b = a.gc_off[g0 + lane]; n = a.gc_off[g0 + lane + 1] - b; }   // (asked for before the plan is known to matter)
    if (!__ballot_sync(FULL, plan != 0)) return;
    if (plan == 0) n = 0;
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
    const int total = __shfl_sync(FULL, incl, 31);
    for (int i0 = 0; i0 < total; i0 += 32 * GI) {
        int gg[GI], pl[GI], k[GI];
#pragma unroll
        for (int u = 0; u < GI; u++) {
            const int i = i0 + u * 32 + lane;
            int lo = 0;   // the lane whose group owns item i: the first one with incl > i
#pragma unroll
            for (int step = 16; step; step >>= 1) { const int v = __shfl_sync(FULL, incl, min(lo + step - 1, 31)); if (v <= i) lo += step; }
            lo = min(lo, 31);
            const int start = __shfl_sync(FULL, incl - n, lo), ob = __shfl_sync(FULL, b, lo);
            pl[u] = __shfl_sync(FULL, plan, lo);
            gg[u] = g0 + lo;
            k[u] = i < total ? ob + (i - start) : -1;
        }
        f(gg, pl, k);
    }
}

// where a tick's time goes: nanoseconds (globaltimer) between the grid-wide barriers, summed over the ticks since the last reset
// by one thread: [0] decide, [1] sideswipes, [2] sweep 0, [3] later sweeps, [4] move + append, [5] sorted append, [6] spawns + lights
__device__ unsigned long long g_tick_phase_ns[16];   // [8..]: thread 0's own share: decide vehicles, decide groups, phase-4 record moves, spawns, group commits, events
__device__ __forceinline__ unsigned long long timer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define OWN_DONE(slot) do { if (tid == 0) { const unsigned long long now_ = timer_ns(); g_tick_phase_ns[slot] += now_ - t_own; t_own = now_; } } while (0)
#define PHASE_DONE(slot) do { if (tid == 0) { const unsigned long long now_ = timer_ns(); g_tick_phase_ns[slot] += now_ - t_phase; t_phase = now_; } } while (0)

template <int MINB>
__global__ void __launch_bounds__(256, MINB) tick2_kernel(TickArgs a) {
    cg::grid_group grid = cg::this_grid();
    constexpr uint32_t FULL = 0xffffffffu;
    const tsim_tick_state &s = a.st;
    const tsim_tick_tapes &tp = a.tp;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x, lane = threadIdx.x & 31;
    const int nv = tp.n_vehicles, ng = a.lt.n_groups;
    const int gwarp = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;   // warps in an order that deals a short list out over all SMs (the light groups)
    const size_t ncell = (size_t)a.W * a.H;
    u64 *plane[2] = {(u64 *)s.claim, (u64 *)s.claim + ncell};
    VRec *recs[2] = {(VRec *)s.recs, (VRec *)s.recs + nv};
    VRec *tmp = (VRec *)s.recs + 2 * (size_t)nv;   // sorted append only: the survivors before they move to their tile's slots
    int32_t *tile_cnt = s.tile_ws, *tile_base = s.tile_ws ? s.tile_ws + a.n_tiles : nullptr;
    VPlan *plans = (VPlan *)s.plans;

    auto scatter_events = [&](int t) {   // route events of tick t: the vehicle picks its own up in phase A (or when it spawns)
        if (t >= tp.n_ticks) return;
        for (int e = tp.ev_first[t] + tid; e < tp.ev_first[t + 1]; e += nth) {
            const int v = tp.ev_vehicle[e];
            s.ev_poff[v] = tp.ev_off[e];
            s.ev_plen[v] = (int32_t)(tp.ev_off[e + 1] - tp.ev_off[e]);
            s.ev_stamp[v] = t;
        }
    };
    scatter_events(*((volatile int32_t *)(s.scalars + S_TICK)));
    grid.sync();
    unsigned long long t_phase = timer_ns(), t_own = t_phase;
    for (int it = 0; it < a.n_ticks; it++) {
        const int t = *((volatile int32_t *)(s.scalars + S_TICK));
        if (t >= tp.n_ticks) { if (tid == 0) s.scalars[S_ERR] = 31; break; }
        const int cur = t & 1, nxt = cur ^ 1;
        VRec *rc = recs[cur], *rn = recs[nxt];
        const int n_live = *((volatile int32_t *)(s.scalars + S_NLIVE0 + cur));
        const uint32_t gen0 = (uint32_t)t * GEN_PER_TICK + 1u;   // generation nobody writes: "no claims yet"
        // vehicles move at most 5 cells a tick and the plain append keeps the vehicles of a warp together: the tile order of the
        // list decays slowly, so the list is only re-sorted every few ticks (two barriers and one more pass over the records)
        const bool sorted = a.n_tiles > 0 && t % a.sort_every == 0;
        if (a.algo == 3 && ng > 0) green_wave_prepass<true>(a, grid, tid, nth);
        // ---- 1: phase A of every live vehicle + light-group decisions (staged)
        int live = 0;
        for (int i = tid; i < n_live; i += nth) {
            VRec r = rc[i];
            VPlan pl;
            decide2(a, r, pl, t, i);
            rc[i] = r;
            plans[i] = pl;
            live += r.pos >= a.own_lo && r.pos < a.own_hi;
        }
        if (tid == 0) t_own = t_phase;
        OWN_DONE(8);
        live = __reduce_add_sync(FULL, live);
        if (lane == 0 && live) atomicAdd((unsigned long long *)(s.scalars + S_UPD_HI), (unsigned long long)live);
        // light groups: one thread per group runs the controller; the groups of a warp that act are then staged by the whole warp,
        // lanes striding over the flat list of the cells their lights control (a few dozen atomics each; one thread doing them
        // one after the other would keep the whole grid waiting at the barrier)
        for (int g0 = gwarp * 32; g0 < ng; g0 += nth) {
            const int g = g0 + lane;
            const int plan = g < ng ? group_controller<true>(a, g) : 0;
            warp_group_cells(a, g0, ng, lane, plan, [&](const int (&gg)[GI], const int (&pl)[GI], const int (&k)[GI]) {
                int role[GI], c[GI];
#pragma unroll
                for (int u = 0; u < GI; u++) { role[u] = k[u] >= 0 ? a.gc_role[k[u]] : 0; c[u] = k[u] >= 0 ? a.gc_cell[k[u]] : -1; }
#pragma unroll
                for (int u = 0; u < GI; u++) {
                    // all red: the lights of g_all stop; phase p goes: its axis goes (0), the other one stops (3)
                    const int val = pl[u] == 1 ? (role[u] == 0 ? 1 : -1) : (role[u] == 0 ? -1 : ((role[u] == 1) == (pl[u] == 2) ? 0 : 3));
                    if (c[u] >= 0 && val >= 0) {   // whoever plans to enter the cell must look at the staged value: second mark
                        int w, b;
                        cell_wb(a, c[u], w, b);
                        atomicMax(s.stopw + c[u], (gg[u] + 1) * 4 + val);
                        atomicOr(bit_plane(a, PL_STG) + w, 1ull << b);
                        atomicOr(bit_plane(a, PL_T2) + w, 1ull << b);
                    }
                }
            });
        }
        OWN_DONE(9);
        if (tid == 0) s.scalars[S_NLIVE0 + nxt] = 0;   // the other half was last read as `cur` one tick ago
        grid.sync();
        PHASE_DONE(0);
        // ---- 1b: sideswipes (only on a tick with a registered candidate: two more barriers)
        const int n_cand = *((volatile int32_t *)(s.scalars + S_NCAND));
        if (n_cand > 0) {
            // who stands on the cells the candidates asked about: slot of that vehicle, tagged with a generation nobody reads as a claim
            for (int i = tid; i < n_live; i += nth) {
                const int p = rc[i].pos;
                if (bit_get(a, PL_WANT, p)) plane[1][p] = ((u64)gen0 << 32) | (u64)(uint32_t)i;
            }
            grid.sync();
            if (tid == 0) sideswipe_fixup(a, rc, plans, plane[1], gen0, n_cand);
            grid.sync();
            PHASE_DONE(1);
        }
        // ---- 2: claim fixed point, one barrier per sweep.  Sweep 0 visits every vehicle: the staged stop_map writes of this tick's
        // light groups are complete now and are folded into the plan; a vehicle none of whose planned cells is marked twice is
        // DONE (nobody can claim a cell it may enter, nobody cares where it ends); the others make their first claim and go on the
        // contested list, which is all the later sweeps read.
        int last = 0;
        int32_t *cont = s.sort_keys;   // free between the sideswipe candidates of phase 1 and the sort keys of phase 3
        for (int iter = 0;; iter++) {
            const int fidx[3] = {S_FLAG0, S_FLAG1, S_FLAG2};   // three flags in rotation: the one zeroed after sweep i is written in sweep i + 2
            int32_t *flag = s.scalars + fidx[iter % 3];
            const int pp = (iter + 1) & 1, pc = iter & 1;
            const uint32_t gen_prev = gen0 + iter, gen_cur = gen0 + iter + 1;
            bool ch = false;
            if (iter == 0) {
                for (int i0 = tid - lane; i0 < n_live; i0 += nth) {   // whole warps: one atomic per warp for the list
                    const int i = i0 + lane;
                    bool contested = false;
                    if (i < n_live) {
                        VPlan pl = plans[i];
                        if (!pl.early && pl.m) {
                            const int m = pl.m;
                            PathBits pb;
                            path_bits(a, pl.cell, m, pb);
                            u64 t2[MAX_SPEED];
                            path_load(a, PL_T2, pb, t2);
#pragma unroll
                            for (int q = 0; q < MAX_SPEED; q++) contested |= (t2[q] & pb.m[q]) != 0;
                            if (contested) {   // the stop bits as of now (staged writes folded in), the first claim
                                uint32_t sm = 0;
#pragma unroll
                                for (int j = 0; j < MAX_SPEED; j++) if (j < m && stop_seen(a, pl.cell[j]) == 1) sm |= 1u << j;
                                int k = 0;
                                bool open = true;
#pragma unroll
                                for (int j = 0; j < MAX_SPEED; j++) {   // _execute_movement :733-753: a stop cell may only be entered on the last step
                                    open = open && j < m && !(((sm >> j) & 1u) && j + 1 != m);
                                    if (open) k = j + 1;
                                }
                                pl.stop = (uint8_t)(sm | 0x80u);
                                pl.k = (uint8_t)k;
                                *reinterpret_cast<uint32_t *>(&plans[i].m) = *reinterpret_cast<const uint32_t *>(&pl.m);   // m, k, stop, early
                                if (k >= 1) {
                                    const int c = pl.cell[k - 1];
                                    if (c != pl.target) claim_cell(plane[pc], c, gen_cur, pl.rank);   // an arriving vehicle is removed at once
                                }
                            }
                        }
                    }
                    const uint32_t mask = __ballot_sync(FULL, contested);
                    if (mask) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(s.scalars + S_NCONT, __popc(mask));
                        base = __shfl_sync(FULL, base, 0);
                        if (contested) cont[base + __popc(mask & ((1u << lane) - 1u))] = i;
                        ch = true;
                    }
                }
            } else {
                const int n_cont = *((volatile int32_t *)(s.scalars + S_NCONT));
                for (int q = tid; q < n_cont; q += nth) {
                    const int i = cont[q];
                    const VPlan pl = plans[i];
                    const int m = pl.m;
                    int cr[MAX_SPEED];
#pragma unroll
                    for (int j = 0; j < MAX_SPEED; j++) cr[j] = j < m ? claim_rank(plane[pp], pl.cell[j], gen_prev) : NO_CLAIM;
                    int k = 0;
                    bool open = true;
#pragma unroll
                    for (int j = 0; j < MAX_SPEED; j++) {   // _execute_movement :733-753
                        // a lower-ranked vehicle ends here; a stop cell may only be entered on the last step
                        open = open && j < m && !(cr[j] < pl.rank) && !(((pl.stop >> j) & 1u) && j + 1 != m);
                        if (open) k = j + 1;
                    }
                    if (k != pl.k) {
                        ch = true;
                        plans[i].k = (uint8_t)k;
                    }
                    if (k >= 1) {
                        const int c = pl.cell[k - 1];
                        if (c != pl.target) claim_cell(plane[pc], c, gen_cur, pl.rank);
                    }
                }
            }
            if (__any_sync(FULL, ch) && lane == 0) *flag = 1;
            grid.sync();
            PHASE_DONE(iter == 0 ? 2 : 3);
            if (iter == 0 && tid == 0) g_tick_phase_ns[7] += (unsigned long long)*((volatile int32_t *)(s.scalars + S_NCONT));
            const int any = *((volatile int32_t *)flag);
            if (tid == 0) { s.scalars[fidx[(iter + 2) % 3]] = 0; s.scalars[S_ITERS]++; }
            last = iter;
            if (!any) break;
            if (iter >= GEN_PER_TICK - 4) { if (tid == 0) s.scalars[S_ERR] = 32; break; }
        }
        const int pf = last & 1, po = pf ^ 1;   // plane of the final sweep; the other one takes the spawner's claims
        const uint32_t gen_spawn = (uint32_t)t * GEN_PER_TICK + GEN_PER_TICK - 1;
        // ---- 3: apply the moves, append the survivors to the other half of the list; spawner claims
        const int k0 = tp.spawn_first[t], k1 = tp.spawn_first[t + 1];
        for (int i0 = tid - lane; i0 < n_live; i0 += nth) {   // whole warps: the append is one atomic per warp
            const int i = i0 + lane;
            bool keep = false;
            VRec r;
            if (i < n_live) {
                r = rc[i];
                const VPlan pl = plans[i];
                int pos = r.pos;
                const int target = r.target;
                if (pl.m) {   // nobody reads the marks of this tick any more: they come off, one atomic per tile
                    PathBits pb;
                    path_bits(a, pl.cell, pl.m, pb);
                    u64 *t1 = bit_plane(a, PL_T1), *t2 = bit_plane(a, PL_T2);
#pragma unroll
                    for (int q = 0; q < MAX_SPEED; q++) {
                        if (q >= pb.n) continue;
                        atomicAnd(t1 + pb.w[q], ~pb.m[q]);
                        if (pl.stop & 0x80u) atomicAnd(t2 + pb.w[q], ~pb.m[q]);
                    }
                }
                if (!pl.early) {
                    const int k = pl.k;
                    if (k >= 1) {
                        const int fin = pl.cell[k - 1], prev = k >= 2 ? pl.cell[k - 2] : pos;
                        // move_vehicle city_model.py:1945-1963.  Nobody ends on the cell this vehicle leaves (it was occupied when
                        // everybody looked ahead), so its occupancy bit simply goes; what stuck_map says about the new cell
                        // travels in the record (:1956-1958, before _move_to resets is_stuck)
                        {
                            int w0, b0, w1, b1;
                            cell_wb(a, pos, w0, b0);
                            cell_wb(a, fin, w1, b1);
                            u64 *occ = bit_plane(a, PL_OCC);
                            if (fin == target) atomicAnd(occ + w0, ~(1ull << b0));
                            else if (w0 == w1) atomicXor(occ + w0, (1ull << b0) | (1ull << b1));   // both bits are known: one flips off, one on
                            else { atomicAnd(occ + w0, ~(1ull << b0)); atomicOr(occ + w1, 1ull << b1); }
                        }
                        r.mark = (k == 1 && r.is_stuck) ? 1 : 0;
                        const int d = fin - prev;                                // compute_direction numba_utilities.py:14-28
                        r.direction = (int8_t)(d == a.W ? DN : d == 1 ? DE : d == -a.W ? DS : d == -1 ? DW : r.direction);
                        if (r.stuck_ticks > 0) { r.is_stuck = 0; r.stuck_ticks = 0; }   // _move_to :528-532
                        r.steps += k;
                        r.path_off += k; r.path_len -= k;
                        r.pos = pos = fin;
                    }
                    r.prev_valid = 1;   // step() :677
                } else {                // :679-680 tick_stuck :687-693
                    if (r.prev_valid && stop_seen(a, pos) != 1) {
                        const int st = ++r.stuck_ticks;
                        if (st > STUCK_THRESHOLD && !r.is_stuck) r.is_stuck = 1;
                    }
                    if (pos == target) bit_clear(a, PL_OCC, pos);
                }
                keep = pos != target;   // on_target_reached :755-775 -> remove_vehicle city_model.py:1920-1929
            }
            if (sorted) {   // rank inside the tile of the new position now, the slot once every tile's count is known
                if (i < n_live) {
                    int tile = -1, rk = 0;
                    if (keep) {
                        tile = ((r.pos / a.W) >> a.tile_sy) * a.tiles_x + ((r.pos % a.W) >> a.tile_sx);
                        rk = atomicAdd(tile_cnt + tile, 1);
                        tmp[i] = r;
                    }
                    s.sort_keys[2 * i] = tile; s.sort_keys[2 * i + 1] = rk;
                }
                continue;
            }
            const uint32_t mask = __ballot_sync(FULL, keep);
            if (mask) {
                int base = 0;
                if (lane == 0) base = atomicAdd(s.scalars + S_NLIVE0 + nxt, __popc(mask));
                base = __shfl_sync(FULL, base, 0);
                if (keep) rn[base + __popc(mask & ((1u << lane) - 1u))] = r;
            }
        }
        for (int k = k0 + tid; k < k1; k += nth)   // lowest attempt index per origin cell; whether the cell is free is known after the barrier
            if (tp.origin[k] >= 0) claim_cell(plane[po], tp.origin[k], gen_spawn, k);   // read back without asking the probe byte
        if (tid == 0) { s.scalars[S_FLAG0] = 0; s.scalars[S_FLAG1] = 0; s.scalars[S_FLAG2] = 0; }   // nobody touches the sweep flags here
        grid.sync();
        PHASE_DONE(4);
        if (sorted) {
            // ---- 3b: first slot of every tile = exclusive scan of the tile counts (one CTA; the counts are zeroed for the next tick)
            if (blockIdx.x == 0) {
                __shared__ int s_part[256];
                const int per = (a.n_tiles + 255) / 256, t0 = threadIdx.x * per, t1 = min(t0 + per, a.n_tiles);
                int sum = 0;
                for (int q = t0; q < t1; q++) sum += tile_cnt[q];
                s_part[threadIdx.x] = sum;
                __syncthreads();
                if (threadIdx.x < 32) {   // 256 partial sums: 8 per lane
                    int loc[8], tot = 0;
#pragma unroll
                    for (int q = 0; q < 8; q++) { loc[q] = s_part[threadIdx.x * 8 + q]; tot += loc[q]; }
                    int incl = tot;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
                    int run = incl - tot;
#pragma unroll
                    for (int q = 0; q < 8; q++) { s_part[threadIdx.x * 8 + q] = run; run += loc[q]; }
                    if (lane == 31) s.scalars[S_NLIVE0 + nxt] = incl;   // the spawns of this tick are appended behind the survivors
                }
                __syncthreads();
                int run = s_part[threadIdx.x];
                for (int q = t0; q < t1; q++) { const int c = tile_cnt[q]; tile_base[q] = run; run += c; tile_cnt[q] = 0; }
            }
            grid.sync();
            PHASE_DONE(5);
            // ---- 3c: the survivors move to their tile's slots
            for (int i = tid; i < n_live; i += nth) {
                const int tile = s.sort_keys[2 * i];
                if (tile >= 0) rn[tile_base[tile] + s.sort_keys[2 * i + 1]] = tmp[i];
            }
        }
        if (tid == 0) t_own = timer_ns();
        // ---- 4: spawns (appended like the survivors), commit of the staged stop_map writes, the next tick's route events
        for (int k0w = k0 + tid - lane; k0w < k1; k0w += nth) {
            const int k = k0w + lane;
            bool born = false;
            int o = -1;
            if (k < k1) {
                o = tp.origin[k];
                born = o >= 0 && !bit_get(a, PL_OCC, o) && claim_rank(plane[po], o, gen_spawn) == k;
            }
            const uint32_t mask = __ballot_sync(FULL, born);
            if (mask) {
                int base = 0;
                if (lane == 0) base = atomicAdd(s.scalars + S_NLIVE0 + nxt, __popc(mask));
                base = __shfl_sync(FULL, base, 0);
                if (born) {
                    VRec r;
                    r.v = k; r.pos = o; r.target = tp.target[k];
                    const bool ev = s.ev_stamp[k] == t;   // the route planned at spawn time
                    r.path_off = ev ? s.ev_poff[k] : 0; r.path_len = ev ? s.ev_plen[k] : 0;
                    r.steps = 0; r.stranded = 0; r.stuck_ticks = 0; r.base_speed = 0; r.cur_speed = 0;
                    r.is_stuck = 0; r.prev_valid = 0; r.malfunction = 0; r.direction = -1; r.collision = 0; r.prev_cur = 0; r.prev_flags = 0; r.mark = 0; r.pad = 0;
                    rn[base + __popc(mask & ((1u << lane) - 1u))] = r;
                    bit_set(a, PL_OCC, o);   // place_vehicle city_model.py:1904-1907
                }
            }
        }
        OWN_DONE(11);
        for (int g0 = gwarp * 32; g0 < ng; g0 += nth) {   // commit of the staged writes, the same way
            const int g = g0 + lane;
            const int plan = g < ng ? s.g_plan[g] : 0;
            warp_group_cells(a, g0, ng, lane, plan, [&](const int (&gg)[GI], const int (&pl)[GI], const int (&k)[GI]) {
                int c[GI], w[GI];
#pragma unroll
                for (int u = 0; u < GI; u++) c[u] = (k[u] >= 0 && (a.gc_role[k[u]] == 0) == (pl[u] == 1)) ? a.gc_cell[k[u]] : -1;
#pragma unroll
                for (int u = 0; u < GI; u++) w[u] = c[u] >= 0 ? __ldcg(s.stopw + c[u]) : 0;
#pragma unroll
                for (int u = 0; u < GI; u++)
                    if (c[u] >= 0 && (w[u] >> 2) == gg[u] + 1) {   // this group's write won the cell (written twice if the cell is listed twice: same values)
                        s.stopw[c[u]] = 0;
                        int tw, tb;
                        cell_wb(a, c[u], tw, tb);
                        if (w[u] & 1) atomicOr(bit_plane(a, PL_STOP) + tw, 1ull << tb); else atomicAnd(bit_plane(a, PL_STOP) + tw, ~(1ull << tb));
                        atomicAnd(bit_plane(a, PL_STG) + tw, ~(1ull << tb));
                        atomicAnd(bit_plane(a, PL_T2) + tw, ~(1ull << tb));
                    }
            });
        }
        OWN_DONE(12);
        scatter_events(t + 1);
        OWN_DONE(13);
        if (tid == 0) { s.scalars[S_TICK] = t + 1; s.scalars[S_NCONT] = 0; }
        grid.sync();
        PHASE_DONE(6);
    }
}

// live records -> the vehicle SoA of tsim_tick_state (the caller zeroed `alive`)
__global__ void __launch_bounds__(256) tick2_export_kernel(tsim_tick_state s, int nv) {
    const int cur = s.scalars[S_TICK] & 1;
    const int n = s.scalars[S_NLIVE0 + cur];
    const VRec *rc = (const VRec *)s.recs + (size_t)cur * nv;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const VRec r = rc[i];
        const int v = r.v;
        s.alive[v] = 1; s.pos[v] = r.pos; s.path_off[v] = r.path_off; s.path_len[v] = r.path_len; s.steps[v] = r.steps; s.stranded[v] = r.stranded;
        s.stuck_ticks[v] = r.stuck_ticks; s.base_speed[v] = r.base_speed; s.cur_speed[v] = r.cur_speed; s.is_stuck[v] = r.is_stuck;
        s.prev_valid[v] = r.prev_valid; s.malfunction[v] = (int8_t)((r.malfunction ? 1 : 0) | (r.collision ? 2 : 0)); s.direction[v] = r.direction;
        s.occupancy[r.pos] = 1; s.stuck_map[r.pos] = (uint8_t)r.mark;   // the caller zeroed both maps
    }
}

// stop_map = the committed stop plane, cell by cell
__global__ void __launch_bounds__(256) tick2_export_stop_kernel(int W, int H, int tiles_x, const unsigned long long *stop, uint8_t *stop_map) {
    const long long n = (long long)W * H;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(c / W), x = (int)(c - (long long)y * W);
        stop_map[c] = (uint8_t)((stop[(y >> 3) * tiles_x + (x >> 3)] >> ((y & 7) * 8 + (x & 7))) & 1ull);
    }
}

// tsim_tick_state.group_ws: | header int64[8] = entries of the N-S lane, W-E lane and cluster lists, occupancy words, flat light
// cells | masks u64[entries] | tiles int32[entries] | counts int32[3][n_groups] | flat offsets int32[n_groups
// + 1] | flat cells int32[flat] | flat roles u8[flat] |  (the occupancy words are plane PL_OCC of `probe`).  The (tile, mask) pairs of a group's list start where its cells start in
// the CSR table of tsim_light_tables (W-E lanes and clusters shifted by the lists before them).
struct GroupWs {
    long long n_ns, n_ew, n_cl, n_occ, n_flat;
    unsigned long long *mask;
    int32_t *tile, *cnt, *gc_off, *gc_cell;
    uint8_t *gc_role;
    size_t bytes;
};
static GroupWs group_ws_layout(void *base, long long n_ns, long long n_ew, long long n_cl, long long n_occ, long long n_flat, int ng) {
    GroupWs w{n_ns, n_ew, n_cl, n_occ, n_flat, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    const long long e = n_ns + n_ew + n_cl;
    char *p = (char *)base + 64;
    w.mask = (unsigned long long *)p; p += e * 8;
    w.tile = (int32_t *)p; p += e * 4;
    w.cnt = (int32_t *)p; p += (long long)3 * ng * 4;
    w.gc_off = (int32_t *)p; p += ((long long)ng + 1) * 4;
    w.gc_cell = (int32_t *)p; p += n_flat * 4;
    w.gc_role = (uint8_t *)p; p += n_flat;
    w.bytes = (size_t)(p - (char *)base);
    return w;
}

// cells controlled by the lights of list `off/lights` of group g
__device__ __forceinline__ int light_list_cells(const tsim_light_tables &lt, const int32_t *off, const int32_t *lights, int g) {
    int n = 0;
    for (int k = off[g]; k < off[g + 1]; k++) n += lt.tl_off[lights[k] + 1] - lt.tl_off[lights[k]];
    return n;
}
// out[g + 1] = flat entries of group g (total != NULL: only their sum)
__global__ void __launch_bounds__(256) group_flat_count_kernel(tsim_light_tables lt, int32_t *out, unsigned long long *total) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= lt.n_groups) return;
    const int n = light_list_cells(lt, lt.g_all_off, lt.g_all, g) + light_list_cells(lt, lt.g_ns_off, lt.g_ns, g) + light_list_cells(lt, lt.g_ew_off, lt.g_ew, g);
    if (total) atomicAdd(total, (unsigned long long)n); else out[g + 1] = n;
}
// in place: counts in off[1..n] -> offsets (one CTA; init time)
__global__ void __launch_bounds__(1024) group_flat_scan_kernel(int n, int32_t *off) {
    __shared__ int part[1024];
    const int per = (n + 1023) / 1024, b = 1 + threadIdx.x * per, e = min(b + per, n + 1);
    int sum = 0;
    for (int i = b; i < e; i++) sum += off[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) { int run = 0; for (int i = 0; i < 1024; i++) { const int v = part[i]; part[i] = run; run += v; } off[0] = 0; }
    __syncthreads();
    int run = part[threadIdx.x];
    for (int i = b; i < e; i++) { run += off[i]; off[i] = run; }
}
__global__ void __launch_bounds__(256) group_flat_fill_kernel(tsim_light_tables lt, const int32_t *off, int32_t *cell, uint8_t *role) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= lt.n_groups) return;
    int o = off[g];
    for (int r = 0; r < 3; r++) {
        const int32_t *lo = r == 0 ? lt.g_all_off : r == 1 ? lt.g_ns_off : lt.g_ew_off, *ll = r == 0 ? lt.g_all : r == 1 ? lt.g_ns : lt.g_ew;
        for (int k = lo[g]; k < lo[g + 1]; k++)
            for (int q = lt.tl_off[ll[k]]; q < lt.tl_off[ll[k] + 1]; q++) { cell[o] = lt.tl_cells[q]; role[o] = (uint8_t)r; o++; }
    }
}

__global__ void __launch_bounds__(256) group_masks_kernel(int ng, int W, int tiles_x, const int32_t *off0, const int32_t *cells0, const int32_t *off1,
                                                          const int32_t *cells1, const int32_t *off2, const int32_t *cells2, int base1, int base2,
                                                          unsigned long long *mask, int32_t *tile, int32_t *cnt) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= 3 * ng) return;
    const int kind = id / ng, g = id - kind * ng;
    const int32_t *off = kind == 0 ? off0 : kind == 1 ? off1 : off2, *cells = kind == 0 ? cells0 : kind == 1 ? cells1 : cells2;
    const int base = (kind == 0 ? 0 : kind == 1 ? base1 : base2) + off[g];
    int n = 0;
    for (int k = off[g]; k < off[g + 1]; k++) {
        const int c = cells[k];
        if (c < 0) continue;
        const int y = c / W, x = c - y * W, w = (y >> 3) * tiles_x + (x >> 3);
        const unsigned long long bit = 1ull << ((y & 7) * 8 + (x & 7));
        int i = 0;
        for (; i < n; i++) if (tile[base + i] == w && !(mask[base + i] & bit)) break;   // a cell listed twice opens an entry of its own
        if (i == n) { tile[base + n] = w; mask[base + n] = bit; n++; } else mask[base + i] |= bit;
    }
    cnt[kind * ng + g] = n;
}

__global__ void fill_i32_kernel2(long long n, int32_t *p, int32_t v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace tsim

using namespace tsim;

// tiles of the sorted append: 64 x 64 cells, doubled (the smaller side first) until there are at most 32768 of them
static void tick_tiles(const tsim_cfg *cfg, int &sx, int &sy, int &tiles_x, int &n_tiles) {
    sx = 6; sy = 6;
    for (;;) {
        tiles_x = (cfg->width + (1 << sx) - 1) >> sx;
        const long long n = (long long)tiles_x * ((cfg->win_rows + (1 << sy) - 1) >> sy);
        if (n <= 32768) { n_tiles = (int)n; return; }
        if (sx <= sy) sx++; else sy++;
    }
}

bool tick2_enabled(const tsim_tick_state *st) { return st->probe != nullptr; }

static size_t probe_bytes(const tsim_cfg *cfg) { return (size_t)N_PLANES * ((cfg->width + 7) / 8) * ((cfg->win_rows + 7) / 8) * 8; }

extern "C" tsim_status tsim_tick_probe_bytes(const tsim_cfg *cfg, long long *bytes) {
    tsim_status r = check_cfg(cfg);
    if (r != TSIM_OK) return r;
    if (!bytes) { set_error("tsim_tick_probe_bytes: NULL output"); return TSIM_ERR_CONFIG; }
    *bytes = (long long)probe_bytes(cfg);
    return TSIM_OK;
}

static tsim_status group_list_sizes(const tsim_light_tables *lt, long long *n_ns, long long *n_ew, long long *n_cl, long long *n_flat) {
    int32_t v[3] = {0, 0, 0};
    unsigned long long flat = 0;
    if (lt->n_groups > 0) {
        if (!lt->g_nsin_off || !lt->g_ewin_off || !lt->g_cl_off || !lt->tl_off || !lt->g_all_off || !lt->g_ns_off || !lt->g_ew_off) {
            set_error("tick: NULL light-group table");
            return TSIM_ERR_CONFIG;
        }
        TSIM_CUDA(cudaMemcpy(&v[0], lt->g_nsin_off + lt->n_groups, 4, cudaMemcpyDeviceToHost));
        TSIM_CUDA(cudaMemcpy(&v[1], lt->g_ewin_off + lt->n_groups, 4, cudaMemcpyDeviceToHost));
        TSIM_CUDA(cudaMemcpy(&v[2], lt->g_cl_off + lt->n_groups, 4, cudaMemcpyDeviceToHost));
        unsigned long long *d = nullptr;
        TSIM_CUDA(cudaMalloc(&d, 8));
        TSIM_CUDA(cudaMemset(d, 0, 8));
        group_flat_count_kernel<<<div_up(lt->n_groups, 256), 256>>>(*lt, nullptr, d);
        cudaError_t e = cudaMemcpy(&flat, d, 8, cudaMemcpyDeviceToHost);
        cudaFree(d);
        TSIM_CUDA(e);
    }
    *n_ns = v[0]; *n_ew = v[1]; *n_cl = v[2]; *n_flat = (long long)flat;
    return TSIM_OK;
}

extern "C" tsim_status tsim_tick_group_ws_bytes(const tsim_cfg *cfg, const tsim_light_tables *lt, long long *bytes) {
    tsim_status r = check_cfg(cfg);
    if (r != TSIM_OK) return r;
    if (!lt || !bytes) { set_error("tsim_tick_group_ws_bytes: NULL argument"); return TSIM_ERR_CONFIG; }
    long long a = 0, b = 0, c = 0, f = 0;
    if ((r = group_list_sizes(lt, &a, &b, &c, &f)) != TSIM_OK) return r;
    const long long n_occ = (long long)((cfg->width + 7) / 8) * ((cfg->win_rows + 7) / 8);
    *bytes = (long long)group_ws_layout(nullptr, a, b, c, n_occ, f, lt->n_groups).bytes;
    return TSIM_OK;
}

// what tick2_init found in a workspace (tick2_run must not synchronise to read the header back)
static struct { const void *ws; long long n_ns, n_ew, n_cl, n_occ, n_flat; int ng; } g_ws_seen = {nullptr, 0, 0, 0, 0, 0, 0};

tsim_status tick2_check(const tsim_tick_state *st, const tsim_tick_tapes *tp) {
    if (!st->probe || !st->recs || !st->plans || !st->ev_stamp || !st->ev_plen || !st->ev_poff || !st->sort_keys || !st->group_ws) {
        set_error("tick (live list): probe / recs / plans / ev_stamp / ev_plen / ev_poff / sort_keys / group_ws must all be set");
        return TSIM_ERR_CONFIG;
    }
    if (st->own_row_lo != 0 || st->own_row_hi != 0) { set_error("tick (live list): row-band shards use the vehicle-indexed kernel"); return TSIM_ERR_UNSUPPORTED; }
    (void)tp;
    return TSIM_OK;
}

tsim_status tick2_init(const tsim_cfg *cfg, const tsim_light_tables *lt, const tsim_tick_tapes *tp, const tsim_tick_state *st, cudaStream_t cs) {
    const size_t n = (size_t)cfg->width * cfg->win_rows, nv = (size_t)tp->n_vehicles;
    {   // light groups: empty occupancy tiles, lanes and clusters as (tile, mask) pairs, the cells of their lights as flat lists
        long long a = 0, b = 0, c = 0, f = 0;
        tsim_status r = group_list_sizes(lt, &a, &b, &c, &f);
        if (r != TSIM_OK) return r;
        const int tiles_x = (cfg->width + 7) / 8;
        const long long n_occ = (long long)tiles_x * ((cfg->win_rows + 7) / 8);
        const GroupWs w = group_ws_layout(st->group_ws, a, b, c, n_occ, f, lt->n_groups);
        const long long hdr[8] = {a, b, c, n_occ, f, 0, 0, 0};
        TSIM_CUDA(cudaMemcpyAsync(st->group_ws, hdr, sizeof(hdr), cudaMemcpyHostToDevice, cs));
        TSIM_CUDA(cudaStreamSynchronize(cs));   // hdr lives on this stack frame
        TSIM_CUDA(cudaMemsetAsync(w.gc_off, 0, 4, cs));
        if (lt->n_groups > 0) {
            group_masks_kernel<<<div_up(3ll * lt->n_groups, 256), 256, 0, cs>>>(lt->n_groups, cfg->width, tiles_x, lt->g_nsin_off, lt->g_nsin, lt->g_ewin_off,
                                                                                 lt->g_ewin, lt->g_cl_off, lt->g_cl, (int)a, (int)(a + b), w.mask, w.tile, w.cnt);
            TSIM_LAUNCH_CHECK();
            group_flat_count_kernel<<<div_up(lt->n_groups, 256), 256, 0, cs>>>(*lt, w.gc_off, nullptr);
            TSIM_LAUNCH_CHECK();
            group_flat_scan_kernel<<<1, 1024, 0, cs>>>(lt->n_groups, w.gc_off);
            TSIM_LAUNCH_CHECK();
            group_flat_fill_kernel<<<div_up(lt->n_groups, 256), 256, 0, cs>>>(*lt, w.gc_off, w.gc_cell, w.gc_role);
            TSIM_LAUNCH_CHECK();
        }
        g_ws_seen = {st->group_ws, a, b, c, n_occ, f, lt->n_groups};
    }
    TSIM_CUDA(cudaMemsetAsync(st->probe, 0, probe_bytes(cfg), cs));   // every plane empty: nobody on the road, every light at go
    if (nv) {
        fill_i32_kernel2<<<div_up((long long)nv, 256) < 1184 ? div_up((long long)nv, 256) : 1184, 256, 0, cs>>>((long long)nv, st->ev_stamp, -1);
        TSIM_LAUNCH_CHECK();
    }
    if (st->tile_ws) {
        int sx, sy, tx, nt;
        tick_tiles(cfg, sx, sy, tx, nt);
        TSIM_CUDA(cudaMemsetAsync(st->tile_ws, 0, (size_t)nt * 2 * sizeof(int32_t), cs));
    }
    return TSIM_OK;   // the live-list counters are part of `scalars`, zeroed by the caller
}

extern "C" tsim_status tsim_debug_tick_phases(unsigned long long *ns16, int32_t reset) {
    if (ns16) TSIM_CUDA(cudaMemcpyFromSymbol(ns16, g_tick_phase_ns, sizeof(unsigned long long) * 16));
    if (reset) { unsigned long long z[16] = {0}; TSIM_CUDA(cudaMemcpyToSymbol(g_tick_phase_ns, z, sizeof(z))); }
    return TSIM_OK;
}

extern "C" tsim_status tsim_tick_tiles(const tsim_cfg *cfg, int32_t *n_tiles) {
    tsim_status r = check_cfg(cfg);
    if (r != TSIM_OK) return r;
    if (!n_tiles) { set_error("tsim_tick_tiles: NULL output"); return TSIM_ERR_CONFIG; }
    int sx, sy, tx, n;
    tick_tiles(cfg, sx, sy, tx, n);
    *n_tiles = n;
    return TSIM_OK;
}

tsim_status tick2_run(const tsim_cfg *cfg, const tsim_light_tables *lt, const tsim_tick_tapes *tp, const tsim_tick_state *st, int32_t n_ticks,
                      int32_t algo, cudaStream_t cs) {
    TickArgs a{cfg->width, cfg->win_rows, n_ticks, algo, 0, cfg->win_rows * cfg->width, *lt, *tp, *st};
    a.sort_every = 16;
    if (st->sort_keys && st->tile_ws) {   // sorted append: worth its extra pass and barrier once the fleet no longer fits the caches
        bool on = tp->n_vehicles >= 200000;
        if (const char *e = getenv("TSIM_TICK_SORT")) on = *e != '0';
        if (on) tick_tiles(cfg, a.tile_sx, a.tile_sy, a.tiles_x, a.n_tiles);
        if (const char *e = getenv("TSIM_TICK_SORT_EVERY")) { const int v = atoi(e); if (v >= 1) a.sort_every = v; }
    }
    if (g_ws_seen.ws != st->group_ws || g_ws_seen.ng != lt->n_groups) {   // another state than the last tsim_tick_init prepared: read its header
        long long hdr[8];
        TSIM_CUDA(cudaMemcpy(hdr, st->group_ws, sizeof(hdr), cudaMemcpyDeviceToHost));
        g_ws_seen = {st->group_ws, hdr[0], hdr[1], hdr[2], hdr[3], hdr[4], lt->n_groups};
    }
    {
        const GroupWs w = group_ws_layout(st->group_ws, g_ws_seen.n_ns, g_ws_seen.n_ew, g_ws_seen.n_cl, g_ws_seen.n_occ, g_ws_seen.n_flat, lt->n_groups);
        a.occ_tiles_x = (cfg->width + 7) / 8;
        a.n_tw = (long long)a.occ_tiles_x * ((cfg->win_rows + 7) / 8);
        a.bits = (unsigned long long *)st->probe;
        if (cfg->width < 2) { set_error("tick (live list): the grid must be at least 2 cells wide"); return TSIM_ERR_CONFIG; }
        a.w_magic = ~0ull / (unsigned long long)cfg->width + 1ull;
        a.occ = a.bits + PL_OCC * a.n_tw; a.gq_mask = w.mask; a.gq_tile = w.tile; a.gq_cnt = w.cnt;
        a.gq_base_ew = (int)g_ws_seen.n_ns; a.gq_base_cl = (int)(g_ws_seen.n_ns + g_ws_seen.n_ew);
        a.gc_off = w.gc_off; a.gc_cell = w.gc_cell; a.gc_role = w.gc_role;
    }
    int dev = 0, sms = 0, per_sm = 0;
    TSIM_CUDA(cudaGetDevice(&dev));
    TSIM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int minb = 2;   // registers per thread: 2 CTAs per SM = as many as the kernel wants, 3 / 4 = capped at 85 / 64
    if (const char *e = getenv("TSIM_TICK_MINB")) { const int v = atoi(e); if (v >= 2 && v <= 4) minb = v; }
    const void *kern = minb == 2 ? (const void *)tick2_kernel<2> : minb == 3 ? (const void *)tick2_kernel<3> : (const void *)tick2_kernel<4>;
    if (minb == 2) TSIM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tick2_kernel<2>, 256, 0));
    else if (minb == 3) TSIM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tick2_kernel<3>, 256, 0));
    else TSIM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tick2_kernel<4>, 256, 0));
    if (per_sm < 1) per_sm = 1;
    // A barrier costs more the more CTAs arrive, a thread with many vehicles serialises their load chains: one CTA per SM up to
    // ~4 vehicles per thread, then more (measured: 100 k vehicles 0.081 / 0.094 / 0.088 ms per tick at 1 / 2 / 4 CTAs per SM,
    // 1 M vehicles 1.22 / 0.85 / 0.71).  TSIM_TICK_CTAS_PER_SM overrides.
    int k = (int)(((long long)tp->n_vehicles + (long long)sms * 256 * 4 - 1) / ((long long)sms * 256 * 4));
    k = k < 1 ? 1 : (k > 4 ? 4 : k);
    if (const char *e = getenv("TSIM_TICK_CTAS_PER_SM")) { const int v = atoi(e); if (v >= 1) k = v; }
    if (k > per_sm) k = per_sm;
    long long want = ((long long)tp->n_vehicles + 255) / 256;
    if (want < (lt->n_groups + 255) / 256) want = (lt->n_groups + 255) / 256;
    const long long cap = (long long)sms * k;
    const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    void *args[] = {&a};
    // The bit planes are what every look-ahead gathers from: ask for them to stay in L2 (persisting lines; everything else a tick
    // reads is streamed once per phase).  TSIM_TICK_L2_PERSIST=0 launches without the window.
    static int max_persist = -1, max_window = 0;
    if (max_persist < 0) {
        TSIM_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
        TSIM_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
        if (const char *e = getenv("TSIM_TICK_L2_PERSIST")) if (*e == '0') max_persist = 0;
        if (getenv("TSIM_DEBUG")) fprintf(stderr, "[tsim] tick: persisting L2 up to %d bytes, access window up to %d bytes\n", max_persist, max_window);
    }
    const size_t plane_bytes = probe_bytes(cfg);
    cudaLaunchAttribute attr[2];
    int n_attr = 0;
    attr[n_attr].id = cudaLaunchAttributeCooperative;
    attr[n_attr++].val.cooperative = 1;
    if (max_persist > 0 && max_window > 0) {
        const size_t win = plane_bytes < (size_t)max_window ? plane_bytes : (size_t)max_window;
        const size_t keep = win < (size_t)max_persist ? win : (size_t)max_persist;
        static size_t limit_set = 0;
        if (limit_set != keep) { TSIM_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, keep)); limit_set = keep; }
        attr[n_attr].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[n_attr].val.accessPolicyWindow.base_ptr = (void *)st->probe;
        attr[n_attr].val.accessPolicyWindow.num_bytes = win;
        attr[n_attr].val.accessPolicyWindow.hitRatio = (float)((double)keep / (double)win);
        attr[n_attr].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[n_attr].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        n_attr++;
    }
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(256); lc.dynamicSmemBytes = 0; lc.stream = cs; lc.attrs = attr; lc.numAttrs = n_attr;
    count_launch();
    TSIM_CUDA(cudaLaunchKernelExC(&lc, kern, args));
    return TSIM_OK;
}

extern "C" tsim_status tsim_tick_export(const tsim_cfg *cfg, const tsim_tick_tapes *tp, const tsim_tick_state *st, void *stream) {
    tsim_status r = check_cfg(cfg);
    if (r != TSIM_OK) return r;
    if (!tp || !st || !st->scalars) { set_error("tsim_tick_export: NULL argument"); return TSIM_ERR_CONFIG; }
    if (!tick2_enabled(st)) return TSIM_OK;   // the vehicle-indexed kernel keeps the SoA itself
    if ((r = tick2_check(st, tp)) != TSIM_OK) return r;
    const int nv = tp->n_vehicles;
    cudaStream_t cs = (cudaStream_t)stream;
    const long long n = (long long)cfg->width * cfg->win_rows;
    if (!st->occupancy || !st->stop_map || !st->stuck_map) { set_error("tsim_tick_export: NULL map"); return TSIM_ERR_CONFIG; }
    TSIM_CUDA(cudaMemsetAsync(st->occupancy, 0, (size_t)n, cs));
    TSIM_CUDA(cudaMemsetAsync(st->stuck_map, 0, (size_t)n, cs));
    {
        const int tiles_x = (cfg->width + 7) / 8;
        const long long n_tw = (long long)tiles_x * ((cfg->win_rows + 7) / 8);
        tick2_export_stop_kernel<<<div_up(n, 256) < 2368 ? div_up(n, 256) : 2368, 256, 0, cs>>>(cfg->width, cfg->win_rows, tiles_x,
                                                                                              (const unsigned long long *)st->probe + PL_STOP * n_tw, st->stop_map);
    }
    TSIM_LAUNCH_CHECK();
    if (nv == 0) return TSIM_OK;
    if (!st->alive || !st->pos || !st->path_off || !st->path_len || !st->steps || !st->stranded || !st->stuck_ticks || !st->base_speed ||
        !st->cur_speed || !st->is_stuck || !st->prev_valid || !st->malfunction || !st->direction) {
        set_error("tsim_tick_export: NULL vehicle array");
        return TSIM_ERR_CONFIG;
    }
    TSIM_CUDA(cudaMemsetAsync(st->alive, 0, (size_t)nv, cs));
    tick2_export_kernel<<<div_up(nv, 256) < 1184 ? div_up(nv, 256) : 1184, 256, 0, cs>>>(*st, nv);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}
