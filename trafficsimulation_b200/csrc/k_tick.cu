// k_tick.cu -- the per-tick vehicle cellular automaton with traffic-light gating, as ONE persistent
// cooperative kernel that advances any number of ticks per launch.
//
// Replaces (reference): CityModel.step city_model.py:1831-1860; VehicleAgent.step_decide / step
// agents/vehicles/vehicle_base.py:616-685 with _tick_stranded :552-565, _check_malfunction :608-610,
// _compute_speed :94-112, _scan_ahead_for_obstacles :422-452, _determine_max_steps :719-731,
// _execute_movement :733-753, _move_to :521-532, tick_stuck :687-693, on_target_reached :755-775;
// CityModel.move_vehicle / remove_vehicle / place_vehicle city_model.py:1897-1963;
// IntersectionLightGroup.step intersection_light_group.py:396-423 (run_queue_actuated :463-494,
// run_fixed_time :427-441, _execute_phase_change :348-384) and CellAgent.set_light_stop/go cell.py:241-251.
//
// Deterministic conflict resolution.  The reference moves vehicles one after another in activation
// order; every cell a vehicle plans to enter was free in the tick-start snapshot, so it can only be
// blocked by the FINAL cell of a vehicle with a lower rank.  That is a triangular system: vehicle v's
// step count depends only on lower-ranked vehicles.  It is solved by Jacobi iteration over a per-cell
// claim word (atomicMin of the rank of the vehicles that currently end there); each sweep fixes at least
// the lowest-ranked undecided vehicle of every dependency chain, the fixed point is unique and equals
// the sequential result.  Chains are as long as the platoons that move bumper to bumper (a few cars).
//
// A tick costs what its grid-wide barriers cost (the data of 100 k vehicles is a few MB), so the phases are arranged
// for as few of them as possible:
//   1 phase A for every live vehicle (tick-start snapshot) + light groups decide and STAGE their stop_map writes
//     (`stopw`: atomicMax of a priority-coded word, last writer in activation order wins); vehicles read the staged
//     value where there is one, so the commit needs no barrier of its own
//   2 claim fixed point, ONE barrier per sweep: claims are generation-tagged 64-bit words (atomicMax: newer sweep,
//     then lower rank, wins) in two planes used alternately, so a sweep reads the previous sweep's claims from one
//     plane while writing its own into the other, and nothing ever has to be cleared
//   3 moves are applied in one phase: a cell somebody enters this tick (it holds a claim of the final sweep) is not
//     cleared by the vehicle that leaves it, so clears and sets cannot race
//   4 spawner claims (lowest attempt index per free origin cell) + commit of the staged stop_map writes
//   5 spawns + scatter of the NEXT tick's route events (replayed A* results)
#include <cooperative_groups.h>
#include "tick_common.cuh"

namespace cg = cooperative_groups;

namespace tsim {

// phase A of one vehicle: vehicle_base.py:616-663 on the tick-start snapshot
__device__ void vehicle_decide(const TickArgs &a, int v, int t) {
    const tsim_tick_state &s = a.st;
    const tsim_tick_tapes &tp = a.tp;
    const size_t tv = (size_t)t * tp.n_vehicles + v;
    s.early[v] = 0;
    s.moved[v] = -1;
    s.max_steps[v] = 0;
    if (s.malfunction[v]) {   // _tick_stranded :552-565
        if (--s.stranded[v] <= 0) { s.malfunction[v] = 0; s.stranded[v] = 0; }
        if (s.malfunction[v]) { s.base_speed[v] = 0; s.cur_speed[v] = 0; s.early[v] = 1; return; }
    }
    if (tp.malfunction[tv] & 2) s.scalars[S_ERR] = 34;   // sideswipe draws that fire are only handled by the live-list kernel
    if (tp.malfunction[tv] & 1) {   // _check_malfunction :608-610
        s.malfunction[v] = 1; s.stranded[v] = MALFUNCTION_TICKS; s.base_speed[v] = 0; s.cur_speed[v] = 0; s.early[v] = 1;
        return;
    }
    const int pos = s.pos[v];
    if (s.stop_map[pos] == 1) { s.base_speed[v] = 0; s.cur_speed[v] = 0; s.early[v] = 1; return; }   // :639-643
    int base = s.base_speed[v];
    if (base == 0) { base = tp.speed[tv]; s.base_speed[v] = (int8_t)base; }   // :94-112
    int sp = base;
    if (tp.rain_map && tp.rain_map[pos] == 1) sp = max(1, sp - RAIN_REDUCTION);
    s.cur_speed[v] = (int8_t)sp;
    // _scan_ahead_for_obstacles :422-452 and _determine_max_steps :719-731
    const int len = s.path_len[v];
    const int32_t *path = tp.ev_cells + s.path_off[v];
    int ms = min(sp, len);
    const int look = min(min(len, AWARENESS), ms);   // cells at or beyond `ms` cannot lower it any more
    if (look <= MAX_SPEED) {
        // all probes first (independent loads in flight together), then the decision
        int cell[MAX_SPEED];
        bool blocked[MAX_SPEED];
#pragma unroll
        for (int i = 0; i < MAX_SPEED; i++) cell[i] = i < look ? path[i] : -1;
#pragma unroll
        for (int i = 0; i < MAX_SPEED; i++)   // a cell beyond the window edge blocks (only ghosts deep in the halo can see one)
            blocked[i] = cell[i] == OUTSIDE || (cell[i] >= 0 && (s.stop_map[cell[i]] == 1 || s.occupancy[cell[i]] == 1));
#pragma unroll
        for (int i = MAX_SPEED - 1; i >= 0; i--) if (blocked[i]) ms = i;
    } else {
        for (int i = 0; i < look; i++) {
            const int c = path[i];
            if (c < 0 || s.stop_map[c] == 1 || s.occupancy[c] == 1) { ms = i; break; }
        }
    }
    s.max_steps[v] = (int8_t)ms;
    if (ms <= 0) {
        s.base_speed[v] = 0;
        if (pos == tp.target[v]) s.scalars[S_ERR] = 30;   // tape contract: a live vehicle is never at its target in phase A
        s.early[v] = 1;
    }
}

// how far does v get, given the claims of the previous sweep?  (_execute_movement :733-753)
__device__ __forceinline__ int vehicle_eval(const TickArgs &a, int v, int rank, const u64 *prev, uint32_t gen_prev) {
    const tsim_tick_state &s = a.st;
    const int m = s.max_steps[v];
    const int32_t *path = a.tp.ev_cells + s.path_off[v];
    int k = 0;
    if (m <= MAX_SPEED) {
        // all probes first (independent loads in flight together), then the walk
        int cell[MAX_SPEED], cr[MAX_SPEED], st[MAX_SPEED];
#pragma unroll
        for (int j = 0; j < MAX_SPEED; j++) cell[j] = j < m ? path[j] : -1;
#pragma unroll
        for (int j = 0; j < MAX_SPEED; j++) {
            cr[j] = cell[j] >= 0 ? claim_rank(prev, cell[j], gen_prev) : NO_CLAIM;
            st[j] = cell[j] >= 0 ? stop_now(s, cell[j]) : 0;
        }
        bool open = true;
#pragma unroll
        for (int j = 0; j < MAX_SPEED; j++) {
            open = open && j < m && !(cr[j] < rank) && !(st[j] == 1 && j + 1 != m);
            if (open) k = j + 1;
        }
        return k;
    }
    for (int j = 1; j <= m; j++) {
        const int c = path[j - 1];
        if (claim_rank(prev, c, gen_prev) < rank) break;       // a lower-ranked vehicle ends here
        if (stop_now(s, c) == 1 && j != m) break;              // a stop cell may only be entered on the last step
        k = j;
    }
    return k;
}

__device__ __forceinline__ void scatter_events(const TickArgs &a, int t, int tid, int nth) {
    const tsim_tick_tapes &tp = a.tp;
    const tsim_tick_state &s = a.st;
    if (t >= tp.n_ticks) return;
    for (int e = tp.ev_first[t] + tid; e < tp.ev_first[t + 1]; e += nth) {
        const int v = tp.ev_vehicle[e];
        s.path_off[v] = tp.ev_off[e];
        s.path_len[v] = (int32_t)(tp.ev_off[e + 1] - tp.ev_off[e]);
    }
}

__global__ void __launch_bounds__(256) tick_kernel(TickArgs a) {
    cg::grid_group grid = cg::this_grid();
    const tsim_tick_state &s = a.st;
    const tsim_tick_tapes &tp = a.tp;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int nv = tp.n_vehicles, ng = a.lt.n_groups;
    const size_t ncell = (size_t)a.W * a.H;
    u64 *plane[2] = {(u64 *)s.claim, (u64 *)s.claim + ncell};
    // route events of the first tick of this launch (later ticks: scattered at the end of the previous tick)
    scatter_events(a, *((volatile int32_t *)(s.scalars + S_TICK)), tid, nth);
    grid.sync();
    for (int it = 0; it < a.n_ticks; it++) {
        const int t = *((volatile int32_t *)(s.scalars + S_TICK));
        if (t >= tp.n_ticks) { if (tid == 0) s.scalars[S_ERR] = 31; break; }
        const uint32_t gen0 = (uint32_t)t * GEN_PER_TICK + 1u;   // generation nobody writes: "no claims yet"
        // a shard iterates the list of the vehicles alive in its window (built after the last halo refresh: valid for one tick)
        const bool use_list = it == 0 && s.live_idx && *((volatile int32_t *)(s.scalars + S_LIST_OK)) != 0;
        const int n_it = use_list ? min(*((volatile int32_t *)(s.scalars + S_NLIST)), nv) : nv;
        if (a.algo == 3 && ng > 0) green_wave_prepass<false>(a, grid, tid, nth);
        // ---- 1: phase A + light-group decisions (staged)
        int live = 0;
        for (int i = tid; i < n_it; i += nth) {
            const int v = use_list ? s.live_idx[i] : i;
            if (s.alive[v]) { vehicle_decide(a, v, t); const int p = s.pos[v]; live += p >= a.own_lo && p < a.own_hi; }
        }
        live = __reduce_add_sync(0xffffffffu, live);
        if ((threadIdx.x & 31) == 0 && live) atomicAdd((unsigned long long *)(s.scalars + S_UPD_HI), (unsigned long long)live);
        for (int g = tid; g < ng; g += nth) group_decide<false>(a, g);
        grid.sync();
        // ---- 2: claim fixed point, one barrier per sweep
        const int32_t *rank = tp.rank + (size_t)t * nv;
        int last = 0;
        for (int iter = 0;; iter++) {
            const int fidx[3] = {S_FLAG0, S_FLAG1, S_FLAG2};   // three flags in rotation: the one zeroed after sweep i is written in sweep i + 2
            int32_t *flag = s.scalars + fidx[iter % 3];
            const u64 *prev = plane[(iter + 1) & 1];
            u64 *cur = plane[iter & 1];
            const uint32_t gen_prev = gen0 + iter, gen_cur = gen0 + iter + 1;
            bool ch = false;
            for (int i = tid; i < n_it; i += nth) {
                const int v = use_list ? s.live_idx[i] : i;
                if (!s.alive[v] || s.early[v]) continue;
                const int r = rank[v];
                const int k = vehicle_eval(a, v, r, prev, gen_prev);
                if (k != s.moved[v]) { s.moved[v] = (int8_t)k; ch = true; }
                if (k >= 1) {
                    const int c = tp.ev_cells[s.path_off[v] + k - 1];
                    if (c != tp.target[v]) claim_cell(cur, c, gen_cur, r);   // an arriving vehicle is removed at once
                }
            }
            if (__any_sync(0xffffffffu, ch) && (threadIdx.x & 31) == 0) *flag = 1;
            grid.sync();
            const int any = *((volatile int32_t *)flag);
            if (tid == 0) { s.scalars[fidx[(iter + 2) % 3]] = 0; s.scalars[S_ITERS]++; }
            last = iter;
            if (!any) break;
            if (iter >= GEN_PER_TICK - 4) { if (tid == 0) s.scalars[S_ERR] = 32; break; }
        }
        const u64 *fin_plane = plane[last & 1];
        const uint32_t gen_fin = gen0 + last + 1;
        // ---- 3: apply the moves
        for (int i = tid; i < n_it; i += nth) {
            const int v = use_list ? s.live_idx[i] : i;
            if (!s.alive[v]) continue;
            int pos = s.pos[v];
            const int target = tp.target[v];
            if (!s.early[v]) {
                const int k = s.moved[v];
                if (k >= 1) {
                    const int32_t *path = tp.ev_cells + s.path_off[v];
                    const int fin = path[k - 1], prev = k >= 2 ? path[k - 2] : pos;
                    // cells left or passed through: cleared unless somebody ends there this tick (their set wins)
                    if (claim_rank(fin_plane, pos, gen_fin) == NO_CLAIM) { s.occupancy[pos] = 0; s.stuck_map[pos] = 0; }   // move_vehicle city_model.py:1952,1957
                    for (int j = 0; j + 1 < k; j++)
                        if (claim_rank(fin_plane, path[j], gen_fin) == NO_CLAIM) s.stuck_map[path[j]] = 0;
                    if (fin != target) {
                        s.occupancy[fin] = 1;
                        s.stuck_map[fin] = (k == 1 && s.is_stuck[v]) ? 1 : 0;   // move_vehicle :1956-1958, before _move_to resets is_stuck
                    } else if (claim_rank(fin_plane, fin, gen_fin) == NO_CLAIM) {
                        s.stuck_map[fin] = 0;                                    // arrives: remove_vehicle clears its cell again (unless a later-ranked vehicle ends there)
                    }
                    const int d = fin - prev;                                    // compute_direction numba_utilities.py:14-28
                    s.direction[v] = (int8_t)(d == a.W ? DN : d == 1 ? DE : d == -a.W ? DS : d == -1 ? DW : s.direction[v]);
                    if (s.stuck_ticks[v] > 0) { s.is_stuck[v] = 0; s.stuck_ticks[v] = 0; }   // _move_to :528-532
                    s.steps[v] += k;
                    s.path_off[v] += k; s.path_len[v] -= k;
                    s.pos[v] = pos = fin;
                }
                s.prev_valid[v] = 1;   // step() :677
            } else {
                s.early[v] = 0;        // :679-680 tick_stuck :687-693
                if (s.prev_valid[v] && stop_now(s, pos) != 1) {
                    const int st = ++s.stuck_ticks[v];
                    if (st > STUCK_THRESHOLD && !s.is_stuck[v]) s.is_stuck[v] = 1;
                }
                if (pos == target) { s.occupancy[pos] = 0; s.stuck_map[pos] = 0; }
            }
            if (pos == target) s.alive[v] = 0;   // on_target_reached :755-775 -> remove_vehicle city_model.py:1920-1929
        }
        grid.sync();
        // ---- 4: spawner claims (attempts of this tick, lowest attempt index wins a free cell) + commit of the staged stop_map writes
        const int k0 = tp.spawn_first[t], k1 = tp.spawn_first[t + 1];
        const uint32_t gen_spawn = (uint32_t)t * GEN_PER_TICK + GEN_PER_TICK - 1;
        for (int k = k0 + tid; k < k1; k += nth)
            if (tp.origin[k] >= 0 && s.occupancy[tp.origin[k]] == 0) claim_cell(plane[0], tp.origin[k], gen_spawn, k);
        for (int g = tid; g < ng; g += nth) group_apply<false>(a, g);
        if (tid == 0) { s.scalars[S_FLAG0] = 0; s.scalars[S_FLAG1] = 0; s.scalars[S_FLAG2] = 0; }   // nobody touches the sweep flags here
        grid.sync();
        // ---- 5: spawns, the next tick's route events
        for (int k = k0 + tid; k < k1; k += nth) {
            const int o = tp.origin[k];
            if (o < 0 || claim_rank(plane[0], o, gen_spawn) != k) continue;
            s.alive[k] = 1; s.pos[k] = o;
            s.base_speed[k] = 0; s.cur_speed[k] = 0; s.max_steps[k] = 0; s.early[k] = 0; s.is_stuck[k] = 0; s.prev_valid[k] = 0;
            s.malfunction[k] = 0; s.direction[k] = -1; s.stuck_ticks[k] = 0; s.stranded[k] = 0; s.steps[k] = 0; s.moved[k] = 0;
            s.occupancy[o] = 1; s.stuck_map[o] = 0;   // place_vehicle city_model.py:1904-1907
        }
        scatter_events(a, t + 1, tid, nth);
        if (tid == 0) { s.scalars[S_TICK] = t + 1; s.scalars[S_LIST_OK] = 0; }   // this tick's spawns are not on the list
        grid.sync();
    }
}

// ---- shards (SURVEY.md 8e "Vehicle step"): after every tick a shard sends ONE message to each neighbour ------------------
// message (int32 words): [0, HDR) header, word 0 = number of vehicle records | records, REC words each | state of the
// light groups both shards simulate, GST words each | the three map planes' rows (occupancy, stop, stuck), bytes.
// record: 0 vehicle, 1 global row, 2 column, 3 path_len, 4 steps, 5 stranded, 6..7 path_off,
// 8 stuck_ticks | base_speed << 16 | cur_speed << 24, 9 is_stuck | prev_valid << 8 | malfunction << 16 | direction << 24
constexpr int REC = TSIM_TICK_REC_WORDS, REC_HDR = TSIM_TICK_REC_HEADER, GST = TSIM_TICK_GROUP_WORDS;
enum { ALIVE = 1, ZOMBIE = 2 };   // ZOMBIE: a ghost of the halo awaiting its owner's record

__host__ __device__ inline long long msg_groups_at(int cap) { return REC_HDR + (long long)cap * REC; }
__host__ __device__ inline long long msg_planes_at(int cap, int n_groups) { return msg_groups_at(cap) + (((long long)n_groups * GST + 3) & ~3LL); }
__host__ __device__ inline long long msg_plane_bytes(int rows, int W) { return ((long long)rows * W + 15) & ~15LL; }
__host__ __device__ inline long long msg_words(int cap, int n_groups, int rows, int W) {
    return msg_planes_at(cap, n_groups) + 3 * msg_plane_bytes(rows, W) / 4;
}

struct StripArgs {
    int W, y0, nv, cap;
    tsim_tick_strips sp;
    tsim_tick_state st;
};

__device__ __forceinline__ void make_record(const tsim_tick_state &s, int v, int W, int y0, int32_t *r) {
    const int pos = s.pos[v];
    const long long off = s.path_off[v];
    r[0] = v; r[1] = pos / W + y0; r[2] = pos % W; r[3] = s.path_len[v]; r[4] = s.steps[v]; r[5] = s.stranded[v];
    r[6] = (int32_t)(off & 0xffffffffLL); r[7] = (int32_t)(off >> 32);
    r[8] = (int32_t)((uint32_t)(uint16_t)s.stuck_ticks[v] | ((uint32_t)(uint8_t)s.base_speed[v] << 16) | ((uint32_t)(uint8_t)s.cur_speed[v] << 24));
    r[9] = (int32_t)((uint32_t)(uint8_t)s.is_stuck[v] | ((uint32_t)(uint8_t)s.prev_valid[v] << 8) | ((uint32_t)(uint8_t)s.malfunction[v] << 16) |
                     ((uint32_t)(uint8_t)s.direction[v] << 24));
    r[10] = 0; r[11] = 0;
}

__device__ __forceinline__ int32_t *group_field(const tsim_tick_state &s, int f) {
    return f == 0 ? s.g_cur : f == 1 ? s.g_pend : f == 2 ? s.g_qt : f == 3 ? s.g_gap : f == 4 ? s.g_last : f == 5 ? s.g_ft_phase : s.g_ft_timer;
}

__device__ __forceinline__ bool differ(uint4 a, uint4 b) { return ((a.x ^ b.x) | (a.y ^ b.y) | (a.z ^ b.z) | (a.w ^ b.w)) != 0; }
__device__ __forceinline__ bool differ(uint32_t a, uint32_t b) { return a != b; }
__device__ __forceinline__ bool differ(uint8_t a, uint8_t b) { return a != b; }

// rows [lo, hi) of the three map planes <-> the plane part of a message, V bytes at a time.  INSTALL: message -> planes,
// and inside the verify rows the ghost rows this shard simulated must already hold what the owner sends.
template <class V, bool INSTALL>
__device__ void plane_rows(const StripArgs &a, int d, int32_t *msg, int n_groups, int lo, int hi, int tid, int nth) {
    const tsim_tick_state &s = a.st;
    const long long nbytes = (long long)(hi - lo) * a.W, first = (long long)lo * a.W;
    const long long v0 = (long long)a.sp.verify_lo[d] * a.W, v1 = (long long)a.sp.verify_hi[d] * a.W;
    uint8_t *body = (uint8_t *)(msg + msg_planes_at(a.cap, n_groups));
    uint8_t *planes[3] = {s.occupancy, s.stop_map, s.stuck_map};
    const long long nvec = nbytes / (long long)sizeof(V);
    bool bad = false;
    for (int p = 0; p < 3; p++) {
        V *m = (V *)(body + p * msg_plane_bytes(hi - lo, a.W));
        V *g = (V *)(planes[p] + first);
        for (long long i = tid; i < nvec; i += nth) {
            if (INSTALL) {
                const V got = m[i];
                const long long at = first + i * (long long)sizeof(V);
                if (at >= v0 && at < v1) {
                    const V mine = g[i];
                    bad |= differ(mine, got);
                }
                g[i] = got;
            } else {
                m[i] = g[i];
            }
        }
    }
    if (bad) s.scalars[S_XERR] = 44;   // a halo row differs from its owner's: the halo is too small for this traffic
}

template <bool INSTALL>
__device__ void plane_rows_any(const StripArgs &a, int d, int32_t *msg, int n_groups, int lo, int hi, int tid, int nth) {
    if (a.W % 16 == 0) plane_rows<uint4, INSTALL>(a, d, msg, n_groups, lo, hi, tid, nth);
    else if (a.W % 4 == 0) plane_rows<uint32_t, INSTALL>(a, d, msg, n_groups, lo, hi, tid, nth);
    else plane_rows<uint8_t, INSTALL>(a, d, msg, n_groups, lo, hi, tid, nth);
}

// vehicles on the own rows next to a cut, the shared groups' state and the own rows themselves go into the message for
// that neighbour; vehicles on halo rows become zombies
__global__ void __launch_bounds__(256) tick_pack_kernel(StripArgs a) {
    const tsim_tick_state &s = a.st;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int v = tid; v < a.nv; v += nth) {
        if (s.alive[v] != ALIVE) continue;
        const int row = s.pos[v] / a.W;
#pragma unroll
        for (int d = 0; d < 2; d++) {
            if (row >= a.sp.halo_lo[d] && row < a.sp.halo_hi[d]) s.alive[v] = ZOMBIE;
            if (a.sp.send_msg[d] && row >= a.sp.send_lo[d] && row < a.sp.send_hi[d]) {
                const int k = atomicAdd(a.sp.send_msg[d], 1);
                if (k < a.cap) make_record(s, v, a.W, a.y0, a.sp.send_msg[d] + REC_HDR + (size_t)k * REC);
                else s.scalars[S_XERR] = 43;   // strip holds more vehicles than the record buffer
            }
        }
    }
    for (int d = 0; d < 2; d++) {
        int32_t *msg = a.sp.send_msg[d];
        if (!msg) continue;
        int32_t *gw = msg + msg_groups_at(a.cap);
        for (int i = tid; i < a.sp.n_g_send[d] * GST; i += nth) gw[i] = group_field(s, i % GST)[a.sp.g_send[d][i / GST]];
        plane_rows_any<false>(a, d, msg, a.sp.n_g_send[d], a.sp.send_lo[d], a.sp.send_hi[d], tid, nth);
    }
}

// the owner's message replaces the halo: rows, group state, vehicles; inside the verify rows what this shard simulated
// must be identical to what arrives
__global__ void __launch_bounds__(256) tick_unpack_kernel(StripArgs a) {
    const tsim_tick_state &s = a.st;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
#pragma unroll
    for (int d = 0; d < 2; d++) {
        int32_t *buf = a.sp.recv_msg[d];
        if (!buf) continue;
        const int n = min(buf[0], a.cap);
        for (int k = tid; k < n; k += nth) {
            const int32_t *r = buf + REC_HDR + (size_t)k * REC;
            const int v = r[0], row = r[1] - a.y0;
            if (v < 0 || v >= a.nv || row < a.sp.halo_lo[d] || row >= a.sp.halo_hi[d]) { s.scalars[S_XERR] = 40; continue; }
            if (row >= a.sp.verify_lo[d] && row < a.sp.verify_hi[d]) {
                int32_t mine[REC];
                bool same = s.alive[v] == ZOMBIE;
                if (same) {
                    make_record(s, v, a.W, a.y0, mine);
                    for (int i = 1; i < 10; i++) same = same && mine[i] == r[i];
                }
                if (!same) s.scalars[S_XERR] = 41;   // the ghost diverged from its owner: the halo is too small for this traffic
            }
            s.alive[v] = ALIVE;
            s.pos[v] = row * a.W + r[2]; s.path_len[v] = r[3]; s.steps[v] = r[4]; s.stranded[v] = r[5];
            s.path_off[v] = (long long)(uint32_t)r[6] | ((long long)r[7] << 32);
            const uint32_t w8 = (uint32_t)r[8], w9 = (uint32_t)r[9];
            s.stuck_ticks[v] = (int16_t)(w8 & 0xffff); s.base_speed[v] = (int8_t)(w8 >> 16); s.cur_speed[v] = (int8_t)(w8 >> 24);
            s.is_stuck[v] = (int8_t)w9; s.prev_valid[v] = (int8_t)(w9 >> 8); s.malfunction[v] = (int8_t)(w9 >> 16); s.direction[v] = (int8_t)(w9 >> 24);
            s.early[v] = 0; s.moved[v] = 0; s.max_steps[v] = 0;
        }
        const int32_t *gw = buf + msg_groups_at(a.cap);
        for (int i = tid; i < a.sp.n_g_recv[d] * GST; i += nth) {
            int32_t *mine = group_field(s, i % GST) + a.sp.g_recv[d][i / GST];
            if (a.sp.g_verify[d][i / GST] && *mine != gw[i]) s.scalars[S_XERR] = 45;   // a ghost light group diverged from its owner
            *mine = gw[i];
        }
        plane_rows_any<true>(a, d, buf, a.sp.n_g_recv[d], a.sp.halo_lo[d], a.sp.halo_hi[d], tid, nth);
    }
}

// zombies nobody claimed are gone (they left the window or were never real); inside the verify rows that is a divergence
__global__ void __launch_bounds__(256) tick_reap_kernel(StripArgs a) {
    const tsim_tick_state &s = a.st;
    const int lane = threadIdx.x & 31;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int v0 = tid - lane; v0 < a.nv; v0 += nth) {   // whole warps: the list append is one atomic per warp
        const int v = v0 + lane;
        bool keep = false;
        if (v < a.nv) {
            const int al = s.alive[v];
            if (al == ZOMBIE) {
                const int row = s.pos[v] / a.W;
                for (int d = 0; d < 2; d++)
                    if (row >= a.sp.verify_lo[d] && row < a.sp.verify_hi[d]) s.scalars[S_XERR] = 42;
                s.alive[v] = 0;
            }
            keep = al == ALIVE;
        }
        if (!s.live_idx) continue;
        const uint32_t mask = __ballot_sync(0xffffffffu, keep);
        if (!mask) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(s.scalars + S_NLIST, __popc(mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) s.live_idx[base + __popc(mask & ((1u << lane) - 1u))] = v;
    }
    if (tid == 0 && s.live_idx) s.scalars[S_LIST_OK] = 1;   // the next tick_run (a later launch) may use the list
}

__global__ void fill_i32_kernel(long long n, int32_t *p, int32_t v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace tsim

using namespace tsim;

// local cell range of the rows the window's shard owns
static void own_cells(const tsim_cfg *cfg, const tsim_tick_state *st, int &lo, int &hi) {
    const bool all = st->own_row_lo == 0 && st->own_row_hi == 0;
    lo = all ? 0 : st->own_row_lo * cfg->width;
    hi = all ? cfg->win_rows * cfg->width : st->own_row_hi * cfg->width;
}

static tsim_status strip_args(const tsim_cfg *cfg, const tsim_tick_tapes *tp, const tsim_tick_state *st, const tsim_tick_strips *sp, StripArgs &a) {
    tsim_status r = check_cfg(cfg);
    if (r != TSIM_OK) return r;
    if (!tp || !st || !sp || sp->cap < 1 || !st->pos || !st->alive || !st->scalars || !st->occupancy || !st->stop_map || !st->stuck_map) {
        set_error("tick strips: bad arguments");
        return TSIM_ERR_CONFIG;
    }
    for (int d = 0; d < 2; d++) {
        if (sp->send_lo[d] < 0 || sp->send_hi[d] > cfg->win_rows || sp->halo_lo[d] < 0 || sp->halo_hi[d] > cfg->win_rows ||
            sp->send_hi[d] < sp->send_lo[d] || sp->halo_hi[d] < sp->halo_lo[d] ||
            (sp->verify_hi[d] > sp->verify_lo[d] && (sp->verify_lo[d] < sp->halo_lo[d] || sp->verify_hi[d] > sp->halo_hi[d]))) {
            set_error("tick strips: row ranges outside the window");
            return TSIM_ERR_CONFIG;
        }
        if (sp->n_g_send[d] < 0 || sp->n_g_recv[d] < 0 || (sp->n_g_send[d] > 0 && !sp->g_send[d]) ||
            (sp->n_g_recv[d] > 0 && (!sp->g_recv[d] || !sp->g_verify[d]))) {
            set_error("tick strips: bad light-group lists");
            return TSIM_ERR_CONFIG;
        }
    }
    a = StripArgs{cfg->width, cfg->win_y0, tp->n_vehicles, sp->cap, *sp, *st};
    return TSIM_OK;
}

static int strip_grid(long long n) { const long long g = (n + 255) / 256; return g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : (int)g); }

extern "C" long long tsim_tick_message_words(int32_t width, int32_t cap, int32_t n_groups, int32_t rows) {
    return msg_words(cap, n_groups, rows, width);
}

extern "C" tsim_status tsim_tick_pack(const tsim_cfg *cfg, const tsim_tick_tapes *tp, const tsim_tick_state *st, const tsim_tick_strips *sp,
                                      void *stream) {
    StripArgs a;
    tsim_status r = strip_args(cfg, tp, st, sp, a);
    if (r != TSIM_OK) return r;
    cudaStream_t cs = (cudaStream_t)stream;
    long long work = a.nv;
    for (int d = 0; d < 2; d++)
        if (sp->send_msg[d]) {
            TSIM_CUDA(cudaMemsetAsync(sp->send_msg[d], 0, REC_HDR * sizeof(int32_t), cs));
            const long long cells = (long long)(sp->send_hi[d] - sp->send_lo[d]) * a.W / 16;
            if (cells > work) work = cells;
        }
    tick_pack_kernel<<<strip_grid(work), 256, 0, cs>>>(a);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}

extern "C" tsim_status tsim_tick_unpack(const tsim_cfg *cfg, const tsim_tick_tapes *tp, const tsim_tick_state *st, const tsim_tick_strips *sp,
                                        void *stream) {
    StripArgs a;
    tsim_status r = strip_args(cfg, tp, st, sp, a);
    if (r != TSIM_OK) return r;
    cudaStream_t cs = (cudaStream_t)stream;
    if (sp->recv_msg[0] || sp->recv_msg[1]) {
        long long work = a.cap;
        for (int d = 0; d < 2; d++) {
            const long long cells = (long long)(sp->halo_hi[d] - sp->halo_lo[d]) * a.W / 16;
            if (sp->recv_msg[d] && cells > work) work = cells;
        }
        tick_unpack_kernel<<<strip_grid(work), 256, 0, cs>>>(a);
        TSIM_LAUNCH_CHECK();
    }
    if (a.nv > 0) {
        if (st->live_idx) TSIM_CUDA(cudaMemsetAsync(st->scalars + S_NLIST, 0, 2 * sizeof(int32_t), cs));
        tick_reap_kernel<<<strip_grid(a.nv), 256, 0, cs>>>(a);
        TSIM_LAUNCH_CHECK();
    }
    return TSIM_OK;
}

// the live-list kernel (k_tick2.cu)
bool tick2_enabled(const tsim_tick_state *st);
tsim_status tick2_check(const tsim_tick_state *st, const tsim_tick_tapes *tp);
tsim_status tick2_init(const tsim_cfg *cfg, const tsim_light_tables *lt, const tsim_tick_tapes *tp, const tsim_tick_state *st, cudaStream_t cs);
tsim_status tick2_run(const tsim_cfg *cfg, const tsim_light_tables *lt, const tsim_tick_tapes *tp, const tsim_tick_state *st, int32_t n_ticks,
                      int32_t algo, cudaStream_t cs);

static tsim_status check_tick_args(const tsim_cfg *cfg, const tsim_light_tables *lt, const tsim_tick_tapes *tp, const tsim_tick_state *st) {
    tsim_status s = check_cfg(cfg);
    if (s != TSIM_OK) return s;
    if (!lt || !tp || !st) { set_error("tick: NULL argument struct"); return TSIM_ERR_CONFIG; }
    if (tp->n_vehicles < 0 || tp->n_ticks < 1 || lt->n_groups < 0) { set_error("tick: bad sizes"); return TSIM_ERR_CONFIG; }
    if (st->own_row_lo < 0 || st->own_row_hi < st->own_row_lo || st->own_row_hi > cfg->win_rows) { set_error("tick: bad own rows"); return TSIM_ERR_CONFIG; }
    if (!st->occupancy || !st->stop_map || !st->stuck_map || !st->claim || !st->stopw || !st->scalars) { set_error("tick: NULL map / scratch plane"); return TSIM_ERR_CONFIG; }
    // the kernels read spawn_first / ev_first every tick whatever the fleet size, and the group arrays whenever there are groups
    if (!tp->spawn_first || !tp->ev_first) { set_error("tick: NULL spawn_first / ev_first"); return TSIM_ERR_CONFIG; }
    if (tp->n_vehicles > 0 && (!st->pos || !st->path_off || !st->path_len || !st->alive || !st->moved || !tp->origin || !tp->target || !tp->speed ||
                               !tp->malfunction || !tp->rank || !st->steps || !st->stranded || !st->stuck_ticks || !st->base_speed || !st->cur_speed ||
                               !st->max_steps || !st->early || !st->is_stuck || !st->prev_valid || !st->malfunction || !st->direction)) {
        set_error("tick: NULL vehicle array / tape");
        return TSIM_ERR_CONFIG;
    }
    if (lt->n_groups > 0 && (!st->g_cur || !st->g_pend || !st->g_qt || !st->g_gap || !st->g_last || !st->g_ft_phase || !st->g_ft_timer || !st->g_plan ||
                             !lt->tl_off || !lt->tl_cells || !lt->g_all_off || !lt->g_all || !lt->g_ns_off || !lt->g_ns || !lt->g_ew_off || !lt->g_ew ||
                             !lt->g_nsin_off || !lt->g_nsin || !lt->g_ewin_off || !lt->g_ewin || !lt->g_cl_off || !lt->g_cl)) {
        set_error("tick: NULL light-group array / table");
        return TSIM_ERR_CONFIG;
    }
    if (tick2_enabled(st)) return tick2_check(st, tp);
    return TSIM_OK;
}

extern "C" tsim_status tsim_tick_init(const tsim_cfg *cfg, const tsim_light_tables *lt, const tsim_tick_tapes *tp, const tsim_tick_state *st,
                                      void *stream) {
    tsim_status r = check_tick_args(cfg, lt, tp, st);
    if (r != TSIM_OK) return r;
    cudaStream_t cs = (cudaStream_t)stream;
    const size_t n = (size_t)cfg->width * cfg->win_rows, nv = (size_t)tp->n_vehicles, ng = (size_t)lt->n_groups;
    TSIM_CUDA(cudaMemsetAsync(st->occupancy, 0, n, cs));
    TSIM_CUDA(cudaMemsetAsync(st->stop_map, 0, n, cs));
    TSIM_CUDA(cudaMemsetAsync(st->stuck_map, 0, n, cs));
    TSIM_CUDA(cudaMemsetAsync(st->stopw, 0, n * 4, cs));
    TSIM_CUDA(cudaMemsetAsync(st->claim, 0, n * 16, cs));   // two planes of 64-bit generation-tagged claims; generation 0 is never read
    if (nv) {
        fill_i32_kernel<<<div_up((long long)nv, 256) < 1184 ? div_up((long long)nv, 256) : 1184, 256, 0, cs>>>((long long)nv, st->pos, -1);
        TSIM_LAUNCH_CHECK();
        TSIM_CUDA(cudaMemsetAsync(st->path_len, 0, nv * 4, cs));
        TSIM_CUDA(cudaMemsetAsync(st->steps, 0, nv * 4, cs));
        TSIM_CUDA(cudaMemsetAsync(st->stranded, 0, nv * 4, cs));
        TSIM_CUDA(cudaMemsetAsync(st->path_off, 0, nv * 8, cs));
        TSIM_CUDA(cudaMemsetAsync(st->stuck_ticks, 0, nv * 2, cs));
        int8_t *bytes[] = {st->alive, st->base_speed, st->cur_speed, st->max_steps, st->early, st->is_stuck, st->prev_valid, st->malfunction, st->moved};
        for (int8_t *b : bytes) TSIM_CUDA(cudaMemsetAsync(b, 0, nv, cs));
        TSIM_CUDA(cudaMemsetAsync(st->direction, 0xff, nv, cs));
    }
    if (ng) {
        fill_i32_kernel<<<div_up((long long)ng, 256) < 1184 ? div_up((long long)ng, 256) : 1184, 256, 0, cs>>>((long long)ng, st->g_cur, -1);
        TSIM_LAUNCH_CHECK();
        int32_t *zero[] = {st->g_pend, st->g_qt, st->g_gap, st->g_last, st->g_ft_phase, st->g_ft_timer, st->g_plan};   // g_pend = 0: apply_phase(0) in __init__ (:115-116)
        for (int32_t *z : zero) TSIM_CUDA(cudaMemsetAsync(z, 0, ng * 4, cs));
    }
    TSIM_CUDA(cudaMemsetAsync(st->scalars, 0, 16 * 4, cs));
    if (tick2_enabled(st)) return tick2_init(cfg, lt, tp, st, cs);
    return TSIM_OK;
}

extern "C" tsim_status tsim_tick_run(const tsim_cfg *cfg, const tsim_light_tables *lt, const tsim_tick_tapes *tp, const tsim_tick_state *st,
                                     int32_t n_ticks, int32_t algo, void *stream) {
    tsim_status r = check_tick_args(cfg, lt, tp, st);
    if (r != TSIM_OK) return r;
    if (n_ticks < 1 || algo < 0 || algo > 3) { set_error("tick: n_ticks %d algo %d", n_ticks, algo); return TSIM_ERR_CONFIG; }
    if (algo == 3 && lt->n_groups > 0 && (!lt->g_nbr || !st->g_wave)) {
        set_error("tick: NEIGHBOR_GREEN_WAVE needs the g_nbr table and the g_wave scratch");
        return TSIM_ERR_CONFIG;
    }
    if (algo == 2 && lt->n_groups > 0 && (!lt->g_nsout_off || !lt->g_nsout || !lt->g_ewout_off || !lt->g_ewout)) {
        set_error("tick: PRESSURE_CONTROL needs the g_nsout / g_ewout tables");
        return TSIM_ERR_CONFIG;
    }
    if (algo >= 2 && (cfg->win_y0 != 0 || cfg->win_rows != cfg->height)) {
        set_error("tick: PRESSURE_CONTROL / NEIGHBOR_GREEN_WAVE read cells / groups outside a shard window (tsim.h): whole cities only");
        return TSIM_ERR_CONFIG;
    }
    if (tick2_enabled(st)) return tick2_run(cfg, lt, tp, st, n_ticks, algo, (cudaStream_t)stream);
    TickArgs a{cfg->width, cfg->win_rows, n_ticks, algo, 0, 0, *lt, *tp, *st};
    own_cells(cfg, st, a.own_lo, a.own_hi);
    int dev = 0, sms = 0, per_sm = 0;
    TSIM_CUDA(cudaGetDevice(&dev));
    TSIM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TSIM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tick_kernel, 256, 0));
    if (per_sm < 1) per_sm = 1;
    // grid.sync cost grows with the CTA count: use just enough CTAs for the work, at most one full wave
    long long want = ((long long)tp->n_vehicles + 255) / 256;
    if (want < (lt->n_groups + 255) / 256) want = (lt->n_groups + 255) / 256;
    long long cap = (long long)sms * ((per_sm > 4 && tp->n_vehicles <= 262144) ? 4 : per_sm);   // small fleets: fewer CTAs, cheaper barriers
    int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    void *args[] = {&a};
    TSIM_COOP_LAUNCH(tick_kernel, dim3(grid), dim3(256), args, (cudaStream_t)stream);
    return TSIM_OK;
}
