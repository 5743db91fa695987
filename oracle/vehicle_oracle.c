/*
 * vehicle_oracle.c -- CPU restatement of one simulation tick.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows, in the reference (kurisu-n/TrafficSimulation):
 *   CityModel.step                       Simulation/city_model.py:1831-1860
 *   VehicleAgent.step_decide / step      Simulation/agents/vehicles/vehicle_base.py:616-685
 *     _tick_stranded :552-565, _check_malfunction :608-610, _check_sideswipe_collision :567-605 (+ _set_collision :534-541),
 *     _is_at_stopped_cell :121-127,
 *     _compute_speed :94-112, _scan_ahead_for_obstacles :422-452, _determine_max_steps :719-731,
 *     _execute_movement :733-753, _move_to :521-532, tick_stuck :687-693, on_target_reached :755-775
 *   CityModel.move_vehicle / remove_vehicle / place_vehicle   city_model.py:1897-1963
 *   IntersectionLightGroup.step          agents/city_structure_entities/intersection_light_group.py:396-423
 *     run_queue_actuated :463-494, run_fixed_time :427-441, run_pressure_control :448-461 (compute_max_pressure
 *     utilities/numba_utilities.py:74-85), run_neighbor_green_wave :522-546,
 *     apply_phase :386-393,
 *     _execute_phase_change :348-384, is_intersection_occupied :285-291
 *   CellAgent.set_light_stop / set_light_go   agents/city_structure_entities/cell.py:241-251
 *
 * under the tape conventions of oracle/refharness/ticks.py (activation order, speed / malfunction /
 * rank tapes, tape-driven spawner that drops attempts onto occupied cells, replayed route events; bit 0 of a
 * malfunction tape entry = the malfunction draw fires, bit 1 = the sideswipe draw fires IF it is made).
 * Pinned against the live reference by tests/test_ticks_vs_reference.py and tests/golden/ticks_*.npz.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t W, H, n_vehicles, n_groups, n_lights;
    int32_t algo;                 /* 0 QUEUE_ACTUATED, 1 FIXED_TIME, 2 PRESSURE_CONTROL, 3 NEIGHBOR_GREEN_WAVE (config.py:341).
                                     NEIGHBOR_PRESSURE_CONTROL (:496-520) is not restated: every group subtracts its neighbours' pressures from
                                     its own, the neighbours do the same, and the values double every tick or two -- Python's integers follow
                                     them past 2^63 within about 60 ticks (seen live: 2.7e9 after 30), fixed-width ones cannot */
    int32_t rain_enabled;         /* Defaults.RAIN_ENABLED */
    int32_t tick;                 /* next tick to run */
    /* maps [H*W] */
    uint8_t *occ, *stop, *stuckmap; const uint8_t *rain;
    /* vehicle tapes */
    const int32_t *spawn_tick, *origin, *target;
    const uint8_t *speed, *malf;  /* [n_ticks][n_vehicles] */
    const int32_t *rank;          /* [n_ticks][n_vehicles] */
    /* route events, sorted by tick; ev_first[t]..ev_first[t+1] index events of tick t */
    const int32_t *ev_first, *ev_vehicle; const int64_t *ev_off; const int32_t *ev_cells;
    /* vehicle state [n_vehicles] */
    int32_t *pos, *path_off, *path_len, *steps;
    int8_t *alive, *base_speed, *cur_speed, *max_steps, *early, *is_stuck, *prev_valid, *malfunction, *direction;
    int16_t *stuck_ticks; int32_t *stranded;
    /* light tables: CSR light -> cells (own cell first, then controlled cells) */
    const int32_t *tl_off, *tl_cells;
    /* CSR group -> all lights / NS lights / EW lights (light indices), lane cells, cluster cells */
    const int32_t *g_all_off, *g_all, *g_ns_off, *g_ns, *g_ew_off, *g_ew;
    const int32_t *g_nsin_off, *g_nsin, *g_ewin_off, *g_ewin, *g_cl_off, *g_cl;
    /* group state [n_groups] */
    int32_t *g_cur, *g_pend, *g_qt, *g_gap, *g_last, *g_ft_phase, *g_ft_timer;
    /* sideswipes: is_in_collision per vehicle, vehicle standing on a cell (-1: none) */
    int8_t *collision; int32_t *veh_at;
    /* pressure controller: lane cells on the far side of their light (ns_out_coords / ew_out_coords :141-154), last pressures */
    const int32_t *g_nsout_off, *g_nsout, *g_ewout_off, *g_ewout;
    int32_t *g_nsp, *g_ewp;
    /* neighbour controllers: neighbor_groups of every group, [n_groups][4] = N, S, E, W, -1 none (populate_links :175-242) */
    const int32_t *g_nbr;
} vsim;

enum { MIN_GREEN = 5, MAX_GREEN = 30, GAP = 3, GREEN_DURATION = 20, AWARENESS = 10,
       MALFUNCTION_TICKS = 400, COLLISION_TICKS = 600, STUCK_THRESHOLD = 30, RAIN_REDUCTION = 2 };

size_t oracle_vsim_sizeof(void) { return sizeof(vsim); }

static void light_set(vsim *s, int l, int v) { /* cell.py:241-251 */
    for (int k = s->tl_off[l]; k < s->tl_off[l + 1]; k++) s->stop[s->tl_cells[k]] = (uint8_t)v;
}

static void group_step(vsim *s, int g) {
    if (s->g_pend[g] < 0) {
        if (s->algo == 0) { /* run_queue_actuated :463-494 */
            s->g_qt[g]++;
            int ns_q = 0, ew_q = 0;
            for (int k = s->g_nsin_off[g]; k < s->g_nsin_off[g + 1]; k++) ns_q += s->occ[s->g_nsin[k]];
            for (int k = s->g_ewin_off[g]; k < s->g_ewin_off[g + 1]; k++) ew_q += s->occ[s->g_ewin[k]];
            int cur_q = s->g_cur[g] == 0 ? ns_q : ew_q, opp_q = s->g_cur[g] == 0 ? ew_q : ns_q;
            if (s->g_qt[g] == 1) { s->g_last[g] = cur_q; s->g_gap[g] = 0; }
            if (cur_q > s->g_last[g]) { s->g_last[g] = cur_q; s->g_gap[g] = 0; } else s->g_gap[g]++;
            if (s->g_qt[g] >= MIN_GREEN && (s->g_gap[g] >= GAP || s->g_qt[g] >= MAX_GREEN || (opp_q > cur_q && cur_q == 0))) {
                int next = 1 - s->g_cur[g];
                if (next != s->g_cur[g] && next != s->g_pend[g]) s->g_pend[g] = next; /* apply_phase :386-393 */
                s->g_qt[g] = 0;
            }
        } else if (s->algo == 2) { /* run_pressure_control :448-461: every tick without a pending phase, phase of the larger pressure */
            int ns_p = 0, ew_p = 0;
            for (int k = s->g_nsin_off[g]; k < s->g_nsin_off[g + 1]; k++) ns_p += s->occ[s->g_nsin[k]];
            for (int k = s->g_nsout_off[g]; k < s->g_nsout_off[g + 1]; k++) ns_p -= s->occ[s->g_nsout[k]];
            for (int k = s->g_ewin_off[g]; k < s->g_ewin_off[g + 1]; k++) ew_p += s->occ[s->g_ewin[k]];
            for (int k = s->g_ewout_off[g]; k < s->g_ewout_off[g + 1]; k++) ew_p -= s->occ[s->g_ewout[k]];
            s->g_nsp[g] = ns_p; s->g_ewp[g] = ew_p;
            int ph = ns_p > ew_p ? 0 : 1;
            if (ph != s->g_cur[g] && ph != s->g_pend[g]) s->g_pend[g] = ph;
        } else if (s->algo == 3) { /* run_neighbor_green_wave :522-546: follow the neighbours' CURRENT phase (this tick's if they stepped
                                      before this group), else the longer own queue */
            int ns_q = 0, ew_q = 0, favor_ns = 0, favor_ew = 0;
            for (int k = s->g_nsin_off[g]; k < s->g_nsin_off[g + 1]; k++) ns_q += s->occ[s->g_nsin[k]];
            for (int k = s->g_ewin_off[g]; k < s->g_ewin_off[g + 1]; k++) ew_q += s->occ[s->g_ewin[k]];
            for (int d = 0; d < 4; d++) {
                int nb = s->g_nbr[4 * g + d];
                if (nb < 0) continue;
                if (d < 2 && s->g_cur[nb] == 0) favor_ns = 1;
                if (d >= 2 && s->g_cur[nb] == 1) favor_ew = 1;
            }
            int ph = (favor_ns && !favor_ew) ? 0 : (favor_ew && !favor_ns) ? 1 : (ns_q > ew_q ? 0 : 1);
            if (ph != s->g_cur[g] && ph != s->g_pend[g]) s->g_pend[g] = ph;
        } else { /* run_fixed_time :427-441 */
            s->g_ft_timer[g]++;
            if (s->g_ft_timer[g] == 1) {
                int ph = s->g_ft_phase[g];
                if (ph != s->g_cur[g] && ph != s->g_pend[g]) s->g_pend[g] = ph;
            }
            if (s->g_ft_timer[g] >= GREEN_DURATION) { s->g_ft_phase[g] = 1 - s->g_ft_phase[g]; s->g_ft_timer[g] = 0; }
        }
    }
    /* _execute_phase_change :348-384 (transition timers disabled by default, clearance enabled) */
    if (s->g_pend[g] < 0) return;
    int occupied = 0;
    for (int k = s->g_cl_off[g]; k < s->g_cl_off[g + 1]; k++) occupied |= s->occ[s->g_cl[k]];
    if (occupied) {
        for (int k = s->g_all_off[g]; k < s->g_all_off[g + 1]; k++) light_set(s, s->g_all[k], 1);
        return;
    }
    const int32_t *go_off = s->g_pend[g] == 0 ? s->g_ns_off : s->g_ew_off, *go = s->g_pend[g] == 0 ? s->g_ns : s->g_ew;
    const int32_t *st_off = s->g_pend[g] == 0 ? s->g_ew_off : s->g_ns_off, *st = s->g_pend[g] == 0 ? s->g_ew : s->g_ns;
    for (int k = go_off[g]; k < go_off[g + 1]; k++) light_set(s, go[k], 0);
    for (int k = st_off[g]; k < st_off[g + 1]; k++) light_set(s, st[k], 1);
    s->g_cur[g] = s->g_pend[g];
    s->g_pend[g] = -1;
}

static void remove_vehicle(vsim *s, int v) { /* city_model.py:1920-1941 */
    s->occ[s->pos[v]] = 0;
    s->stuckmap[s->pos[v]] = 0;
    if (s->veh_at[s->pos[v]] == v) s->veh_at[s->pos[v]] = -1;
    s->alive[v] = 0;
}

static const int32_t *g_ev_of;   /* vehicle -> index of its last route event of the running tick (built per tick by oracle_ticks_run) */

static void set_collision(vsim *s, int v) { /* _set_collision vehicle_base.py:534-541 */
    s->collision[v] = 1; s->malfunction[v] = 0; s->stranded[v] = COLLISION_TICKS; s->base_speed[v] = 0; s->cur_speed[v] = 0;
}

/* _check_sideswipe_collision vehicle_base.py:567-605: the vehicle to my left, then the one to my right; the first one that is
   moving (as far as its last step_decide knows) in the OPPOSITE direction decides: one draw, collision for both or nothing */
static void check_sideswipe(vsim *s, int v, int fires) {
    static const int LEFT[4] = {3, 0, 1, 2}, RIGHT[4] = {1, 2, 3, 0}, DXv[4] = {0, 1, 0, -1}, DYv[4] = {1, 0, -1, 0};
    const int d = s->direction[v];
    if (d < 0) return;
    const int x = s->pos[v] % s->W, y = s->pos[v] / s->W;
    for (int side = 0; side < 2; side++) {
        const int l = side ? RIGHT[d] : LEFT[d], nx = x + DXv[l], ny = y + DYv[l];
        if (nx < 0 || nx >= s->W || ny < 0 || ny >= s->H) continue;
        const int u = s->veh_at[ny * s->W + nx];
        if (u < 0) continue;
        if (s->cur_speed[u] <= 0 || s->is_stuck[u] || s->collision[u] || s->malfunction[u]) continue;
        if (s->direction[u] != ((d + 2) & 3)) continue;
        if (!fires) return;
        set_collision(s, v);
        set_collision(s, u);
        return;
    }
}

/* returns -1 when a tape contract is violated (vehicle already at its target in phase A) */
static int decide(vsim *s, int v, int t) {
    const int nv = s->n_vehicles;
    s->early[v] = 0;
    if (g_ev_of[v] >= 0) { /* the route this vehicle follows from this tick on: the re-plan the reference made in this step_decide
        (:506-517, :454-504; a recorded tape only holds one for a vehicle that got past the early exits below, so where in
        step_decide it is installed makes no difference to it), or the first route of a vehicle that spawned in the tick before
        (trafficsimulation_b200/replan.py hands those over with the next tick's events); the tick's LAST event of the vehicle */
        const int e = g_ev_of[v];
        s->path_off[v] = (int32_t)s->ev_off[e]; s->path_len[v] = (int32_t)(s->ev_off[e + 1] - s->ev_off[e]);
    }
    if (s->malfunction[v] || s->collision[v]) { /* _tick_stranded :552-565 */
        s->stranded[v]--;
        if (s->stranded[v] <= 0) { s->malfunction[v] = 0; s->collision[v] = 0; s->stranded[v] = 0; }
        if (s->malfunction[v] || s->collision[v]) { s->base_speed[v] = 0; s->cur_speed[v] = 0; s->early[v] = 1; return 0; }
    }
    const int draws = s->malf[(size_t)t * nv + v];
    if (draws & 1) { /* _check_malfunction :608-610 */
        s->malfunction[v] = 1; s->collision[v] = 0; s->stranded[v] = MALFUNCTION_TICKS; s->base_speed[v] = 0; s->cur_speed[v] = 0;
        s->early[v] = 1;
        return 0;
    }
    check_sideswipe(s, v, draws & 2);
    if (s->collision[v]) { s->base_speed[v] = 0; s->cur_speed[v] = 0; s->early[v] = 1; return 0; } /* :633-638 */
    if (s->stop[s->pos[v]] == 1) { s->base_speed[v] = 0; s->cur_speed[v] = 0; s->early[v] = 1; return 0; } /* :639-643 */
    if (s->base_speed[v] == 0) s->base_speed[v] = (int8_t)s->speed[(size_t)t * nv + v]; /* :94-112 */
    int sp = s->base_speed[v];
    if (s->rain_enabled && s->rain[s->pos[v]] == 1) { sp -= RAIN_REDUCTION; if (sp < 1) sp = 1; }
    s->cur_speed[v] = (int8_t)sp;
    /* _scan_ahead_for_obstacles :422-452 */
    int idx_stop = -1, idx_veh = -1, look = s->path_len[v] < AWARENESS ? s->path_len[v] : AWARENESS;
    for (int i = 0; i < look; i++) {
        int c = s->ev_cells[s->path_off[v] + i];
        if (idx_stop < 0 && s->stop[c] == 1) idx_stop = i;
        if (idx_veh < 0 && s->occ[c] == 1) idx_veh = i;
        if (idx_stop == 0 || idx_veh == 0) break;
    }
    int ms = sp < s->path_len[v] ? sp : s->path_len[v]; /* :719-731 */
    if (idx_stop >= 0 && idx_stop < ms) ms = idx_stop;
    if (idx_veh >= 0 && idx_veh < ms) ms = idx_veh;
    s->max_steps[v] = (int8_t)ms;
    if (ms <= 0) {
        s->base_speed[v] = 0;
        if (s->pos[v] == s->target[v]) return -1;
        s->early[v] = 1;
    }
    return 0;
}

static void vehicle_step(vsim *s, int v) { /* :666-685 */
    if (!s->early[v]) {
        int old = s->pos[v];
        for (int k = 0; k < s->max_steps[v]; k++) { /* _execute_movement :733-753 */
            if (s->path_len[v] == 0) break;
            int c = s->ev_cells[s->path_off[v]];
            if (s->occ[c] == 1 && c != s->pos[v]) break;
            if (s->stop[c] == 1 && k != s->max_steps[v] - 1) break;
            /* _move_to :521-532 -> move_vehicle city_model.py:1945-1963 */
            s->occ[old] = 0; s->occ[c] = 1;
            s->stuckmap[old] = 0; s->stuckmap[c] = s->is_stuck[v] ? 1 : 0;
            if (s->veh_at[old] == v) s->veh_at[old] = -1;
            s->veh_at[c] = v;
            s->pos[v] = c;
            int d = c - old;
            if (d == s->W) s->direction[v] = 0; else if (d == 1) s->direction[v] = 1;
            else if (d == -s->W) s->direction[v] = 2; else if (d == -1) s->direction[v] = 3;
            if (s->stuck_ticks[v] > 0) { s->is_stuck[v] = 0; s->stuck_ticks[v] = 0; }
            s->steps[v]++;
            old = c;
            s->path_off[v]++; s->path_len[v]--;
        }
        s->prev_valid[v] = 1;
    } else {
        s->early[v] = 0;
        if (s->prev_valid[v] && s->stop[s->pos[v]] != 1) { /* tick_stuck :687-693 */
            s->stuck_ticks[v]++;
            if (s->stuck_ticks[v] > STUCK_THRESHOLD && !s->is_stuck[v]) s->is_stuck[v] = 1;
        }
    }
    if (s->pos[v] == s->target[v]) remove_vehicle(s, v); /* on_target_reached :755-775 */
}

static const int32_t *g_rank_row;
static int cmp_rank(const void *a, const void *b) {
    int x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (g_rank_row[x] > g_rank_row[y]) - (g_rank_row[x] < g_rank_row[y]);
}

/* runs `n` ticks; returns 0, or -(tick+1) on a tape-contract violation */
int oracle_ticks_run(vsim *s, int n) {
    int32_t *order = malloc((size_t)(s->n_vehicles + 1) * sizeof(int32_t));
    int32_t *ev_of = malloc((size_t)(s->n_vehicles + 1) * sizeof(int32_t));   /* vehicle -> its last route event of the tick, -1: none */
    for (int v = 0; v < s->n_vehicles; v++) ev_of[v] = -1;
    g_ev_of = ev_of;
    for (int it = 0; it < n; it++) {
        const int t = s->tick;
        int na = 0;
        for (int e = s->ev_first[t]; e < s->ev_first[t + 1]; e++) ev_of[s->ev_vehicle[e]] = e;
        for (int v = 0; v < s->n_vehicles; v++) /* phase A: run_parallel_decide, one worker, list order */
            if (s->alive[v]) { if (decide(s, v, t) < 0) { free(order); free(ev_of); return -(t + 1); } order[na++] = v; }
        for (int g = 0; g < s->n_groups; g++) group_step(s, g); /* phase B: light groups first */
        g_rank_row = s->rank + (size_t)t * s->n_vehicles;
        qsort(order, (size_t)na, sizeof(int32_t), cmp_rank);
        for (int i = 0; i < na; i++) vehicle_step(s, order[i]);
        for (int v = 0; v < s->n_vehicles; v++) { /* the tape-driven spawner, last */
            if (s->spawn_tick[v] != t) continue;
            if (s->occ[s->origin[v]] == 1) continue; /* dropped attempt */
            s->alive[v] = 1; s->pos[v] = s->origin[v];
            s->occ[s->origin[v]] = 1; s->stuckmap[s->origin[v]] = 0; /* place_vehicle :1897-1908 */
            s->veh_at[s->origin[v]] = v;
            s->path_off[v] = 0; s->path_len[v] = 0;
            if (g_ev_of[v] >= 0) {
                const int e = g_ev_of[v];
                s->path_off[v] = (int32_t)s->ev_off[e]; s->path_len[v] = (int32_t)(s->ev_off[e + 1] - s->ev_off[e]);
            }
        }
        for (int e = s->ev_first[t]; e < s->ev_first[t + 1]; e++) ev_of[s->ev_vehicle[e]] = -1;
        s->tick++;
    }
    free(order); free(ev_of);
    return 0;
}
