/*
 * tsim.h -- C ABI of the B200-native CityModel hot path (libtsim.so).
 *
 * Drop-in boundary.  The reference has no plugin point for this
 * path; its seam is structural:
 *   - layout:  the pass methods called from CityModel.__init__, Simulation/city_model.py:125-139
 *              (each `tsim_layout_*` entry cites the method it replaces), whose only mutator is
 *              place_cell (city_model.py:1864-1870) and only reader get_cell_contents (:1965-1972);
 *   - maps:    _build_simple_maps, city_model.py:2151-2199;
 *   - tick:    CityModel.step (city_model.py:1831-1860) -> VehicleAgent.step_decide/step
 *              (agents/vehicles/vehicle_base.py:616-685) and IntersectionLightGroup.step
 *              (agents/city_structure_entities/intersection_light_group.py:396-423).
 * The only FFI precedent upstream is the pybind11 module utilities/pathfinding/astar_cpp.cpp:117-128
 * (flat C-contiguous [y,x] arrays, borrowed pointers); this ABI keeps those conventions and drops
 * its silent dtype casts and silent CPU fallback.
 *
 * Conventions
 *   - every plane is row-major [H][W] with y the slow axis (y up, N = +y, config.py:64);
 *   - all plane / table pointers are DEVICE pointers owned by the caller (torch tensors);
 *     the library allocates nothing persistent and keeps no global state;
 *   - every entry point is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - every entry point returns a tsim_status; tsim_last_error() gives a thread-local message;
 *   - results that the host must read (counts, flags) are written to caller-provided device
 *     memory; the caller synchronises and reads them.
 */
#ifndef TSIM_H
#define TSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSIM_ABI_VERSION 5

/* cell_type codes = index into Defaults.ZONES (Simulation/config.py:74-95) */
enum tsim_cell_type {
    TSIM_RESIDENTIAL = 0, TSIM_OFFICE, TSIM_MARKET, TSIM_LEISURE, TSIM_OTHER, TSIM_EMPTY,
    TSIM_NOTHING, TSIM_SIDEWALK, TSIM_WALL, TSIM_R1, TSIM_R2, TSIM_R3, TSIM_INTERSECTION,
    TSIM_HIGHWAY_ENTRANCE, TSIM_HIGHWAY_EXIT, TSIM_TRAFFIC_LIGHT, TSIM_TRAFFIC_LIGHT_STOP,
    TSIM_CONTROLLED_ROAD, TSIM_CONTROLLED_ROAD_STOP, TSIM_BLOCK_ENTRANCE
};

/* direction indices; the bit (1 << index) is the allowed_dirs_map bit (city_model.py:2191-2196) */
enum tsim_dir { TSIM_N = 0, TSIM_E = 1, TSIM_S = 2, TSIM_W = 3 };

/* `dirs` plane, u16: bits0-3 mask, bits 4+2i..5+2i i-th entry of the ordered list, bits12-14 length */
/* `aux`  plane, u8 : */
#define TSIM_AUX_ORIG_MASK 0x1f /* original cell_type of a ControlledRoad (CellAgent.road_type) */
#define TSIM_AUX_RING      0x20 /* member of CityModel._ring_road_cells (forced ring corners)   */
#define TSIM_AUX_EVER_INT  0x40 /* member of CityModel._intersection_cells                       */
#define TSIM_AUX_HAS_LIGHT 0x80 /* CellAgent.light is not None                                   */

typedef enum tsim_status {
    TSIM_OK = 0,
    TSIM_ERR_CONFIG = 1,      /* invalid configuration / null pointer / misaligned plane */
    TSIM_ERR_WORKSPACE = 2,   /* workspace too small (see tsim_workspace_bytes)          */
    TSIM_ERR_CUDA = 3,        /* a CUDA call failed; message in tsim_last_error()        */
    TSIM_ERR_TAPE = 4,        /* a tape is too short / holds an illegal decision         */
    TSIM_ERR_UNSUPPORTED = 5, /* configuration not implemented on the GPU path           */
    TSIM_ERR_CAPACITY = 6     /* a bounded device structure overflowed                   */
} tsim_status;

/* POD copy of the CityModel constructor kwargs the passes read (city_model.py:27-53) */
typedef struct tsim_cfg {
    int32_t width, height;
    int32_t wall_thickness, sidewalk_ring_width;
    int32_t ring_road_type;                    /* 0 None, 1 R1, 2 R2, 3 R3 */
    int32_t optimized_intersections;
    int32_t subblock_roads_have_intersections;
    int32_t subblock_road_type;                /* 1..3 */
    int32_t min_subblock_spacing;
    int32_t traffic_light_range;
    int32_t forward_traffic_light_range;
    int32_t forward_intersections_mode;
    int32_t block_entrance_road_level;         /* Defaults.BLOCK_ENTRANCE_ROAD_LEVEL, config.py:26 */
    /* row-band shard window: the planes passed with this cfg hold global rows
       [win_y0, win_y0 + win_rows) of the width x height grid, row-major, nothing else.  Every cell
       index a call takes or returns (component roots, light cells, link tables) is an index into THIS
       window; rows outside it do not exist for the call.  Single device: win_y0 = 0, win_rows = height. */
    int32_t win_y0, win_rows;
    int32_t win_halo;   /* rows at each CUT end of the window that belong to the neighbour shard (0 on a single device) */
} tsim_cfg;

typedef struct tsim_planes {
    uint8_t  *cell_type;
    uint16_t *dirs;
    uint8_t  *aux;
    int32_t  *block_id;
} tsim_planes;

/* per-row / per-column band descriptors (device, uint32 each), built by tsim_build_line_table */
typedef struct tsim_lines {
    const uint32_t *row;   /* [height] */
    const uint32_t *col;   /* [width]  */
    /* optional class tables of tsim_build_class_tables (all NULL / 0 = closed-form evaluation per cell) */
    const uint8_t  *row_class;   /* [height] class of the descriptor triple of rows y-1, y, y+1 */
    const uint8_t  *col_class;   /* [width]                                                     */
    const uint32_t *lut;         /* [n_row_classes][n_col_classes]: type | aux << 8 | dirs << 16 */
    int32_t n_row_classes, n_col_classes;
    /* optional row patterns of tsim_build_row_patterns: the whole bulk part of a row of class k, ready to be copied */
    const uint8_t  *pat_type;    /* [n_row_classes][width] */
    const uint16_t *pat_dirs;    /* [n_row_classes][width] */
    const uint8_t  *pat_aux;     /* [n_row_classes][width] */
} tsim_lines;

/* component table produced by the labelling passes, one row per component of the window in raster
   discovery order: minx, miny (global y), maxx, maxy (global y), size, root cell (window index) */
#define TSIM_BLOB_STRIDE 6

typedef struct tsim_blobs {
    int32_t *table;          /* [cap][TSIM_BLOB_STRIDE], device                                          */
    int32_t  cap;
    int32_t *count;          /* device scalar: components found in the window                            */
    const int32_t *id_base;  /* device scalar or NULL (= 0): id of table row k is k + 1 + *id_base.  Shards
                                set it from the all-gathered root counts so that ids -- the index into
                                every per-block tape and the value stored in block_id -- are the global
                                raster discovery ranks                                     */
} tsim_blobs;

int         tsim_version(void);
const char *tsim_last_error(void);
/* number of CUDA kernels this library has launched in this process (benchmarks report it) */
long long   tsim_launch_count(void);

/* host-side helper: band list (n rows of start,end,type,dir) -> line table of `len` entries.
   Restates _find_band_covering (city_model.py:1269-1273: first band in list order wins) and the
   forced-band membership used by _override_corner_lane_dirs (:519-527) and
   _upgrade_r2_to_intersections (:859-866). */
tsim_status tsim_build_line_table(const int32_t *bands, int32_t n_bands, int32_t len, uint32_t *out_host);

/* host-side helper: away from the frame every output of tsim_layout_frame_roads is a function of the
   band descriptors of the cell's row/column and of their two neighbours only.  This groups rows and
   columns into classes of equal descriptor triples and tabulates the cell for every (row class, column
   class) pair, so the kernel's bulk path is one table look-up per cell.  TSIM_ERR_CAPACITY when an axis
   has more than 255 classes or the table exceeds lut_cap entries (the caller then passes no tables). */
tsim_status tsim_build_class_tables(const tsim_cfg *cfg, const uint32_t *row_host, const uint32_t *col_host,
                                    uint8_t *row_class_host, uint8_t *col_class_host, uint32_t *lut_host,
                                    int32_t lut_cap, int32_t *n_row_classes, int32_t *n_col_classes);

/* host-side helper: away from the frame, all rows of one class are IDENTICAL (the cell is a function of the row class and
   of x).  Writes, for every row class, that row's cells for all x (valid where x lies in the bulk; the kernel evaluates
   the frame columns itself), so the bulk of tsim_layout_frame_roads becomes a copy of ~n_row_classes cache-resident rows. */
tsim_status tsim_build_row_patterns(const tsim_cfg *cfg, const uint32_t *row_host, const uint32_t *col_host,
                                    const uint8_t *row_class_host, int32_t n_row_classes, uint8_t *pat_type_host,
                                    uint16_t *pat_dirs_host, uint8_t *pat_aux_host);

/* bytes of device workspace every tsim_layout_* / tsim_maps call may use */
tsim_status tsim_workspace_bytes(const tsim_cfg *cfg, size_t *out_bytes);

/* _place_thick_wall + _place_sidewalk_inner_ring + _clear_interior (city_model.py:315-369) and
   _build_roads_and_sidewalks from the band lists on (:396-495, incl. _make_intersection :211-306,
   _compute_lane_dirs :1275-1368, _override_corner_lane_dirs :498-558,
   _replace_boundary_highways_with_entrances :1370-1420), fused: writes cell_type, dirs, aux, block_id */
tsim_status tsim_layout_frame_roads(const tsim_cfg *cfg, const tsim_planes *p, const tsim_lines *lines,
                                    void *stream);

/* 4-connected components of `Nothing` in raster discovery order (the flood fills at
   city_model.py:632-647 and :746-763): fills the component table and the count.  The run structure
   of the labelling stays in `workspace`; tsim_layout_zones (same workspace, no other libtsim call in
   between) turns it into the block_id plane.  block_id itself is not touched here. */
tsim_status tsim_layout_label_nothing(const tsim_cfg *cfg, const tsim_planes *p, const tsim_blobs *blobs,
                                      int32_t *err_flag, void *workspace, size_t ws_bytes, void *stream);

/* shard bookkeeping after a labelling call: with this device owning the global rows [own_lo, own_hi) of
   its window, out (device int32[3]) receives the number of component roots below the own rows, the number inside them,
   and 1 if a component meets the own rows AND a cut edge of the window (halo too small). */
tsim_status tsim_shard_counts(const tsim_cfg *cfg, const tsim_blobs *blobs, int32_t own_lo, int32_t own_hi, int32_t *out,
                              void *stream);

/* position-weighted 64-bit digest of the LOCAL rows [row_lo, row_hi) of the planes selected by `what` (1 cell_type, 2 dirs, 4 aux,
   8 block_id), ADDED to *out (device).  The digest of a row depends on its global position (win_y0 + row), so two shards
   that hold the same global rows get the same value exactly when the bytes agree: what row-band shards compare instead of
   exchanging their halos after every pass. */
tsim_status tsim_rows_digest(const tsim_cfg *cfg, const tsim_planes *p, int32_t row_lo, int32_t row_hi, int32_t what, uint64_t *out,
                             void *stream);

/* measurement aid, not part of the path (profiles/write_peak.py): a write-only kernel storing hashed (incompressible) words to
   `streams` planes of n_bytes each starting at `base`, `blocks` CTAs of 256 threads, grid-stride */
tsim_status tsim_debug_write_probe(void *base, long long n_bytes, int32_t streams, int32_t blocks, void *stream);

/* _carve_subblock_roads (city_model.py:563-737) given the table of tsim_layout_label_nothing and
   the carve tape: one row of 8 int32 per blob id (drawn, carved, px, py (global), hor_dir, ver_dir,
   inbound_is_horizontal, tries).  err_flag (device int32) is set non-zero on an illegal row. */
tsim_status tsim_layout_carve(const tsim_cfg *cfg, const tsim_planes *p, const tsim_lines *lines,
                              const tsim_blobs *blobs, const int32_t *tape, int32_t n_tape,
                              int32_t *err_flag, void *stream);

/* _flood_fill_blocks_storing_data (city_model.py:742-806) right after tsim_layout_label_nothing:
   writes block_id (id of the cell's block, 0 elsewhere) and fills each block with Empty (bbox < 3)
   or the zone zone_by_block[id-1]. */
tsim_status tsim_layout_zones(const tsim_cfg *cfg, const tsim_planes *p, const tsim_blobs *blobs,
                              const uint8_t *zone_by_block, int32_t n_tape, int32_t *err_flag,
                              void *workspace, size_t ws_bytes, void *stream);

/* _eliminate_dead_ends (city_model.py:811-840); *sweeps (device) receives the sweep count */
tsim_status tsim_layout_dead_ends(const tsim_cfg *cfg, const tsim_planes *p, int32_t *sweeps,
                                  void *workspace, size_t ws_bytes, void *stream);

/* _upgrade_r2_to_intersections (city_model.py:842-879) */
tsim_status tsim_layout_upgrade_r2(const tsim_cfg *cfg, const tsim_planes *p, const tsim_lines *lines,
                                   int32_t *err_flag, void *stream);

/* _final_place_block_entrances (city_model.py:884-963); run_by_block[id-1] = canonical index of the
   chosen run among the longest ones; entrances[k] (k = table row) receives the cell index or -1 */
tsim_status tsim_layout_entrances(const tsim_cfg *cfg, const tsim_planes *p, const tsim_blobs *blobs,
                                  const int32_t *run_by_block, int32_t n_tape, int32_t *entrances,
                                  int32_t *err_flag, void *workspace, size_t ws_bytes, void *stream);

/* _remove_invalid_intersection_directions + _add_entrance_directions (city_model.py:969-1070), fused */
tsim_status tsim_layout_fix_dirs(const tsim_cfg *cfg, const tsim_planes *p, void *stream);

/* link tables of _add_traffic_lights: CSR by light, lights in ascending cell index */
typedef struct tsim_light_links {
    int32_t *n_lights;        /* device scalar                                          */
    int32_t *light_cell;      /* [cap_lights]                                            */
    int32_t *ctrl_off;        /* [cap_lights + 1]  controlled road cells per light       */
    int32_t *ctrl_cell;       /* [cap_ctrl]                                              */
    int32_t *inc_off;         /* [cap_lights + 1]  assigned incoming lane cells (multiset) */
    int32_t *inc_cell;        /* [cap_inc]                                               */
    int32_t cap_lights, cap_ctrl, cap_inc;
    /* assigned OUTGOING cells (forward_traffic_light_range, city_model.py:1550-1584); may be NULL / 0 when the
       option is off */
    int32_t *out_off;         /* [cap_lights + 1]                                        */
    int32_t *out_cell;        /* [cap_out]                                               */
    int32_t cap_out;
} tsim_light_links;

/* _add_traffic_lights (city_model.py:1422-1548, CellAgent.leads_to cell.py:201-227) */
tsim_status tsim_layout_lights(const tsim_cfg *cfg, const tsim_planes *p, const tsim_light_links *links,
                               int32_t *err_flag, void *workspace, size_t ws_bytes, void *stream);

/* The stages of tsim_layout_lights, for row-band shards: `leads_to` needs reachability over the WHOLE grid, so
   shards run prepare, agree on one pivot, then alternate tsim_lights_reach with an exchange (bitwise OR) of the
   halo rows of the two reachability planes until no shard reports a change, then finish.
     prepare : bit-planes + candidate list; pivot_out (device int32[2], optional) receives this window's pivot
               candidates as window cell indices: [0] first intersection at or after the grid's middle row,
               [1] first intersection at all (0x7fffffff = none)
     seed    : pivot = device scalar with the window cell index of the agreed pivot (negative: outside this
               window); NULL = use this window's own candidate
     reach   : closure inside the window; *changed (device, optional) is set to 1 when a bit was added.
               edge_rows < 0: first call after seed; edge_rows >= 0: resumed call, only that many rows at each end of
               the window were changed from outside since the previous call (the halo exchange)
     reach_planes : byte offsets of the planes inside the workspace ([win_rows][words_per_row] uint64)
     finish  : evaluation, light numbering, link tables, cell conversion */
tsim_status tsim_lights_prepare(const tsim_cfg *cfg, const tsim_planes *p, int32_t *pivot_out, int32_t *err_flag,
                                void *workspace, size_t ws_bytes, void *stream);
tsim_status tsim_lights_seed(const tsim_cfg *cfg, const int32_t *pivot, void *workspace, size_t ws_bytes, void *stream);
tsim_status tsim_lights_reach(const tsim_cfg *cfg, int32_t edge_rows, int32_t *changed, int32_t *err_flag, void *workspace,
                              size_t ws_bytes, void *stream);
tsim_status tsim_lights_reach_planes(const tsim_cfg *cfg, size_t ws_bytes, size_t *fw_off, size_t *bw_off,
                                     int32_t *words_per_row);

tsim_status tsim_lights_finish(const tsim_cfg *cfg, const tsim_planes *p, const tsim_light_links *links,
                               int32_t *err_flag, void *workspace, size_t ws_bytes, void *stream);

/* tsim_layout_lights = tsim_lights_prepare + tsim_lights_eval + tsim_lights_links, for callers (bench.py) that time the stages:
     eval  : every candidate's record.  `leads_to` beyond the lane itself is settled by local searches (Z-shaped witnesses, then
             closures of a 128 x 128 window); the reachability planes of the window are only closed, inside this call, when a
             query is left over
     links : light numbering, link tables, cell conversion (the second half of tsim_lights_finish) */
tsim_status tsim_lights_eval(const tsim_cfg *cfg, const tsim_planes *p, const tsim_light_links *links, int32_t *err_flag,
                             void *workspace, size_t ws_bytes, void *stream);
tsim_status tsim_lights_links(const tsim_cfg *cfg, const tsim_planes *p, const tsim_light_links *links, int32_t *err_flag,
                              void *workspace, size_t ws_bytes, void *stream);

/* _build_simple_maps (city_model.py:2151-2199) */
tsim_status tsim_maps(const tsim_cfg *cfg, const tsim_planes *p, uint8_t *is_road, uint8_t *road_type,
                      uint8_t *intersection, uint8_t *allowed_dirs, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Tick: CityModel.step (city_model.py:1831-1860) = phase A VehicleAgent.step_decide for every active
 * vehicle (vehicle_base.py:616-663), then phase B in activation order: IntersectionLightGroup.step
 * (intersection_light_group.py:396-423, QUEUE_ACTUATED :463-494 / FIXED_TIME :427-441 / PRESSURE_CONTROL :448-461 / NEIGHBOR_GREEN_WAVE :522-546, phase commit
 * :348-384, stop_map writes cell.py:241-251), VehicleAgent.step (vehicle_base.py:666-685: movement
 * :733-753 via CityModel.move_vehicle city_model.py:1945-1963, tick_stuck :687-693, arrival :755-775),
 * and the tape-driven spawner (place_vehicle city_model.py:1897-1908).
 * Activation order, speed / malfunction draws, spawns and planned routes are INPUT tapes.
 * ---------------------------------------------------------------------------------------------- */
typedef struct tsim_light_tables {   /* all device pointers; CSR offsets have n+1 entries */
    int32_t n_groups, n_lights;
    const int32_t *tl_off, *tl_cells;          /* light -> its own cell followed by its controlled road cells */
    const int32_t *g_all_off, *g_all;          /* group -> lights (indices)                                  */
    const int32_t *g_ns_off, *g_ns;            /* group -> lights of the N-S axis (opposite_pairs["N-S"])    */
    const int32_t *g_ew_off, *g_ew;            /* group -> lights of the W-E axis                            */
    const int32_t *g_nsin_off, *g_nsin;        /* group -> ns_in_coords cells (multiset)                     */
    const int32_t *g_ewin_off, *g_ewin;        /* group -> ew_in_coords cells (multiset)                     */
    const int32_t *g_cl_off, *g_cl;            /* group -> intersection cluster cells                        */
    /* PRESSURE_CONTROL only (NULL otherwise): the lane cells on the far side of their light, ns_out_coords / ew_out_coords
       (intersection_light_group.py:141-154).  NOTE for that controller all four lane lists hold the cells the reference
       actually reads: run_pressure_control (:448-461) indexes the occupancy map reshaped to (-1, 2) (`to_int32`, :443-446),
       i.e. flat element 2 * y + x instead of W * y + x -- the host tables carry that index (light_groups.pressure_cells).  */
    const int32_t *g_nsout_off, *g_nsout;
    const int32_t *g_ewout_off, *g_ewout;
    /* NEIGHBOR_GREEN_WAVE only (NULL otherwise): neighbor_groups of every group, [n_groups][4] = N, S, E, W, -1 = none
       (IntersectionLightGroup.populate_links, intersection_light_group.py:175-242; host: light_groups.neighbor_links)     */
    const int32_t *g_nbr;
} tsim_light_tables;

typedef struct tsim_tick_tapes {     /* all device pointers */
    int32_t n_ticks, n_vehicles;
    const int32_t *spawn_first;      /* [n_ticks+1] vehicles (= spawn attempts) are sorted by spawn tick      */
    const int32_t *origin, *target;  /* [n_vehicles] cell indices                                             */
    const uint8_t *speed;            /* [n_ticks][n_vehicles] value of random.randint(1,5) if drawn           */
    const uint8_t *malfunction;      /* [n_ticks][n_vehicles] bit 0 = the malfunction draw fires (vehicle_base.py:608-610);
                                        bit 1 = the sideswipe draw fires IF the vehicle gets to make it (:567-605; live-list
                                        kernel only, the vehicle-indexed one raises error flag 34).  After a run
                                        state.malfunction[v] bit 1 (tsim_tick_export) = is_in_collision               */
    const int32_t *rank;             /* [n_ticks][n_vehicles] activation rank (lower steps first)             */
    const int32_t *ev_first;         /* [n_ticks+1] route events sorted by tick                               */
    const int32_t *ev_vehicle;       /* [n_events]                                                            */
    const int64_t *ev_off;           /* [n_events+1] into ev_cells                                            */
    const int32_t *ev_cells;         /* route arena: cell indices, next cell first                            */
    const uint8_t *rain_map;         /* [H*W] or NULL (Defaults.RAIN_ENABLED False)                           */
} tsim_tick_tapes;
/* Route events.  Event e of tick t (ev_first[t] <= e < ev_first[t+1]) makes vehicle ev_vehicle[e] follow
   ev_cells[ev_off[e] .. ev_off[e+1]) from phase A of tick t on (a spawn of tick t takes it as its first route): every return of
   VehicleAgent._compute_path (vehicle_base.py:143-167).  The kernels read the four arrays anew at every tsim_tick_run, and a live
   vehicle keeps POINTING into ev_cells; so a caller that plans the routes itself between two ticks (trafficsimulation_b200/replan.py:
   the re-plan triggers and the planner of vehicle_base.py:143-517 around tsim_astar_batch) rewrites ev_first[t], ev_first[t+1..],
   ev_vehicle and ev_off before it runs tick t and APPENDS the new routes to ev_cells -- or, when that arena is full, starts it again
   with an event for the remaining route of every live vehicle.                                                                    */

typedef struct tsim_tick_state {     /* all device pointers, owned by the caller */
    uint8_t *occupancy, *stop_map, *stuck_map;   /* [H*W] CityModel.occupancy_map / stop_map / stuck_map      */
    uint64_t *claim;                             /* [2][H*W] generation-tagged claim words (scratch), zeroed by tsim_tick_init */
    int32_t *stopw;                              /* [H*W] staged stop_map writes (scratch), zeroed by tsim_tick_init          */
    /* vehicle SoA [n_vehicles] */
    int32_t *pos, *path_len, *steps, *stranded;
    int64_t *path_off;
    int16_t *stuck_ticks;
    int8_t *alive, *base_speed, *cur_speed, *max_steps, *early, *is_stuck, *prev_valid, *malfunction, *direction, *moved;
    /* light-group state [n_groups] */
    int32_t *g_cur, *g_pend, *g_qt, *g_gap, *g_last, *g_ft_phase, *g_ft_timer, *g_plan;
    int32_t *scalars;                            /* [16]: [0] next tick, [1] error flag, [2] fixed-point iterations (sum),
                                                    [6..7] vehicle updates (int64: sum over ticks of the live vehicles on
                                                    own rows), [9] shard-exchange error flag, rest internal           */
    int32_t own_row_lo, own_row_hi;              /* LOCAL rows of the window this shard owns; 0, 0 = the whole window          */
    /* Working set of the LIVE-LIST kernel (all NULL: the vehicle-indexed kernel, which row-band shards use).  With these
       the tick iterates a compacted list of the live vehicles instead of the attempt array: a vehicle is a 48-byte record
       in `recs` that moves to a new slot every tick (warp-aggregated append into the other half), its plan for the tick
       a 32-byte record in `plans`, and what a vehicle asks about the cells ahead is answered by BIT PLANES over 8 x 8-cell
       tiles in `probe` (occupancy, stop, "planned by one vehicle", "planned by several", staged stop, sideswipe query; one
       64-bit word per tile and plane).  A tick then touches neither the vehicle SoA nor the occupancy / stop_map / stuck_map
       byte maps above: tsim_tick_export writes all of them (tsim_tick_init starts from an empty city with every light at go). */
    uint32_t *probe;                             /* tsim_tick_probe_bytes() bytes, 8-byte aligned                              */
    void *recs;                                  /* [2][n_vehicles] x TSIM_TICK_VREC_BYTES                                     */
    void *plans;                                 /* [n_vehicles] x TSIM_TICK_PLAN_BYTES                                        */
    int32_t *ev_stamp, *ev_plen;                 /* [n_vehicles] tick of the vehicle's pending route event, its length         */
    int64_t *ev_poff;                            /* [n_vehicles] ... and its offset in ev_cells                                */
    /* vehicle-indexed kernel on a shard (optional, NULL = scan the whole attempt array): [n_vehicles] indices of the vehicles
       alive in THIS window, rebuilt by tsim_tick_unpack after every halo refresh (scalars[12] = entries, [13] = valid for the
       next tick), so that a shard's tick costs what its own vehicles and ghosts cost, not the fleet of the whole city          */
    int32_t *live_idx;
    /* live-list kernel: scratch lists of a tick (sideswipe candidates, the vehicles in the claim fixed point, sort keys)     */
    int32_t *sort_keys;                          /* [2 * n_vehicles]                                                           */
    /* live-list kernel, optional (NULL = plain append): with it the survivors of a tick are appended TILE BY TILE
       (counting sort by the tile of the new position, tsim_tick_tiles), so that the vehicles a warp handles next tick are
       neighbours on the grid and their cell probes share cache lines -- the cell-sorted vehicle SoA of large fleets.
       `recs` then holds THREE halves of n_vehicles records.                                                                 */
    int32_t *tile_ws;                            /* [2 * n_tiles] vehicles per tile, first slot of the tile                    */
    /* live-list kernel: what the light groups read, tsim_tick_group_ws_bytes() bytes, 8-byte aligned, filled by tsim_tick_init:
       every group's incoming lanes and cluster as (tile, mask) pairs over the occupancy plane of `probe`, and the cells its
       lights control as one flat list, so that a queue count (intersection_light_group.py:463-494) is a handful of popcounts instead of one load per lane cell */
    void *group_ws;
    /* NEIGHBOR_GREEN_WAVE only (NULL otherwise): [3 * n_groups + 4] scratch of the per-tick fixed point that settles, in
       activation order, which phase every group asks for (a group reads the phase its lower-ranked neighbours hold AFTER
       their step of this very tick)                                                                                      */
    int32_t *g_wave;
} tsim_tick_state;
#define TSIM_TICK_VREC_BYTES 48
#define TSIM_TICK_PLAN_BYTES 32

/* zero the maps / vehicle / group state and prepare the scratch planes */
tsim_status tsim_tick_init(const tsim_cfg *cfg, const tsim_light_tables *lt, const tsim_tick_tapes *tp,
                           const tsim_tick_state *st, void *stream);

/* bytes of tsim_tick_state.probe for this grid */
tsim_status tsim_tick_probe_bytes(const tsim_cfg *cfg, long long *bytes);

/* bytes of tsim_tick_state.group_ws for these light tables (reads three table entries back: synchronises) */
tsim_status tsim_tick_group_ws_bytes(const tsim_cfg *cfg, const tsim_light_tables *lt, long long *bytes);

/* advance n ticks (one persistent cooperative launch); algo: 0 QUEUE_ACTUATED, 1 FIXED_TIME, 2 PRESSURE_CONTROL (needs the
   g_nsout / g_ewout tables), 3 NEIGHBOR_GREEN_WAVE (:522-546; needs g_nbr and state.g_wave).  2 and 3 run on whole cities only:
   the cells / neighbour groups they read lie outside a row-band window */
tsim_status tsim_tick_run(const tsim_cfg *cfg, const tsim_light_tables *lt, const tsim_tick_tapes *tp,
                          const tsim_tick_state *st, int32_t n_ticks, int32_t algo, void *stream);

/* debug / profiling: where the live-list kernel's time went since the last reset, in nanoseconds summed over ticks (device
   globaltimer between the grid-wide barriers): ns16[0] decide, [1] sideswipes, [2] sweep 0, [3] later sweeps, [4] move + append
   (sorted append: + ranks), [5] tile scan, [6] (sorted append: record moves +) spawns + light commits; [7] is not a time: vehicles that took part in the claim fixed point;
   [8..13] thread 0's own work inside those phases: decide vehicles, decide light groups, -, spawns, light commits, route events.
   Synchronises the device. */
tsim_status tsim_debug_tick_phases(unsigned long long *ns16, int32_t reset);

/* number of tiles the live-list kernel sorts by for this grid (tiles are 64 x 64 cells, doubled until there are at most 32768) */
tsim_status tsim_tick_tiles(const tsim_cfg *cfg, int32_t *n_tiles);

/* live-list kernel only: scatter the live records into the vehicle SoA of `st` (alive = 0 for everybody else) and write the
   occupancy / stop_map / stuck_map byte maps (from the records and the probe bytes), so that the host reads the same arrays
   whichever kernel ran.  stop_map must be 4-byte aligned. */
tsim_status tsim_tick_export(const tsim_cfg *cfg, const tsim_tick_tapes *tp, const tsim_tick_state *st, void *stream);

/* ---- row-band shards of the tick (SURVEY.md 8e "Vehicle step"; no reference counterpart: the reference is one process).
   A shard runs tsim_tick_run on its WINDOW (tsim_cfg.win_y0 / win_rows / win_halo): every cell index in
   tsim_light_tables, tsim_tick_state and the origin / target / ev_cells tapes is LOCAL to the window
   ((y - win_y0) * width + x); a tape cell outside the window is TSIM_CELL_OUTSIDE.  The vehicle arrays keep
   their global length (a vehicle is the same index on every shard).  After every tick the shards refresh the
   halo through the two calls below and one message per neighbour in between (the caller's transport: NCCL, a copy).                                        */
#define TSIM_CELL_OUTSIDE (-2)
#define TSIM_TICK_REC_WORDS 12      /* int32 words per vehicle record                                          */
#define TSIM_TICK_REC_HEADER 16     /* int32 words before the first record; word 0 = number of records        */
#define TSIM_TICK_GROUP_WORDS 7     /* g_cur, g_pend, g_qt, g_gap, g_last, g_ft_phase, g_ft_timer             */

/* One MESSAGE per neighbour and tick (int32 words): header | cap vehicle records | state of the n_groups light groups
   both shards simulate | `rows` rows of the occupancy, stop and stuck planes.  Sender and receiver agree on cap,
   n_groups and rows; this is its length.                                                                        */
long long tsim_tick_message_words(int32_t width, int32_t cap, int32_t n_groups, int32_t rows);

typedef struct tsim_tick_strips {    /* [0] = neighbour below (lower rows), [1] = neighbour above; LOCAL row ranges */
    int32_t send_lo[2], send_hi[2];      /* own rows the neighbour holds as halo (empty: no neighbour)             */
    int32_t halo_lo[2], halo_hi[2];      /* halo rows owned by that neighbour                                      */
    int32_t verify_lo[2], verify_hi[2];  /* halo rows where this shard's ghosts must equal what the owner sends   */
    int32_t *send_msg[2], *recv_msg[2];  /* device message buffers (NULL: no neighbour)                            */
    int32_t cap;                         /* vehicle records per message                                            */
    const int32_t *g_send[2];            /* local indices of the groups this shard owns and neighbour d simulates */
    int32_t n_g_send[2];
    const int32_t *g_recv[2];            /* local indices of the groups neighbour d owns and this shard simulates */
    const uint8_t *g_verify[2];          /* per received group: 1 = this shard's copy must already be identical   */
    int32_t n_g_recv[2];
} tsim_tick_strips;

/* builds send_msg[d]: vehicles on send rows d, the state of g_send[d], the send rows of the three map planes;
   marks the vehicles on halo rows as awaiting their owner                                                      */
tsim_status tsim_tick_pack(const tsim_cfg *cfg, const tsim_tick_tapes *tp, const tsim_tick_state *st,
                           const tsim_tick_strips *strips, void *stream);

/* installs recv_msg[d] (the neighbour's send_msg): halo rows, group state, the owners' vehicles; drops the ghosts
   nobody sent.  Inside the verify rows everything this shard simulated must equal what arrives, else scalars[9]
   is set (40..45): the halo is too small for the dependency chains of this traffic                             */
tsim_status tsim_tick_unpack(const tsim_cfg *cfg, const tsim_tick_tapes *tp, const tsim_tick_state *st,
                             const tsim_tick_strips *strips, void *stream);

/* ---- batched route planning.  Replaces astar_numba(width, height, start_x, start_y, goal_x, goal_y,
   occupancy_map, stop_map, is_road_map, road_type_map, allowed_dirs_map, respect_awareness, awareness_range, density_map,
   soft_obstacles, ignore_flow, maximum_steps) -> [(x, y)] (utilities/pathfinding/astar_numba.py:240-281; same signature as
   the reference's pybind11 module, utilities/pathfinding/astar_cpp.cpp:35-50) for MANY queries on one set of maps.
   The result is the reference's path cell for cell (same heap, same tie-breaking, same penalties), not just a path of
   equal cost.                                                                                                       */
typedef struct tsim_astar_maps {      /* device pointers, [height][width], values as in city_model.py:109-115 */
    const uint8_t *occupancy, *stop_map, *is_road_map, *road_type_map, *allowed_dirs_map;
    const double *density_map;        /* NULL = 0 everywhere (it only scales the soft vehicle penalty) */
    const uint8_t *spawn_rank;        /* NULL, or per cell: 0 = whoever stands there was on the grid before this tick's spawner ran,
                                         k in 1..127 = the k-th vehicle the spawner placed this tick stands there (127: that one or a
                                         later one).  The spawns of a tick plan one after the other (VehicleAgent.__init__,
                                         vehicle_base.py:72-76, inside dynamic_traffic_generator.py's loop): see spawn_rank_limit */
} tsim_astar_maps;

#define TSIM_ASTAR_RESPECT_AWARENESS 1
#define TSIM_ASTAR_SOFT_OBSTACLES 2
#define TSIM_ASTAR_IGNORE_FLOW 4

typedef struct tsim_astar_query {
    int32_t sx, sy, gx, gy;
    int32_t flags;                    /* TSIM_ASTAR_* */
    int32_t awareness_range;          /* Defaults.VEHICLE_AWARENESS_RANGE = 10 */
    int32_t maximum_steps;            /* 0x7FFFFFFF = unbounded */
    int32_t spawn_rank_limit;         /* an occupied cell whose spawn_rank is above this counts as FREE for this query (0 with no
                                         spawn_rank plane: every occupied cell is occupied).  Query of the k-th spawn of a tick: k */
} tsim_astar_query;

/* every query works on its own dist / came_from / heap arrays (26 bytes per cell), like the reference's wrapper (:264-272) */
tsim_status tsim_astar_scratch_bytes(const tsim_cfg *cfg, int32_t n_queries, size_t *out);

/* path_len[q] = cells of query q's path (0: none, or start == goal), path_cells[q * max_path ...] = cell indices
   y * width + x, first step first, goal last.  err_flag: 50 a path is longer than max_path (path_len[q] = -(needed)),
   51 open list overflow, 52 query outside the grid.                                                              */
tsim_status tsim_astar_batch(const tsim_cfg *cfg, const tsim_astar_maps *maps, const tsim_astar_query *queries,
                             int32_t n_queries, int32_t *path_len, int32_t *path_cells, int32_t max_path,
                             int32_t *err_flag, void *scratch, size_t scratch_bytes, void *stream);

/* CityModel._update_density_map (city_model.py:1764-1778): density[y][x] = occupied cells / road cells of the 21 x 21 window
   around (x, y), bit for bit what the reference gets from scipy.ndimage.uniform_filter in float32.  Either output may be
   NULL; density64 is the same values widened, the form tsim_astar_maps.density_map takes.  scratch: 2 bytes per cell.   */
tsim_status tsim_density_map(const tsim_cfg *cfg, const uint8_t *occupancy, const uint8_t *is_road_map, float *density32,
                             double *density64, void *scratch, size_t scratch_bytes, void *stream);

/* CityModel.rain_map (city_model.py:113) from the clouds of a tick: writes `value` (0 or 1) into every cell of the discs
   {(cx + dx, cy + dy) : dx^2 + dy^2 <= r^2}, discs = device int32 [n][3] = (cx, cy, r) with (cx, cy) = (int(x), int(y)) of the
   cloud in GLOBAL coordinates (RainAgent.step, agents/rain.py:61-72; cells outside the grid / the window are skipped).
   RainManager.step (:154-185) = one call with value 0 on the previous tick's discs, one with value 1 on this tick's. */
tsim_status tsim_rain_discs(const tsim_cfg *cfg, const int32_t *discs, int32_t n_discs, int32_t value, uint8_t *rain_map, void *stream);

/* labels the 4-connected components of mask == 1 (u8 plane) in raster discovery order: component
   table as tsim_layout_label_nothing, plus the label plane (id, 0 elsewhere).  Used for the
   intersection clusters of _create_intersection_light_groups (city_model.py:1587-1650). */
tsim_status tsim_label_mask(const tsim_cfg *cfg, const uint8_t *mask, int32_t *labels, const tsim_blobs *blobs,
                            int32_t *err_flag, void *workspace, size_t ws_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TSIM_H */
