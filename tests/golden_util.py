"""Loading of the committed reference fixtures (tests/golden/*.npz)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def layout_fixtures():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "layout_*.npz")))


def load(path):
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(bytes(d["meta"]).decode())
    d["name"] = os.path.basename(path)[len("layout_"):-4]
    return d


PLANES = ("cell_type", "dirs", "aux", "block_id")
MAPS = ("is_road_map", "road_type_map", "intersection_map", "allowed_dirs_map")


def tick_fixtures():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "ticks_*.npz")))


def load_ticks(path):
    """Tick fixture with route events decoded back to cell lists and bit-packed tapes unpacked."""
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(bytes(d["meta"]).decode())
    d["name"] = os.path.basename(path)[len("ticks_"):-4]
    W, H, nv = d["meta"]["W"], d["meta"]["H"], d["meta"]["n_attempts"]
    d["W"], d["H"], d["n_ticks"] = W, H, d["meta"]["n_ticks"]
    d["malfunction"] = np.unpackbits(d["malfunction"], axis=1)[:, :nv]
    if "sideswipe" in d:   # bit 1 of a tape entry: the sideswipe draw of that (tick, vehicle) fires
        d["malfunction"] = d["malfunction"] | (np.unpackbits(d.pop("sideswipe"), axis=1)[:, :nv] << 1)
    d["rain_map"] = np.unpackbits(d["rain_map"], axis=1)[:, :W]
    ev_len = d["ev_len"]
    off = np.zeros(len(ev_len) + 1, np.int64)
    off[1:] = np.cumsum(ev_len)
    cells = np.zeros(int(off[-1]), np.int32)
    delta = np.array([W, 1, -W, -1], np.int32)
    steps, sp = d["ev_steps"], 0
    for i, n in enumerate(ev_len):
        if n == 0:
            continue
        seg = np.empty(n, np.int32)
        seg[0] = d["ev_first"][i]
        if n > 1:
            seg[1:] = seg[0] + np.cumsum(delta[steps[sp:sp + n - 1]])
            sp += n - 1
        cells[off[i]:off[i + 1]] = seg
    d["ev_off"], d["ev_cells"] = off, cells
    groups = []
    ng = len(d["g_cluster_off"]) - 1
    for g in range(ng):
        groups.append({f: d["g_" + f][d["g_" + f + "_off"][g]:d["g_" + f + "_off"][g + 1]]
                       for f in ("cluster", "lights", "ns_lights", "ew_lights", "ns_in", "ew_in", "ns_out", "ew_out") if "g_" + f in d})
    if "g_nbr" in d:
        for g in range(ng):
            groups[g]["nbr"] = d["g_nbr"][g]
    d["groups"] = groups
    d["algo"] = d["meta"]["case"].get("algo") or "QUEUE_ACTUATED"
    d["group_state"] = d["group_state"].astype(np.int32)
    return d


def compare_tick(t, got, r):
    """got: dict(pos, base_speed, stuck_ticks, vflags, occ, stop, stuckmap, groups) after tick t."""
    want_pos = r["pos"][t]
    bad = np.flatnonzero(got["pos"] != want_pos)
    assert len(bad) == 0, (t, "pos", [(int(v), int(got["pos"][v]), int(want_pos[v])) for v in bad[:6]])
    alive = want_pos >= 0
    for k in ("base_speed", "stuck_ticks", "vflags"):
        bad = np.flatnonzero((got[k] != r[k][t]) & alive)
        assert len(bad) == 0, (t, k, [(int(v), int(got[k][v]), int(r[k][t][v])) for v in bad[:6]])
    for k in ("occ", "stop", "stuckmap"):
        want = r[k + "_cells"][r[k + "_off"][t]:r[k + "_off"][t + 1]]
        assert np.array_equal(got[k], want), (t, k, len(got[k]), len(want))
    assert np.array_equal(got["groups"], r["group_state"][t]), (t, "groups")


def load_astar(path):
    """tests/golden/astar_*.npz -> maps ([H][W] uint8 / float64), queries [n, 8], paths (list of int64 cell arrays)."""
    d = np.load(path, allow_pickle=True)
    lay = np.load(os.path.join(os.path.dirname(path), str(d["fixture"])), allow_pickle=True)
    H, W = lay["is_road_map"].shape
    step = np.array([W, 1, -W, -1], np.int64)
    first, off, codes = d["path_first"], d["path_off"], d["path_steps"]
    paths = []
    for i in range(len(first)):
        n = int(off[i + 1] - off[i])
        if n == 0:
            paths.append(np.zeros(0, np.int64))
            continue
        c = codes[off[i]:off[i + 1]]
        paths.append(first[i] + np.concatenate([[0], np.cumsum(step[c[:-1]])]))
    return dict(W=W, H=H, is_road_map=lay["is_road_map"].astype(np.uint8), road_type_map=lay["road_type_map"].astype(np.uint8),
                allowed_dirs_map=lay["allowed_dirs_map"].astype(np.uint8), occupancy=d["occupancy"], stop_map=d["stop_map"],
                density=d["density"], queries=d["queries"], paths=paths)
