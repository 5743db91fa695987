"""Intersection light groups from the layout's link tables (host side, O(#lights) table work).

Mirrors, for the default QUEUE_ACTUATED / FIXED_TIME controllers, what the reference builds in
``CityModel._create_intersection_light_groups`` (city_model.py:1587-1650) and
``IntersectionLightGroup.__init__`` (intersection_light_group.py:33-116):

* clusters = 4-connected components of ``_intersection_cells`` (the "ever intersection" aux bit),
  labelled on the device (``tsim_label_mask``); a cluster whose four diagonal corner cells (one cell
  outside its bounding box) hold no TrafficLight forms no group (:1623-1635);
* canonical group order = cluster order = raster order of the cluster's first cell (the reference
  iterates a Python set there; the tick tapes fix the activation order to this canonical one);
* ``opposite_pairs`` (:243-279): a light joins the N-S (W-E) list if one of its controlled road cells
  has, as its FIRST arrow that enters an Intersection cell of this very cluster, a N/S (E/W) arrow;
* lane cells (:141-154): every assigned incoming -- and, with ``forward_traffic_light_range``, outgoing -- lane cell (with
  multiplicity) that has a N or S arrow goes to ``ns_in`` if it lies below the light, else to ``ns_out``; a cell without one
  but with an E or W arrow goes to ``ew_in`` if it lies left of the light, else to ``ew_out`` (the *_out lists are only read
  by PRESSURE_CONTROL, :448-461).

Everything is returned as CSR int32 arrays in the layout of ``tsim_light_tables`` (include/tsim.h).
"""
from __future__ import annotations

import numpy as np

T_INTER, T_TL = 12, 15
DX = np.array([0, 1, 0, -1], np.int64)   # N, E, S, W
DY = np.array([1, 0, -1, 0], np.int64)


def _csr_from_pairs(owner, value, n_owner):
    """rows sorted by owner (stable) -> (off[n_owner+1], values)"""
    order = np.argsort(owner, kind="stable")
    off = np.zeros(n_owner + 1, np.int32)
    np.add.at(off, owner + 1, 1)
    return np.cumsum(off).astype(np.int32), value[order].astype(np.int32)


def build_light_tables(W, H, cell_type, dirs, cluster_label, cluster_table, lights, ctrl_off, ctrl_cell, inc_off, inc_cell,
                       out_off=None, out_cell=None):
    """All inputs are host numpy arrays; planes are [H, W].  out_off / out_cell: the lights' assigned OUTGOING lane cells
    (only a layout with ``forward_traffic_light_range`` has them)."""
    T = cell_type.reshape(-1)
    D = dirs.reshape(-1).astype(np.int64)
    L = cluster_label.reshape(-1)
    lights = np.asarray(lights, np.int64)
    n_lights = len(lights)
    ctrl_off = np.asarray(ctrl_off, np.int64); inc_off = np.asarray(inc_off, np.int64)
    nC = len(cluster_table)
    # light -> own cell followed by its controlled cells
    tl_off = (ctrl_off[: n_lights + 1] + np.arange(n_lights + 1)).astype(np.int32)
    tl_cells = np.zeros(int(tl_off[-1]) if n_lights else 0, np.int32)
    if n_lights:
        tl_cells[tl_off[:-1]] = lights
        body = np.ones(len(tl_cells), bool); body[tl_off[:-1]] = False
        tl_cells[body] = ctrl_cell[: int(ctrl_off[n_lights])]
    # corner lights per cluster, corner order of the reference (:1623-1624)
    minx, miny, maxx, maxy = (cluster_table[:, k].astype(np.int64) for k in range(4))
    cl_ids, cl_lights = [], []
    for cx, cy in ((minx - 1, miny - 1), (maxx + 1, miny - 1), (minx - 1, maxy + 1), (maxx + 1, maxy + 1)):
        ok = (cx >= 0) & (cx < W) & (cy >= 0) & (cy < H)
        cell = np.where(ok, cy * W + cx, 0)
        ok &= T[cell] == T_TL
        idx = np.searchsorted(lights, cell[ok])
        cl_ids.append(np.flatnonzero(ok)); cl_lights.append(idx)
    slot = np.concatenate([np.full(len(c), k) for k, c in enumerate(cl_ids)]) if nC else np.zeros(0, np.int64)
    cl = np.concatenate(cl_ids) if nC else np.zeros(0, np.int64)
    li = np.concatenate(cl_lights) if nC else np.zeros(0, np.int64)
    order = np.lexsort((slot, cl))
    cl, li = cl[order], li[order]
    group_clusters = np.unique(cl)                      # cluster index (0-based) of every group, canonical order
    ng = len(group_clusters)
    gi = np.searchsorted(group_clusters, cl)            # group index of every (group, light) pair
    g_all_off, g_all = _csr_from_pairs(gi, li, ng)
    # opposite pairs: per (group, light, controlled cell) the first arrow entering this cluster's Intersection
    n_ctrl = (ctrl_off[li + 1] - ctrl_off[li]).astype(np.int64)
    pg = np.repeat(gi, n_ctrl); pl = np.repeat(li, n_ctrl)
    start = np.repeat(ctrl_off[li], n_ctrl)
    within = np.arange(len(pg)) - np.repeat(np.cumsum(n_ctrl) - n_ctrl, n_ctrl)
    cb = ctrl_cell[start + within].astype(np.int64)
    want_label = (group_clusters[pg] + 1)
    axis = np.full(len(pg), -1, np.int64)
    cbx, cby, cbd = cb % W, cb // W, D[cb]
    for i in range(4):
        has = ((cbd >> 12) & 7) > i
        d = (cbd >> (4 + 2 * i)) & 3
        nx, ny = cbx + DX[d], cby + DY[d]
        inb = (nx >= 0) & (nx < W) & (ny >= 0) & (ny < H)
        nc = np.where(inb, ny * W + nx, 0)
        hit = has & inb & (T[nc] == T_INTER) & (L[nc] == want_label) & (axis < 0)
        axis[hit] = (d[hit] & 1)                        # N,S -> 0 (vertical), E,W -> 1 (horizontal)
    tabs = {}
    for key, ax in (("g_ns", 0), ("g_ew", 1)):
        sel = axis == ax
        pairs = np.unique(np.stack([pg[sel], pl[sel]], 1), axis=0) if sel.any() else np.zeros((0, 2), np.int64)
        tabs[key + "_off"], tabs[key] = _csr_from_pairs(pairs[:, 0], pairs[:, 1], ng)
    # lane cells: per light its incoming list, then its outgoing list (:141-142)
    def lane_pairs(off, cell):
        off = np.asarray(off, np.int64)
        cnt = (off[li + 1] - off[li]).astype(np.int64)
        start = np.repeat(off[li], cnt)
        within = np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        pair = np.repeat(np.arange(len(li)), cnt)          # index of the (group, light) pair: keeps the reference's order
        return pair, np.asarray(cell)[start + within].astype(np.int64), np.zeros(len(pair), np.int64)
    pi, rb, half = lane_pairs(inc_off, inc_cell)
    if out_off is not None and len(np.asarray(out_cell)):
        po, rbo, _ = lane_pairs(out_off, out_cell)
        pi, rb, half = np.concatenate([pi, po]), np.concatenate([rb, rbo]), np.concatenate([half, np.ones(len(po), np.int64)])
        order = np.lexsort((half, pi))                     # stable inside a half: a light's incoming cells, then its outgoing ones
        pi, rb = pi[order], rb[order]
    qg, ql = gi[pi], li[pi]
    rbd = D[rb] & 0xF
    tlx, tly = lights[ql] % W, lights[ql] // W
    vertical = (rbd & 0b0101) != 0
    horizontal = ~vertical & ((rbd & 0b1010) != 0)
    below, left = rb // W < tly, rb % W < tlx
    for key, sel in (("g_nsin", vertical & below), ("g_nsout", vertical & ~below), ("g_ewin", horizontal & left), ("g_ewout", horizontal & ~left)):
        tabs[key + "_off"], tabs[key] = _csr_from_pairs(qg[sel], rb[sel], ng)
    # cluster cells of the groups
    cells = np.flatnonzero(L > 0)
    lab = L[cells].astype(np.int64) - 1
    keep = np.isin(lab, group_clusters)
    cells, lab = cells[keep], lab[keep]
    tabs["g_cl_off"], tabs["g_cl"] = _csr_from_pairs(np.searchsorted(group_clusters, lab), cells, ng)
    tabs.update(tl_off=tl_off, tl_cells=tl_cells, g_all_off=g_all_off, g_all=g_all, n_groups=ng, n_lights=n_lights,
                group_clusters=group_clusters.astype(np.int32))
    return tabs


def pressure_cells(tabs, W):
    """The lane tables as PRESSURE_CONTROL reads them.  ``run_pressure_control`` (intersection_light_group.py:448-461) passes
    ``to_int32(occ_map)`` -- the [H, W] occupancy map reshaped to (-1, 2), :443-446 -- to ``compute_max_pressure``
    (utilities/numba_utilities.py:74-85), whose ``occupancy_map[y, x]`` then reads flat element ``2 * y + x`` of the map (Numba
    does not check bounds; the element always exists).  Bit-exact parity means counting the vehicles on THOSE cells."""
    out = dict(tabs)
    for k in ("g_nsin", "g_ewin", "g_nsout", "g_ewout"):
        c = np.asarray(tabs[k], np.int64)
        out[k] = (2 * (c // W) + c % W).astype(np.int32)
    return out


def groups_as_cell_lists(tabs, lights):
    """Group tables as sorted cell-index arrays (the form the reference fixtures use), for parity checks."""
    lights = np.asarray(lights)
    out = []
    for g in range(tabs["n_groups"]):
        sl = lambda k: tabs[k][tabs[k + "_off"][g]:tabs[k + "_off"][g + 1]]
        out.append(dict(cluster=np.sort(sl("g_cl")), lights=np.sort(lights[sl("g_all")]), ns_lights=np.sort(lights[sl("g_ns")]),
                        ew_lights=np.sort(lights[sl("g_ew")]), ns_in=np.sort(sl("g_nsin")), ew_in=np.sort(sl("g_ewin")),
                        ns_out=np.sort(sl("g_nsout")), ew_out=np.sort(sl("g_ewout"))))
    return out


# ---------------------------------------------------------------------------------------------------------------
# Neighbour links: IntersectionLightGroup.populate_links (intersection_light_group.py:175-242)
# ---------------------------------------------------------------------------------------------------------------
LINK_DIRS = ("N", "S", "E", "W")                      # Defaults.AVAILABLE_DIRECTIONS (config.py:62)
_STEP = {"N": (0, 1), "S": (0, -1), "E": (1, 0), "W": (-1, 0)}


def neighbor_links(W, H, cell_type, cluster_label, tabs, lights, hbands, vbands, creation_order=None, max_depth=1000):
    """``neighbor_groups`` of every light group as an int32 table [n_groups, 4] (columns N, S, E, W; -1 = none).

    Restates the ray marches of ``populate_links`` INCLUDING their order dependence: whether a group ``g`` "blocks all lanes" for
    marches in direction ``d`` is computed once, by the first march that reaches it, at that march's hit cell, and cached on ``g``
    (``_blocks_<d>``, :226-228).  The reference runs the marches twice: once inside every group's constructor (only groups created
    earlier are visible then, :141 / city_model.py:1638-1650) and once more when ``get_opposite_traffic_lights`` finds the axis lists
    still empty (:303-307) -- under the harness that second pass happens for all groups, in creation order, before the first tick.
    Both passes are replayed here in ``creation_order`` (group indices; default: the canonical order of ``tabs``).  The reference
    creates its groups in the iteration order of a Python set (city_model.py:1595), so exact parity for a city whose cached answers
    depend on that order needs the reference's order as an input; `order_dependent` in the result says whether any answer did.

    Host-side and sequential (a Python loop per march): meant for reference-sized cities.
    """
    T = np.asarray(cell_type).reshape(H, W)
    L = np.asarray(cluster_label).reshape(H, W)
    ng = int(tabs["n_groups"])
    lights = np.asarray(lights, np.int64)
    group_of_cluster = np.full(int(L.max()) + 1, -1, np.int64)
    group_of_cluster[np.asarray(tabs["group_clusters"], np.int64) + 1] = np.arange(ng)
    hb = [tuple(int(v) for v in b[:2]) for b in np.asarray(hbands).reshape(-1, 4)]
    vb = [tuple(int(v) for v in b[:2]) for b in np.asarray(vbands).reshape(-1, 4)]

    def band_or_single(idx, bands):                    # _find_band_covering (city_model.py:1269-1273): the first band that covers idx
        for st, en in bands:
            if st <= idx <= en:
                return st, en
        return idx, idx

    inter = lambda x, y: T[y, x] == T_INTER

    def blocks_all_lanes(ix, iy, d):                   # :185-205
        if d in ("N", "S"):
            vx0, vx1 = band_or_single(ix, vb)
            if vx1 == vx0:
                hy0, hy1 = band_or_single(iy, hb)
                return bool(inter(vx0, iy) and (hy1 != hy0 or inter(ix, hy0)))
            return bool(all(inter(xx, iy) for xx in range(vx0, vx1 + 1)))
        hy0, hy1 = band_or_single(iy, hb)
        if hy1 == hy0:
            vx0, vx1 = band_or_single(ix, vb)
            return bool(inter(ix, hy0) and (vx1 != vx0 or inter(vx0, iy)))
        return bool(all(inter(ix, yy) for yy in range(hy0, hy1 + 1)))

    diag = []
    for g in range(ng):                                # :213-221, the group's lights in corner order
        cells = []
        for l in tabs["g_all"][tabs["g_all_off"][g]:tabs["g_all_off"][g + 1]]:
            lx, ly = int(lights[l] % W), int(lights[l] // W)
            for dx, dy in ((1, 1), (1, -1), (-1, 1), (-1, -1)):
                nx, ny = lx + dx, ly + dy
                if 0 <= nx < W and 0 <= ny < H and inter(nx, ny):
                    cells.append((nx, ny))
        diag.append(cells)

    blocks, answers = {}, {}                           # (group, direction) -> cached answer / every answer any march would get

    def march(me, visible):
        nbr = {}
        for cx, cy in diag[me]:
            for d in LINK_DIRS:
                sx, sy = _STEP[d]
                x, y, steps = cx, cy, 0
                while steps < max_depth:
                    x, y = x + sx, y + sy
                    if not (0 <= x < W and 0 <= y < H):
                        break
                    g = int(group_of_cluster[L[y, x]]) if inter(x, y) else -1
                    if g < 0 or g == me or not visible[g]:
                        steps += 1
                        continue
                    ans = blocks_all_lanes(x, y, d)
                    answers.setdefault((g, d), set()).add(ans)
                    if (g, d) not in blocks:
                        blocks[(g, d)] = ans
                    if blocks[(g, d)]:
                        nbr[d] = g
                        break
                    steps += 1
        return nbr

    order = list(range(ng)) if creation_order is None else [int(g) for g in creation_order]
    visible = np.zeros(ng, bool)
    for g in order:                                    # pass 1: inside the constructors
        march(g, visible)
        visible[g] = True
    table = np.full((ng, 4), -1, np.int32)
    for g in order:                                    # pass 2: everybody visible, cached answers kept
        for d, n in march(g, visible).items():
            table[g, LINK_DIRS.index(d)] = n
    return dict(nbr=table, order_dependent=any(len(v) > 1 for v in answers.values()))
