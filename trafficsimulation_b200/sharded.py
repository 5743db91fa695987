"""Row-band shards of the layout pipeline across the GPUs of one box.

The city's ``[H][W]`` planes are cut into ``n_shards`` contiguous row bands.  Shard ``s`` OWNS rows
``[own_lo, own_hi)`` and holds a WINDOW ``[win_lo, win_hi)`` = its own rows plus ``halo`` rows of each
neighbour.  Every pass runs on the whole window with the single-GPU kernels (a window is a small city whose
out-of-window rows do not exist), then the shards refresh their halo rows from the owners:

* closed-form pass (frame + roads): no communication at all, band tables are replicated;
* stencil and per-block passes: one neighbour exchange of ``halo`` rows per plane after the pass.  A pass is
  exact on a shard's own rows when everything those rows depend on lies inside the window: 1 row for the
  stencils, ``traffic_light_range + 2`` for the lights, one block height for carving / zoning / entrances.
  ``halo`` (default 64) must exceed the tallest block (``max_block_spacing`` = 18 by default); a component
  that touches both a shard's own rows and a non-grid window edge raises ``TSIM_ERR_CAPACITY`` -- never a
  silent wrong answer;
* labelling: ids are global raster discovery ranks.  A component that meets a shard's own rows lies inside
  its window, so its window root is its true root; every shard counts the roots in its own rows, one
  all-gather + prefix gives ``id_base`` (device scalar consumed by the kernels through ``tsim_blobs``);
* fixed points (dead ends, reachability for ``leads_to``): local fixed point, halo exchange, all-reduce of
  the "changed" flags, repeat until no shard changed.  Dead-end removal treats rows beyond a shard cut as
  unknown (= road), so missing information can only delay a removal (the pass is monotone).

LEAN mode (``lean=True``; what ``bench.py --gpus N`` runs): no halo refresh at all.  A shard computes its halo rows itself
on every pass; what it cannot see beyond its window edge spoils at most the dependency radius of the pass, pass after pass,
from the edge inwards, so with a halo deeper than the sum of the radii (block height for carving / zoning / entrances,
``traffic_light_range`` + the reach of the local `leads_to` searches for the lights, 1 for the stencils: about 150 rows
with the defaults, hence ``halo=192``) the own rows come out exact without a single exchange.  "Deeper than the sum" is
not taken on trust (dead-end chains have no a-priori length): after EVERY pass both shards of a cut digest the
``verify`` rows on either side of it (``tsim_rows_digest``, position-weighted, a few hundred KB read per band), and the
digests are compared once at the end -- one all-gather of a few hundred bytes.  Spoilt data can only reach an own row by
first spoiling the verify band of the previous pass, where the owner's copy is still exact, so equal digests on both sides
of every cut prove the own rows exact; a mismatch raises ``TSIM_ERR_CAPACITY`` (flag 26: halo too small).  What is left of
the communication is the all-gather of the root counts per labelling (global block ids) and that comparison.

Two deployments share this code: one process per GPU under ``torch.distributed`` (NCCL; gloo on CPU for
the host-logic tests) with one local shard each, or ONE process holding all shards on one device (how the
``-m gpu`` tests check N-shard == 1-shard on a single GPU).  The data path has no other collective.
"""
from __future__ import annotations

import numpy as np
import torch

INF32 = 0x7FFFFFFF
ERR_HALO = 26   # a component meets both a shard's own rows and a cut edge of its window: halo too small


class ShardPlan:
    """Row ranges of every shard: own rows (even split) and window (own rows + halo, clipped to the grid)."""

    def __init__(self, height: int, n_shards: int, halo: int, cuts=None):
        """``cuts``: the n_shards - 1 rows where the bands meet (default: an even split)."""
        if n_shards < 1 or height < n_shards:
            raise ValueError(f"cannot cut {height} rows into {n_shards} shards")
        self.height, self.n, self.halo = int(height), int(n_shards), int(halo)
        if cuts is None:
            cuts = [s * height // n_shards for s in range(1, n_shards)]
        cuts = [int(c) for c in cuts]
        if len(cuts) != n_shards - 1 or any(b <= a for a, b in zip([0] + cuts, cuts + [height])):
            raise ValueError(f"cuts {cuts} do not split {height} rows into {n_shards} bands")
        self.own_lo = [0] + cuts
        self.own_hi = cuts + [int(height)]
        if n_shards > 1 and min(h - l for l, h in zip(self.own_lo, self.own_hi)) < halo:
            raise ValueError(f"halo {halo} exceeds the rows of a shard ({height} rows / {n_shards} shards)")
        self.win_lo = [max(0, l - halo) for l in self.own_lo]
        self.win_hi = [min(height, h + halo) for h in self.own_hi]

    def up_rows(self, s):
        """(global row range) shard s sends UP to s+1 = the lower halo of s+1 (rows s owns)."""
        return self.win_lo[s + 1], self.own_lo[s + 1]

    def down_rows(self, s):
        """(global row range) shard s sends DOWN to s-1 = the upper halo of s-1 (rows s owns)."""
        return self.own_hi[s - 1], self.win_hi[s - 1]


class Comm:
    """Neighbour exchange and the two tiny collectives, for shards that are all local (``group is None`` and
    world size 1) or spread one per rank over a ``torch.distributed`` group."""

    def __init__(self, n_shards: int, dist_enabled: bool, group=None):
        self.n = n_shards
        self.dist = None
        if dist_enabled:
            import torch.distributed as dist
            self.dist = dist
            self.group = group
            self.rank = dist.get_rank(group)
            if dist.get_world_size(group) != n_shards:
                raise ValueError("one shard per rank: world size must equal n_shards")
            self.local = [self.rank]
        else:
            self.local = list(range(n_shards))

    # halo exchange.  items: list of (get, put) with get(s, lo, hi) -> view of shard s's rows [lo, hi) (global row
    # numbers) and put(s, lo, hi, src).  All items travel in ONE batch of sends / receives per neighbour pair.
    def exchange(self, plan: ShardPlan, get, put=None):
        """put(s, lo, hi, src) merges received rows; put = None means a plain refresh of the rows get(s, lo, hi) views."""
        items = [(get, put)] if callable(get) else list(get)
        if self.n == 1:
            return
        if self.dist is None:
            for g, p in items:
                staged = []
                for s in range(self.n - 1):
                    lo, hi = plan.up_rows(s)
                    staged.append((s + 1, lo, hi, g(s, lo, hi).clone()))
                    lo, hi = plan.down_rows(s + 1)
                    staged.append((s, lo, hi, g(s + 1, lo, hi).clone()))
                for dst, lo, hi, src in staged:
                    if p is None:
                        g(dst, lo, hi).copy_(src)
                    else:
                        p(dst, lo, hi, src)
            return
        dist, s = self.dist, self.rank
        ops, recvs, get_of = [], [], {}

        def add(g, p, peer, send_rows, recv_rows):   # rows travel as raw bytes (NCCL has no int16)
            src = g(s, *send_rows).contiguous()
            ops.append(dist.P2POp(dist.isend, src.view(torch.uint8).reshape(-1), peer, self.group))
            like = g(s, *recv_rows)
            if p is None and like.is_contiguous():           # plain refresh: receive straight into the halo rows
                ops.append(dist.P2POp(dist.irecv, like.view(torch.uint8).reshape(-1), peer, self.group))
                return
            buf = torch.empty(like.numel() * like.element_size(), dtype=torch.uint8, device=like.device)
            get_of[id(buf)] = g
            ops.append(dist.P2POp(dist.irecv, buf, peer, self.group))
            recvs.append((p, recv_rows, buf, like.dtype, like.shape))

        for g, p in items:
            if s + 1 < self.n:
                add(g, p, s + 1, plan.up_rows(s), plan.down_rows(s + 1))
            if s > 0:
                add(g, p, s - 1, plan.down_rows(s), plan.up_rows(s - 1))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for p, (lo, hi), buf, dtype, shape in recvs:
            if p is None:
                get_of[id(buf)](s, lo, hi).copy_(buf.view(dtype).view(shape))
            else:
                p(s, lo, hi, buf.view(dtype).view(shape))

    def all_gather(self, vals: dict):
        """vals: {local shard -> 1-D tensor[k]} -> {local shard -> tensor [n_shards, k]} (same device / dtype)."""
        if self.dist is None:
            return {s: torch.stack([vals[r].to(vals[s].device) for r in range(self.n)]) for s in self.local}
        v = vals[self.rank].contiguous().reshape(-1)
        out = torch.empty(self.n * v.numel(), dtype=v.dtype, device=v.device)
        self.dist.all_gather_into_tensor(out, v, group=self.group)
        return {self.rank: out.view(self.n, v.numel())}

    def max(self, vals: dict) -> int:
        """vals: {local shard -> 0-d / 1-element integer tensor}; the maximum over ALL shards (host sync)."""
        if self.dist is None:
            return max(int(v.reshape(-1)[0].item()) for v in vals.values())
        f = vals[self.rank].reshape(-1)[:1].to(torch.int32).clone()
        self.dist.all_reduce(f, op=self.dist.ReduceOp.MAX, group=self.group)
        return int(f.item())

    def any(self, flags: dict) -> bool:
        """flags: {local shard -> 0-d / 1-element integer tensor}; True if any shard's flag is non-zero (host sync)."""
        if self.dist is None:
            return any(bool(int(f.reshape(-1)[0].item())) for f in flags.values())
        f = flags[self.rank].reshape(-1)[:1].to(torch.int32).clone()
        self.dist.all_reduce(f, op=self.dist.ReduceOp.MAX, group=self.group)
        return bool(int(f.item()))


def id_bases(own_counts: torch.Tensor, n_lo: torch.Tensor, shard: int) -> torch.Tensor:
    """Global id base of a shard's window-local component numbering (0-d int32 tensor).

    own_counts[r] = roots in the own rows of shard r (all-gathered), n_lo = roots of THIS window below its
    own rows.  Window component k (0-based, raster order) has global 0-based id k + base with
    base = sum(own_counts[:shard]) - n_lo: the roots between the window's first owned root and any root that
    matters are all true roots (module docstring).
    """
    before = own_counts[:shard].sum() if shard > 0 else torch.zeros((), dtype=own_counts.dtype, device=own_counts.device)
    return (before - n_lo).to(torch.int32)


class ShardedCityLayout:
    """``GpuCityLayout`` over row-band shards.  Same constructor kwargs as the reference ``CityModel`` plus the
    shard geometry; ``generate`` runs the reference's pass sequence (city_model.py:125-139, 148)."""

    PASSES = ("frame", "carve", "zones", "dead_ends", "upgrade_r2", "entrances", "fix_dirs")

    def __init__(self, n_shards, halo=64, devices=None, distributed=False, group=None, global_reach=False, cuts=None, lean=False,
                 verify=None, **city_kwargs):
        from .layout import GpuCityLayout
        self.kw = dict(city_kwargs)
        self.width, self.height = int(city_kwargs.get("width", 200)), int(city_kwargs.get("height", 200))
        self.plan = ShardPlan(self.height, n_shards, halo if n_shards > 1 else 0, cuts)
        self.comm = Comm(n_shards, distributed, group)
        self.carve = bool(city_kwargs.get("carve_subblock_roads", False))
        self.global_reach = bool(global_reach)
        self.lean = bool(lean) and n_shards > 1
        self.verify = int(min(80, halo // 2) if verify is None else verify)   # rows on either side of a cut whose digests are compared
        if self.lean and (self.global_reach or not 0 < self.verify <= halo):
            raise ValueError("lean shards: verify rows must fit the halo, and the reachability exchange is not available")
        self.shards = {}
        for i, s in enumerate(self.comm.local):
            dev = (devices[i] if devices else (city_kwargs.get("device", "cuda:0")))
            kw = {k: v for k, v in city_kwargs.items() if k != "device"}
            self.shards[s] = GpuCityLayout(device=dev, win_y0=self.plan.win_lo[s], win_rows=self.plan.win_hi[s] - self.plan.win_lo[s],
                                           win_halo=self.plan.halo, **kw)
        self.n_blocks = None
        self.trace = None   # set to [] to collect (label, cuda event) pairs of one generate() call (see phase_times)
        # lean mode: digests[pass][band] per local shard; bands: 0 own rows above the lower cut, 1 halo rows below it,
        # 2 own rows below the upper cut, 3 halo rows above it (int64 bit patterns of the unsigned digests)
        self._dig = {s: torch.zeros(len(self.PASSES), 4, dtype=torch.int64, device=L.device) for s, L in self.shards.items()} if self.lean else None

    # ------------------------------------------------------------------ plumbing
    def set_bands(self, hbands, vbands):
        for L in self.shards.values():
            L.set_bands(hbands, vbands)
        self.global_cap = 3 * (len(hbands) + 2) * (len(vbands) + 2) + 64   # bound on the blocks of the whole city

    def _plane(self, s, name):
        L = self.shards[s]
        t = getattr(L, name)
        return t.view(L.win_rows, self.width)

    def _exchange(self, *names):
        wl = self.plan.win_lo
        self.comm.exchange(self.plan, [(lambda s, lo, hi, name=name: self._plane(s, name)[lo - wl[s]: hi - wl[s]], None) for name in names])

    def _digest(self, pass_name, what):
        """Lean mode: add the digests of the verify bands around this shard's cuts after pass `pass_name` (`what`: plane mask)."""
        if not self.lean:
            return
        import ctypes as C
        from . import _lib
        p, v, k = self.plan, self.verify, self.PASSES.index(pass_name)
        for s, L in self.shards.items():
            w0 = p.win_lo[s]
            bands = []
            if s > 0:
                bands += [(0, p.own_lo[s], p.own_lo[s] + v), (1, p.own_lo[s] - v, p.own_lo[s])]
            if s + 1 < p.n:
                bands += [(2, p.own_hi[s] - v, p.own_hi[s]), (3, p.own_hi[s], p.own_hi[s] + v)]
            for b, lo, hi in bands:
                out = C.c_void_p(self._dig[s].data_ptr() + 8 * (4 * k + b))
                _lib.check(L.lib.tsim_rows_digest(C.byref(L.cfg), C.byref(L._planes), lo - w0, hi - w0, what, out, L._stream))

    def _verify_digests(self):
        """Lean mode, end of the pipeline: every cut's two shards must have digested the same bytes after every pass."""
        self._mark("verify: last digests")
        gathered = self.comm.all_gather({s: d.reshape(-1) for s, d in self._dig.items()})
        self._mark("verify: all-gather")
        for s, L in self.shards.items():
            g = gathered[s].view(self.plan.n, len(self.PASSES), 4)
            bad = torch.zeros((), dtype=torch.bool, device=L.device)
            if s > 0:
                bad = bad | (g[s, :, 1] != g[s - 1, :, 2]).any() | (g[s, :, 0] != g[s - 1, :, 3]).any()
            if s + 1 < self.plan.n:
                bad = bad | (g[s, :, 2] != g[s + 1, :, 1]).any() | (g[s, :, 3] != g[s + 1, :, 0]).any()
            L.flags[0] = torch.where(bad & (L.flags[0] == 0), torch.full_like(L.flags[0], ERR_HALO), L.flags[0])

    def _label_and_number(self):
        """Label every window, then turn the window-local numbering into global raster ranks (id_base)."""
        p, W = self.plan, self.width
        if p.n == 1:                                          # one window = the whole grid: window ranks ARE global ranks
            L = self.shards[0]
            L._label_async()
            self._total = L.flags[2]
            return
        import ctypes as C
        from . import _lib
        counts, n_lo = {}, {}
        for s, L in self.shards.items():
            L._label_async()
            out = L.flags[12:15]                              # n_lo, roots in the own rows, halo-too-small marker
            _lib.check(L.lib.tsim_shard_counts(C.byref(L.cfg), C.byref(L._blobs), p.own_lo[s], p.own_hi[s], C.c_void_p(out.data_ptr()), L._stream))
            n_lo[s] = out[0]
            counts[s] = out[1:2].to(torch.int64)
            # a component that meets the own rows must not reach a cut edge of the window
            L.flags[0] = torch.where((out[2] != 0) & (L.flags[0] == 0), torch.full_like(L.flags[0], ERR_HALO), L.flags[0])
        gathered = self.comm.all_gather(counts)
        for s, L in self.shards.items():
            own = gathered[s].reshape(-1)
            L.flags[4] = id_bases(own, n_lo[s], s)
            self._total = own.sum()

    def synth_carve_tapes(self, seed):
        """Synthetic carve tapes for cities the reference cannot reach (tapes.synth_carve_tape_device): runs the frame
        pass and the labelling once, returns {local shard -> int32 tape [global_cap + 1, 8]} for ``generate``."""
        from . import tapes
        for L in self.shards.values():
            L._build_roads_and_sidewalks()
        self._label_and_number()
        uni = tapes.synth_carve_uniforms(seed, self.global_cap)
        out = {}
        for s, L in self.shards.items():
            u = torch.from_numpy(uni).to(L.device)
            t = torch.zeros((self.global_cap + 1, 8), dtype=torch.int32, device=L.device)
            out[s] = tapes.synth_carve_tape_device(u, L.blobs.view(-1, 6), L.flags[2], L.flags[4], t)
        return out

    def _mark(self, label):
        if self.trace is not None:
            L = next(iter(self.shards.values()))
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(L.device))
            self.trace.append((label, ev))

    def phase_times(self):
        """ms between the marks of the traced generate() call (device time line of the first local shard)."""
        torch.cuda.synchronize()
        return [(b[0], a[1].elapsed_time(b[1])) for a, b in zip(self.trace[:-1], self.trace[1:])]

    def _check(self, what):
        for L in self.shards.values():
            L._check_flag(what)

    # ------------------------------------------------------------------ the pipeline
    def generate(self, tape_zone, tape_carve=None, tape_entrance=None, check=True, lights=True, maps=True):
        S = self.shards
        if self.plan.n == 1:                                  # no cuts: the plain single-GPU sequence, no host round trips
            L = S[0]
            L._place_thick_wall(); L._place_sidewalk_inner_ring(); L._clear_interior()
            L._build_roads_and_sidewalks()
            if self.carve:
                L._carve_subblock_roads(tape_carve[0] if isinstance(tape_carve, dict) else tape_carve, check=False)
            L._flood_fill_blocks_storing_data(tape_zone, check=False)
            self._total = L.flags[2]
            L._eliminate_dead_ends()
            L._upgrade_r2_to_intersections(check=False)
            L._final_place_block_entrances(tape_entrance, check=False)
            L._remove_invalid_intersection_directions(); L._add_entrance_directions()
            if lights:
                L._add_traffic_lights(check=False)
            if maps:
                L._build_simple_maps()
            self.dead_end_rounds = self.reach_rounds = 1
            if check:
                self._check("generate")
                self.n_blocks = int(self._total.item())
            return
        if tape_entrance is None:                             # per-block tapes are indexed by GLOBAL block id
            tape_entrance = np.zeros(self.global_cap, np.int32)
        if self.lean:
            return self._generate_lean(tape_zone, tape_carve, tape_entrance, check, lights, maps)
        self._mark("start")
        for L in S.values():
            L._build_roads_and_sidewalks()                    # closed form: exact on the whole window, no exchange
        self._mark("frame_roads")
        if self.carve:
            self._label_and_number()
            self._mark("carve: label + number")
            for s, L in S.items():
                L._carve_subblock_roads(tape_carve[s] if isinstance(tape_carve, dict) else tape_carve, check=False, relabel=False)
            self._mark("carve")
            self._exchange("cell_type", "dirs", "aux")
            self._mark("carve: exchange")
        self._label_and_number()
        self._mark("zones: label + number")
        for L in S.values():
            L._flood_fill_blocks_storing_data(tape_zone, check=False, relabel=False)
        self._mark("zones")
        self._exchange("cell_type", "block_id")
        self._mark("zones: exchange")
        self.dead_end_rounds = 0
        while True:                                           # monotone pruning: local fixed point, exchange, repeat
            for L in S.values():
                L._eliminate_dead_ends()
            self._exchange("cell_type", "dirs", "aux")
            self.dead_end_rounds += 1
            if not self.comm.any({s: (L.flags[1] > 1).to(torch.int32) for s, L in S.items()}):
                break
        self._mark("dead_ends (+exchange, host round trip)")
        for L in S.values():
            L._upgrade_r2_to_intersections(check=False)
        self._exchange("cell_type", "dirs", "aux")
        self._mark("upgrade_r2 + exchange")
        for L in S.values():
            L._final_place_block_entrances(tape_entrance, check=False)
        self._exchange("cell_type", "dirs", "aux", "block_id")
        self._mark("entrances + exchange")
        for L in S.values():
            L._remove_invalid_intersection_directions()
            L._add_entrance_directions()
        self._exchange("dirs")
        self._mark("fix_dirs + exchange")
        if lights:
            self._lights()
            self._mark("lights + exchange")
        if maps:
            for L in S.values():
                L._build_simple_maps()
            self._mark("maps")
        if check:
            self._check("generate")
            self.n_blocks = int(self._total.item())

    def _generate_lean(self, tape_zone, tape_carve, tape_entrance, check, lights, maps):
        """The pipeline without halo exchanges (module docstring): every pass on the whole window, a digest of the verify bands
        after it, one comparison at the end.  Communication: the root counts of the labellings and the digests."""
        S = self.shards
        T, D, A, B = 1, 2, 4, 8
        for d in self._dig.values():
            d.zero_()
        self._mark("start")
        for L in S.values():
            L._build_roads_and_sidewalks()
        self._digest("frame", T | D | A)
        self._mark("frame_roads")
        if self.carve:
            self._label_and_number()
            self._mark("carve: label + number")
            for s, L in S.items():
                L._carve_subblock_roads(tape_carve[s] if isinstance(tape_carve, dict) else tape_carve, check=False, relabel=False)
            self._digest("carve", T | D | A)
            self._mark("carve")
        self._label_and_number()
        self._mark("zones: label + number")
        for L in S.values():
            L._flood_fill_blocks_storing_data(tape_zone, check=False, relabel=False)
        self._digest("zones", T | B)
        self._mark("zones")
        for L in S.values():
            L._eliminate_dead_ends()                          # local fixed point; a chain that comes in over the window edge shows in the digests
        self.dead_end_rounds = 1
        self._digest("dead_ends", T | D | A)
        self._mark("dead_ends")
        for L in S.values():
            L._upgrade_r2_to_intersections(check=False)
        self._digest("upgrade_r2", T | D | A)
        self._mark("upgrade_r2")
        for L in S.values():
            L._final_place_block_entrances(tape_entrance, check=False)
        self._digest("entrances", T | D | A | B)
        self._mark("entrances")
        for L in S.values():
            L._remove_invalid_intersection_directions()
            L._add_entrance_directions()
        self._digest("fix_dirs", D)
        self._mark("fix_dirs")
        if lights:
            for L in S.values():
                L._add_traffic_lights(check=False)
            self.reach_rounds = 1
            self._mark("lights")                              # the last pass that changes the planes: nothing downstream reads its halo rows
        if maps:
            for L in S.values():
                L._build_simple_maps()
            self._mark("maps")
        self._verify_digests()
        self._mark("verify")
        if check:
            self._check("generate")
            self.n_blocks = int(self._total.item())

    def _lights(self):
        """`leads_to` (cell.py:201-227) is reachability over the whole arrow graph, but every query of the lights pass
        relates two cells at most ``traffic_light_range + 1`` apart on one lane.  Default: every shard answers from
        reachability planes closed INSIDE its window around its own pivot (a -> pivot -> b inside the window is a
        real path); a query the planes cannot answer is searched exactly, and a search that fails after touching a
        shard cut raises (flag 13) instead of answering "unreachable" -- so no reachability data crosses shards and
        the pass costs what it costs on one GPU.  ``global_reach=True`` runs the exchange protocol instead: one pivot
        for the whole city, closure + OR-exchange of the halo rows of the planes until no shard changes."""
        p, W, S = self.plan, self.width, self.shards
        if not self.global_reach:
            for L in S.values():
                L._add_traffic_lights(check=False)
            self.reach_rounds = 1
            self._exchange("cell_type", "aux", "block_id")
            return
        cands = {}
        for s, L in S.items():
            L._lights_prepare()
            c = L.flags[8:10].to(torch.int64)
            cands[s] = torch.where(c == INF32, torch.full_like(c, 2 ** 62), c + p.win_lo[s] * W)   # global raster index
        gathered = self.comm.all_gather(cands)
        for s, L in S.items():
            g = gathered[s]
            mid, first = g[:, 0].min(), g[:, 1].min()
            piv = torch.where(mid < 2 ** 62, mid, first)
            local = piv - p.win_lo[s] * W
            inside = (piv < 2 ** 62) & (local >= 0) & (local < L.win_rows * W)
            L.flags[10] = torch.where(inside, local, torch.full_like(local, -1)).to(torch.int32)
            L._lights_seed()
        self.reach_rounds = 0
        planes = {s: L.reach_planes() for s, L in S.items()}
        while True:
            changed = {}
            for s, L in S.items():
                L.flags[11] = 0
                L._lights_reach(-1 if self.reach_rounds == 0 else p.halo)   # later rounds: only the halo rows received new bits
                changed[s] = L.flags[11].clone()
            items = []
            for k in (0, 1):                                  # OR the owners' rows into the neighbours' halo rows
                def get(s, lo, hi, k=k):
                    return planes[s][k][lo - p.win_lo[s]: hi - p.win_lo[s]]

                def put(s, lo, hi, src, k=k):
                    dst = planes[s][k][lo - p.win_lo[s]: hi - p.win_lo[s]]
                    merged = dst | src
                    changed[s] = changed[s] | (merged != dst).any().to(torch.int32)
                    dst.copy_(merged)
                items.append((get, put))
            self.comm.exchange(p, items)
            self.reach_rounds += 1
            if not self.comm.any(changed):
                break
        for L in S.values():
            L._lights_finish(check=False)
        self._exchange("cell_type", "aux", "block_id")

    # ------------------------------------------------------------------ read-back (own rows)
    def planes_host(self):
        """Own rows of every LOCAL shard, stacked in row order (the whole city when all shards are local)."""
        out = {}
        for s in sorted(self.shards):
            L, p = self.shards[s], self.plan
            h = L.planes_host()
            a, b = p.own_lo[s] - p.win_lo[s], p.own_hi[s] - p.win_lo[s]
            for k, v in h.items():
                out.setdefault(k, []).append(v[a:b])
        return {k: np.concatenate(v, 0) for k, v in out.items()}

    def maps_host(self):
        out = {}
        for s in sorted(self.shards):
            L, p = self.shards[s], self.plan
            a, b = p.own_lo[s] - p.win_lo[s], p.own_hi[s] - p.win_lo[s]
            for k, v in L.maps_host().items():
                out.setdefault(k, []).append(v[a:b])
        return {k: np.concatenate(v, 0) for k, v in out.items()}

    def light_links_host(self):
        """Link tables of the lights each local shard OWNS, as sorted (light, cell) pairs in GLOBAL cell indices."""
        p, W = self.plan, self.width
        lights, ctrl, inc = [], [], []
        for s in sorted(self.shards):
            L = self.shards[s]
            g = L.light_links_host()
            off = p.win_lo[s] * W
            lo, hi = (p.own_lo[s] - p.win_lo[s]) * W, (p.own_hi[s] - p.win_lo[s]) * W
            keep = (g["lights"] >= lo) & (g["lights"] < hi)
            lights.append(g["lights"][keep].astype(np.int64) + off)
            for name, acc in (("ctrl", ctrl), ("incoming", inc)):
                pr = g[name].astype(np.int64)
                k = (pr[:, 0] >= lo) & (pr[:, 0] < hi)
                acc.append(pr[k] + off)
        return {"lights": np.concatenate(lights), "ctrl": np.concatenate(ctrl, 0), "incoming": np.concatenate(inc, 0)}
