"""Host-side logic of the row-band shards under torch.distributed (gloo, world size 2 and 3, CPU tensors):
shard plan, neighbour halo exchange, the id-base prefix over all-gathered root counts and the fixed-point
flag reduction (the kernels themselves are covered by tests/test_gpu_sharded.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from trafficsimulation_b200.sharded import Comm, ShardPlan, id_bases


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, H, W, halo, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ShardPlan(H, world, halo)
        comm = Comm(world, True)
        assert comm.local == [rank]
        lo, hi = plan.win_lo[rank], plan.win_hi[rank]
        # the "truth": cell (y, x) holds y * W + x; every rank fills only its OWN rows, halos start as garbage
        for dtype in (torch.uint8, torch.int16, torch.int32, torch.int64):
            truth = (torch.arange(H * W, dtype=torch.int64).view(H, W) % 251).to(dtype)
            win = torch.full((hi - lo, W), 99, dtype=dtype)
            win[plan.own_lo[rank] - lo: plan.own_hi[rank] - lo] = truth[plan.own_lo[rank]: plan.own_hi[rank]]
            if dtype in (torch.uint8, torch.int32):   # merge callback path
                comm.exchange(plan, lambda s, a, b: win[a - lo: b - lo], lambda s, a, b, src: win[a - lo: b - lo].copy_(src))
            else:                                      # plain refresh: received straight into the halo rows
                comm.exchange(plan, [(lambda s, a, b: win[a - lo: b - lo], None)])
            assert torch.equal(win, truth[lo:hi]), f"halo exchange rank {rank} {dtype}"
        # root counts -> id bases: rank r owns 10 + r roots and sees 3 * r roots below its own rows
        own = torch.tensor([10 + rank], dtype=torch.int64)
        g = comm.all_gather({rank: own})[rank].reshape(-1)
        assert g.tolist() == [10 + r for r in range(world)]
        base = id_bases(g, torch.tensor(3 * rank), rank)
        assert int(base) == sum(10 + r for r in range(rank)) - 3 * rank
        # fixed-point flag: only the last rank reports a change
        assert comm.any({rank: torch.tensor([1 if rank == world - 1 else 0], dtype=torch.int32)}) is True
        assert comm.any({rank: torch.tensor([0], dtype=torch.int32)}) is False
        out.put((rank, "ok"))
    except Exception as e:   # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shard_comm_under_gloo(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 96, 40, 8, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def test_shard_plan_geometry():
    p = ShardPlan(1000, 4, 64)
    assert p.own_lo == [0, 250, 500, 750] and p.own_hi == [250, 500, 750, 1000]
    assert p.win_lo == [0, 186, 436, 686] and p.win_hi == [314, 564, 814, 1000]
    assert p.up_rows(0) == (186, 250) and p.down_rows(1) == (250, 314)
    with pytest.raises(ValueError):
        ShardPlan(100, 4, 64)


def test_local_comm_matches_distributed_semantics():
    """All shards in one process (how the GPU tests emulate N shards on one device)."""
    H, W, n, halo = 90, 16, 3, 5
    plan, comm = ShardPlan(H, n, halo), Comm(n, False)
    truth = torch.arange(H * W, dtype=torch.int32).view(H, W)
    wins = {}
    for s in range(n):
        w = torch.full((plan.win_hi[s] - plan.win_lo[s], W), -1, dtype=torch.int32)
        w[plan.own_lo[s] - plan.win_lo[s]: plan.own_hi[s] - plan.win_lo[s]] = truth[plan.own_lo[s]: plan.own_hi[s]]
        wins[s] = w
    comm.exchange(plan, lambda s, a, b: wins[s][a - plan.win_lo[s]: b - plan.win_lo[s]],
                  lambda s, a, b, src: wins[s][a - plan.win_lo[s]: b - plan.win_lo[s]].copy_(src))
    for s in range(n):
        assert torch.equal(wins[s], truth[plan.win_lo[s]: plan.win_hi[s]])
    g = comm.all_gather({s: torch.tensor([s + 1]) for s in range(n)})
    assert all(g[s].reshape(-1).tolist() == [1, 2, 3] for s in range(n))
    assert comm.any({s: torch.tensor(0) for s in range(n)}) is False
    assert comm.any({0: torch.tensor(0), 1: torch.tensor(1), 2: torch.tensor(0)}) is True


def _tick_message_worker(rank, world, port, H, halo, out):
    """The sharded tick's exchange (ShardedTraffic._message through Comm.exchange): one message per neighbour, different
    lengths per direction, received in place -- with CPU tensors standing in for the device message buffers."""
    import torch.distributed as dist
    from trafficsimulation_b200.sharded_traffic import ShardedTraffic
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class Stub:
            pass
        st = Stub()
        st.plan, st.n = ShardPlan(H, world, halo), world
        comm = Comm(world, True)
        # message from shard a to shard b has 100 + 10 a + b words, every word = 1000 a + b
        length = lambda a, b: 100 + 10 * a + b
        mk = lambda a, b: torch.full((length(a, b),), 1000 * a + b, dtype=torch.int32)
        st.send = {rank: [mk(rank, rank - 1) if rank > 0 else None, mk(rank, rank + 1) if rank + 1 < world else None]}
        st.recv = {rank: [torch.zeros(length(rank - 1, rank), dtype=torch.int32) if rank > 0 else None,
                          torch.zeros(length(rank + 1, rank), dtype=torch.int32) if rank + 1 < world else None]}
        st._role = lambda s, lo, hi: ShardedTraffic._role(st, s, lo, hi)
        for _ in range(3):   # every tick reuses the buffers
            comm.exchange(st.plan, lambda s, lo, hi: ShardedTraffic._message(st, s, lo, hi))
            if rank > 0:
                assert torch.equal(st.recv[rank][0], mk(rank - 1, rank)), "message from the shard below"
            if rank + 1 < world:
                assert torch.equal(st.recv[rank][1], mk(rank + 1, rank)), "message from the shard above"
            for b in st.recv[rank]:
                if b is not None:
                    b.zero_()
        out.put((rank, "ok"))
    except Exception as e:   # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tick_messages_under_gloo(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tick_message_worker, args=(r, world, port, 600, 128, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
