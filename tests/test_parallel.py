"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard arithmetic, the weight-arena
broadcast and the waveform gather (the only two collectives on the path)."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_range_covers_everything():
    from voicefixer_b200.parallel import shard_range
    for n in (0, 1, 7, 32, 512, 513):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ragged, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from voicefixer_b200 import parallel
    parallel.init_from_env(backend="gloo")
    # --- arena broadcast: rank 0 owns the bytes and the layout table
    if rank == 0:
        table = [("a.w", 0, 10), ("b.w", 256, 7)]
        arena = torch.arange(300, dtype=torch.int64).to(torch.uint8)
    else:
        table, arena = None, None
    table, arena = parallel.broadcast_arena(table, arena, "cpu")
    ok = table == [("a.w", 0, 10), ("b.w", 256, 7)] and arena.numel() == 300 and int(arena[299]) == 299 % 256
    # --- batch shard + gather of "waveforms"
    n_items, L = (5, 11) if ragged else (6, 11)
    lo, hi = parallel.shard_range(n_items, rank, world)
    full = torch.arange(n_items * L, dtype=torch.float32).reshape(n_items, L)
    out = parallel.gather_waveforms(full[lo:hi] * 2.0)
    if rank == 0:
        ok = ok and out is not None and torch.equal(out, full * 2.0)
    else:
        ok = ok and out is None
    # --- the per-step form: sizes exchanged once at construction, then exactly ONE collective per call
    plan = parallel.WaveformGather(hi - lo, L, torch.float32, "cpu")
    ok = ok and plan.total_items == n_items and plan.sizes == [parallel.shard_range(n_items, r, world)[1] -
                                                             parallel.shard_range(n_items, r, world)[0] for r in range(world)]
    calls = []
    real_gather, real_all_gather = dist.gather, dist.all_gather
    dist.gather = lambda *a, **k: (calls.append("gather"), real_gather(*a, **k))[1]
    dist.all_gather = lambda *a, **k: (calls.append("all_gather"), real_all_gather(*a, **k))[1]
    for step in range(3):
        out = plan(full[lo:hi] * float(step + 1))
        if rank == 0:
            ok = ok and torch.equal(out, full * float(step + 1))
        else:
            ok = ok and out is None
    dist.gather, dist.all_gather = real_gather, real_all_gather
    ok = ok and calls == ["gather"] * 3
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ragged", [False, True])
def test_broadcast_and_gather_gloo_world2(ragged):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ragged, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
