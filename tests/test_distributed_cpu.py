"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous sharding, the all-gather of
partial results, rank-order reassembly.  The per-rank compute is stood in for by the CPU oracle, the
plumbing under test is bulletproof-perm_b200/parallel.py (what bench.py --gpus N uses)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bpperm_b200
        from oracle import cref
        par = bpperm_b200.parallel
        rs = np.random.RandomState(5)
        blobs = rs.randint(0, 256, size=(n, 64), dtype=np.uint8).tobytes()
        sc = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
        sc[:, 31] &= 0x0F
        scb = sc.tobytes()
        pts = cref.from_uniform(blobs)
        # --- sharded MSM: partial per rank -> all-gather -> sum
        off, cnt = par.shard_bounds(n, world, rank)
        part = cref.msm_raw(scb[32 * off:32 * (off + cnt)], pts[160 * off:160 * (off + cnt)])
        allp = par.gather_bytes(torch.frombuffer(bytearray(part), dtype=torch.uint8), world)
        import ctypes
        acc = bytes(allp[:160].numpy().tobytes())
        for r in range(1, world):
            o = ctypes.create_string_buffer(160)
            cref.lib().orc_point_add(acc, bytes(allp[160 * r:160 * (r + 1)].numpy().tobytes()), o)
            acc = o.raw
        full = cref.msm(scb, pts)
        ok_msm = cref.compress(acc) == full
        # --- sharded batch: each rank decides its slice, decisions gathered in rank order
        total = 11
        truth = bytes((i * 7) % 3 != 0 for i in range(total))
        off2, cnt2 = par.shard_bounds(total, world, rank)
        mine = torch.zeros((total + world - 1) // world, dtype=torch.uint8)
        mine[:cnt2] = torch.tensor(list(truth[off2:off2 + cnt2]), dtype=torch.uint8)
        allacc = par.gather_bytes(mine, world)
        per = mine.numel()
        rebuilt = b"".join(bytes(allacc[per * r: per * r + par.shard_bounds(total, world, r)[1]].numpy().tobytes())
                           for r in range(world))
        q.put((rank, ok_msm, rebuilt == truth))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 257])
def test_sharded_msm_and_batch_over_gloo(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res)


def test_shard_bounds_cover_everything():
    import bpperm_b200
    sb = bpperm_b200.parallel.shard_bounds
    for n in (0, 1, 7, 8, 4096, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            spans = [sb(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
