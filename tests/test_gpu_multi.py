"""Multi-GPU parity (SURVEY 8 row e), through the same code the benchmark uses: the library's NCCL communicator
(bpp_comm_init), bulletproof-perm_b200/parallel.py ShardedMsm and Batch.gather_accept.  One process per GPU;
skipped on a box with fewer than two devices (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).
Results are compared with the C restatement of the reference's algorithms on the UNSHARDED inputs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sb(v):
    return b"".join(int(x).to_bytes(32, "little") for x in v)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = {"rank": rank}
    try:
        import bpperm_b200
        from oracle import cref
        par = bpperm_b200.parallel
        G = bpperm_b200.acproof
        be = bpperm_b200.Backend(rank)
        be.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        par.init_comm(be, world, rank, dev)
        res["comm"] = be.comm_info() == (world, rank)
        # ---- sharded MSM: uniform scalars at two sizes, and a skewed set (one scalar repeated: hot buckets on one rank)
        ok_msm = []
        for n, skew in ((1 << 14, False), (1 << 16, False), (1 << 14, True), (1000, False)):
            rs = np.random.RandomState(100 + n + skew)
            blobs = rs.randint(0, 256, size=(n, 64), dtype=np.uint8)
            sc = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
            sc[:, 31] &= 0x0F
            if skew:
                sc[: n // 3] = sc[0]
            off, cnt = par.shard_bounds(n, world, rank)
            table = be.points_from_uniform(blobs[off:off + cnt].tobytes())
            d_sc = torch.from_numpy(sc[off:off + cnt].copy()).to(dev)
            sm = par.ShardedMsm(be, table, world, dev)
            one = bytes(sm.run(d_sc)[:32].cpu().numpy().tobytes())
            many = sm.run_many([d_sc] * 5)
            torch.cuda.synchronize()
            outs = [bytes(o[:32].cpu().numpy().tobytes()) for o in sm.d_outs]
            want = cref.msm(sc.tobytes(), cref.from_uniform(blobs.tobytes()))
            ok_msm.append(one == want and all(o == want for o in outs) and bytes(many[:32].cpu().numpy().tobytes()) == want)
            table.free()
        res["msm"] = ok_msm
        # ---- sharded batch verification with corrupted proofs: decisions vs the C restatement, proof by proof
        k, total = 4, 37
        W = bpperm_b200.weights
        n, Q, m, WL, WR, WO, WV, c = W.shuffle_circuit(k)
        ng = G.next_pow2(n)
        rs = np.random.RandomState(5)
        enc = cref.compress(cref.from_uniform(rs.randint(0, 256, size=(2 * ng + 2, 64), dtype=np.uint8).tobytes()))
        gl = [enc[32 * i:32 * i + 32] for i in range(2 * ng + 2)]
        inst = cref.AcpFixedInstance(n, Q, m, WL, WR, WO, WV, _sb(c), gl[0], gl[1], enc[64:64 + 32 * ng], enc[64 + 32 * ng:])
        cir = G.Circuit.shuffle(be, k)
        gens = G.Generators(be, gl[0], gl[1], gl[2:2 + ng], gl[2 + ng:], 6)
        deck = np.zeros((k, 32), dtype=np.uint8)
        deck[:, 0] = np.arange(1, k + 1)
        perm = np.stack([rs.permutation(k) for _ in range(total)]).astype(np.uint32)
        x = rs.randint(0, 256, size=(total, 32), dtype=np.uint8)
        x[:, 31] = 0
        gamma = rs.randint(0, 256, size=(total * m, 32), dtype=np.uint8)
        gamma[:, 31] &= 0x0F
        seeds = rs.randint(0, 256, size=(total, 32), dtype=np.uint8)
        off, cnt = par.shard_bounds(total, world, rank)
        per = (total + world - 1) // world
        batch = G.Batch(be, cir, gens, cnt, "fixed")
        batch.gen_shuffle_witness(deck.tobytes(), perm[off:off + cnt].tobytes(), x[off:off + cnt].tobytes(),
                                  gamma[m * off:m * (off + cnt)].tobytes(), seeds[off:off + cnt].tobytes())
        V = batch.commit(None)
        batch.prove()
        proofs = bytearray(batch.download_proofs())
        plen = batch.proof_len
        want = []
        for g in range(off, off + cnt):
            i = g - off
            if g % 5 == 2:
                proofs[i * plen + 32 * (g % (plen // 32)) + 7] ^= 0x10     # one bit in a different field per victim
            Ve = V[32 * m * i:32 * m * (i + 1)]
            want.append(1 if inst.verify(bytes(proofs[i * plen:(i + 1) * plen]), cref.decompress(Ve), V_enc=Ve) else 0)
        batch.upload_proofs(bytes(proofs), V)
        batch.verify()
        allacc = batch.gather_accept(per)
        mine = list(allacc[per * rank:per * rank + cnt])
        res["verify"] = mine == want and 0 in want and 1 in want
        # every rank holds every rank's decisions: compare the other ranks' slices over gloo
        t = torch.zeros(per, dtype=torch.uint8)
        t[:cnt] = torch.tensor(want, dtype=torch.uint8)
        ref = par.gather_bytes(t, world)
        res["gathered"] = bytes(ref.numpy().tobytes()) == allacc
        batch.free(); gens.free(); cir.free()
        be.close()
    except Exception as e:  # pragma: no cover
        import traceback
        res["error"] = traceback.format_exc()
    finally:
        q.put(res)
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_msm_and_sharded_batch_verify_on_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert "error" not in r, r.get("error")
        assert r["comm"] and all(r["msm"]) and r["verify"] and r["gathered"], r
