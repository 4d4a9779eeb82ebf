"""Multi-GPU plumbing (one process per GPU, torch.distributed): only what SURVEY 8(e) shards.

* proofs are independent units  -> contiguous slices per rank, no data-path collective; accept bytes can be
  gathered with `gather_bytes` when one rank needs all decisions;
* a large MSM shards its points  -> every rank reduces its slice to one extended point (128 B), a single
  all-gather of those partials, then every rank adds them and compresses.
The collective payloads are tiny (128 B per rank): latency, not bandwidth, is what matters.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous split of n units over `world` ranks: (offset, count); the first n % world ranks get one more."""
    base, rem = divmod(n, world)
    cnt = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return off, cnt


def gather_bytes(local: torch.Tensor, world: int) -> torch.Tensor:
    """All-gather a fixed-size uint8 tensor; result is [world * len(local)] in rank order."""
    if world == 1:
        return local.clone()
    out = torch.empty(world * local.numel(), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


class ShardedMsm:
    """MSM over a point set sharded across ranks.  `table` holds this rank's slice on its GPU."""

    PARTIAL_BYTES = 128

    def __init__(self, backend, table, world: int, device):
        self.be, self.table, self.world = backend, table, world
        self.d_part = torch.zeros(self.PARTIAL_BYTES, dtype=torch.uint8, device=device)
        self.d_all = torch.zeros(world * self.PARTIAL_BYTES, dtype=torch.uint8, device=device)
        self.d_out = torch.zeros(160, dtype=torch.uint8, device=device)

    def run(self, d_scalars: torch.Tensor) -> torch.Tensor:
        """d_scalars: this rank's scalars (n_local x 32, uint8, on the GPU).  Returns the device buffer whose
        first 32 bytes are the compressed result (identical on every rank)."""
        n = len(self.table)
        if self.world == 1:
            self.be.msm_dev(d_scalars.data_ptr(), self.table, 0, n, self.d_out.data_ptr())
            return self.d_out
        self.be.msm_partial_dev(d_scalars.data_ptr(), self.table, 0, n, self.d_part.data_ptr())
        dist.all_gather_into_tensor(self.d_all, self.d_part)
        self.be.points_sum_compress_dev(self.d_all.data_ptr(), self.world, self.d_out.data_ptr())
        return self.d_out

    # Throughput form for a sequence of independent MSMs (two in flight per rank, bpp_msm_submit_*), three buffer
    # slots (i % 3):
    #     for i, sc in enumerate(sets):
    #         m.submit(sc, i % 3); m.wait_previous()
    #         if i >= 1: m.gather((i - 1) % 3)       # 128 B/rank all-gather + sum + compress on a side stream
    #         if i >= 2: m.finish((i - 2) % 3)       # event wait only: that step's result is final
    #     m.wait(); m.gather(last % 3); m.finish((last - 1) % 3); m.finish(last % 3)
    # Nothing but event waits goes onto the caller's stream between two submits: a kernel there (even the one-warp
    # sum) queues behind the accumulate blocks of the MSM in flight and would hold back the next submit's fork.  The
    # gather of step i-1 and its sum + compress run on a side stream (the sum through a second context bound to it)
    # beside the MSM of step i; finish() only makes the caller's stream wait for them.
    SLOTS = 3

    def _slots(self):
        if not hasattr(self, "_slot_bufs"):
            dev = self.d_out.device
            self._slot_bufs = [(torch.zeros(self.PARTIAL_BYTES, dtype=torch.uint8, device=dev),
                                torch.zeros(self.world * self.PARTIAL_BYTES, dtype=torch.uint8, device=dev),
                                torch.zeros(160, dtype=torch.uint8, device=dev)) for _ in range(self.SLOTS)]
            self._gathered = [torch.cuda.Event() for _ in range(self.SLOTS)]
            self._ready = torch.cuda.Event()
            self._side = torch.cuda.Stream(dev)
            if self.world > 1:
                self._side_be = type(self.be)(dev.index if dev.index is not None else 0)
                self._side_be.set_stream(self._side.cuda_stream)
        return self._slot_bufs

    def submit(self, d_scalars: torch.Tensor, slot: int):
        part, _, out = self._slots()[slot]
        n = len(self.table)
        if self.world == 1:
            self.be.msm_submit_dev(d_scalars.data_ptr(), self.table, 0, n, out.data_ptr())
        else:
            self.be.msm_submit_partial_dev(d_scalars.data_ptr(), self.table, 0, n, part.data_ptr())

    def wait_previous(self):
        self.be.msm_wait_previous()

    def wait(self):
        self.be.msm_wait()

    def gather(self, slot: int):
        """After the slot's MSM has been waited for on the caller's stream: all-gather of the partials on the side
        stream (nothing to do on one GPU)."""
        if self.world == 1:
            return
        part, allp, _ = self._slots()[slot]
        cur = torch.cuda.current_stream(part.device)
        self._ready.record(cur)
        self._side.wait_event(self._ready)
        _, _, out = self._slots()[slot]
        with torch.cuda.stream(self._side):
            dist.all_gather_into_tensor(allp, part)
            self._side_be.points_sum_compress_dev(allp.data_ptr(), self.world, out.data_ptr())
            self._gathered[slot].record(self._side)

    def finish(self, slot: int) -> torch.Tensor:
        """The caller's stream waits for the slot's gather + sum.  Returns the slot's 160-byte output buffer (first
        32 bytes = the compressed result, identical on every rank)."""
        _, allp, out = self._slots()[slot]
        if self.world > 1:
            torch.cuda.current_stream(out.device).wait_event(self._gathered[slot])
        return out

    def close(self):
        """Release the side context of the throughput form (multi-GPU only)."""
        be2 = getattr(self, "_side_be", None)
        if be2 is not None:
            torch.cuda.synchronize()
            be2.close()
            self._side_be = None

    def run_many(self, scalar_sets):
        """The loop above over a list of device scalar tensors; returns the output buffer of the last step."""
        k = len(scalar_sets)
        for i, sc in enumerate(scalar_sets):
            self.submit(sc, i % 3)
            self.wait_previous()
            if i >= 1:
                self.gather((i - 1) % 3)
            if i >= 2:
                self.finish((i - 2) % 3)
        self.wait()
        self.gather((k - 1) % 3)
        if k >= 2:
            self.finish((k - 2) % 3)
        return self.finish((k - 1) % 3)
