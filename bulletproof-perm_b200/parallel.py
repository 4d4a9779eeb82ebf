"""Multi-GPU plumbing (one process per GPU): only what SURVEY 8(e) shards, and all of it behind the C ABI.

* proofs are independent units  -> contiguous slices per rank (`shard_bounds`), no data-path collective; the accept
  bytes are all-gathered by the library (bpp_acp_batch_gather_accept) when one rank needs all decisions;
* a large MSM shards its points  -> every rank reduces its slice to one extended point (128 B), the library all-gathers
  those partials over NCCL, every rank adds them and compresses (bpp_msm_sharded_dev / _submit_dev / _wait).
The communicator lives in the library (bpp_comm_init, created once per context); this module only carries the 128-byte
id from rank 0 to the others over whatever channel the host program has - here torch.distributed, which the
benchmark and the tests use for their barriers anyway.  Payloads are tiny (128 B per rank): latency, not bandwidth.
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


def init_comm(backend, world: int, rank: int, device=None):
    """Create the library's NCCL communicator on `backend`: rank 0 makes the id, torch.distributed broadcasts it."""
    if world == 1:
        return
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(type(backend).comm_unique_id()), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        uid = uid.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.broadcast(uid, src=0)
    backend.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))


def gather_bytes(local: torch.Tensor, world: int) -> torch.Tensor:
    """All-gather a fixed-size uint8 tensor through torch.distributed (host-side tests on gloo); result is
    [world * len(local)] in rank order.  The GPU path uses the library: Backend.all_gather_dev / Batch.gather_accept."""
    if world == 1:
        return local.clone()
    out = torch.empty(world * local.numel(), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


class ShardedMsm:
    """MSM over a point set sharded across ranks.  `table` holds this rank's slice on its GPU; the backend carries the
    communicator (init_comm).  With world == 1 the same calls are the single-GPU MSM."""

    def __init__(self, backend, table, world: int, device):
        self.be, self.table, self.world = backend, table, world
        if backend.comm_info()[0] != world:
            raise RuntimeError(f"ShardedMsm over {world} ranks needs the library communicator: call init_comm(backend, world, rank) first")
        self.d_out = torch.zeros(160, dtype=torch.uint8, device=device)
        self.d_outs = [torch.zeros(160, dtype=torch.uint8, device=device) for _ in range(3)]

    def run(self, d_scalars: torch.Tensor) -> torch.Tensor:
        """d_scalars: this rank's scalars (n_local x 32, uint8, on the GPU).  Returns the device buffer whose
        first 32 bytes are the compressed result (identical on every rank)."""
        self.be.msm_sharded_dev(d_scalars.data_ptr(), self.table, 0, len(self.table), self.d_out.data_ptr())
        return self.d_out

    def run_many(self, scalar_sets):
        """Throughput form over a list of device scalar tensors: two MSMs in flight per rank, the 128-byte all-gather
        + sum of step i - 1 on the library's communication stream beside the MSM of step i.  Returns the output buffer
        of the last step (results rotate over three buffers: self.d_outs[i % 3])."""
        n = len(self.table)
        for i, sc in enumerate(scalar_sets):
            self.be.msm_sharded_submit_dev(sc.data_ptr(), self.table, 0, n, self.d_outs[i % 3].data_ptr())
        self.be.msm_sharded_wait()
        return self.d_outs[(len(scalar_sets) - 1) % 3]

    def close(self):
        pass
