"""Multi-GPU plumbing (one process per GPU).

The collective of a sharded exchange build is issued by the library itself (hfq_comm_init: owner-computes
sharding by output block + one in-place ncclAllGather of the compact result, helfem_b200/csrc/engine.cu).  What is
left for the host language is the bootstrap -- handing the 128-byte NCCL id from rank 0 to the others -- done here
through torch.distributed, and `CompactAllReduce`, the reduction for contexts WITHOUT a communicator (legacy
shard / nshards arguments: every rank holds a partial dense matrix whose sum is the result).
"""
import numpy as np
import torch
import torch.distributed as dist


def torch_bcast():
    """(rank, world, bcast) for _BasisBase.comm_init: broadcast of a bytes object from rank 0 over the default
    torch.distributed process group (any backend)."""
    rank, world = dist.get_rank(), dist.get_world_size()

    def bcast(data):
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    return rank, world, bcast


def shard_assign(costs, nranks):
    """Owner of every work unit (hfq_shard_assign: longest first onto the least loaded rank)."""
    import ctypes
    from . import lib, _check
    c = np.ascontiguousarray(costs, dtype=np.float64)
    out = np.zeros(len(c), dtype=np.int32)
    _check(lib().hfq_shard_assign(c.ctypes.data, len(c), int(nranks), out.ctypes.data))
    return out


class CompactAllReduce:
    """All-reduce(sum) of the non-zero blocks of a column-major Nbf x Nbf device matrix held in a
    torch tensor ``dK`` of shape (Nbf, Nbf) (dK[c, r] = K[r, c]).  The block pattern is the pattern of the
    COMPLETE matrix (identical on every rank, whatever the rank's share of the work) and is re-read from the
    context on every call, so a density whose block structure changes between SCF iterations is handled."""

    def __init__(self, basis, device, coulomb=False):
        self.basis, self.device, self.coulomb = basis, device, coulomb
        self.pattern, self.idx, self.nbytes = None, None, 0

    def _refresh(self):
        bf_sector, pairs = self.basis.exchange_output_pattern(self.coulomb)
        if self.pattern == tuple(pairs):
            return
        n = len(bf_sector)
        nsec = int(bf_sector.max()) + 1
        members = [np.nonzero(bf_sector == s)[0] for s in range(nsec)]
        idx = []
        for (sj, sk) in pairs:
            rows, cols = members[sj], members[sk]
            idx.append((cols[:, None].astype(np.int64) * n + rows[None, :]).ravel())
        flat = np.concatenate(idx) if idx else np.zeros(0, dtype=np.int64)
        self.pattern = tuple(pairs)
        self.idx = torch.from_numpy(flat).to(self.device)
        self.nbytes = int(flat.size) * 8

    def __call__(self, dK):
        self._refresh()
        if self.idx.numel() == 0:
            return
        flat = dK.view(-1)
        packed = flat[self.idx]
        dist.all_reduce(packed)
        flat[self.idx] = packed
