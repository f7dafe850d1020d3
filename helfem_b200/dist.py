"""Multi-GPU plumbing for sharded exchange builds (one process per GPU, torch.distributed/NCCL).

Each rank runs ``exchange_device(..., shard=rank, nshards=world)`` and obtains a partial sum of
every non-zero block of K.  One all-reduce completes the matrix; it is restricted to the blocks
the engine actually wrote (``hfq_exchange_output_pattern``), which for the m-diagonal densities
of linear molecules is a few percent of the dense matrix.
"""
import numpy as np
import torch
import torch.distributed as dist


class CompactAllReduce:
    """All-reduce(sum) of the non-zero blocks of a column-major Nbf x Nbf device matrix held in a
    torch tensor ``dK`` of shape (Nbf, Nbf) (dK[c, r] = K[r, c])."""

    def __init__(self, basis, device, coulomb=False):
        bf_sector, pairs = basis.exchange_output_pattern(coulomb)
        n = len(bf_sector)
        nsec = int(bf_sector.max()) + 1
        members = [np.nonzero(bf_sector == s)[0] for s in range(nsec)]
        idx = []
        for (sj, sk) in pairs:
            rows, cols = members[sj], members[sk]
            idx.append((cols[:, None].astype(np.int64) * n + rows[None, :]).ravel())
        flat = np.concatenate(idx) if idx else np.zeros(0, dtype=np.int64)
        self.pattern = tuple(pairs)
        self.idx = torch.from_numpy(flat).to(device)
        self.nbytes = int(flat.size) * 8

    def __call__(self, dK):
        if self.idx.numel() == 0:
            return
        flat = dK.view(-1)
        packed = flat[self.idx]
        dist.all_reduce(packed)
        flat[self.idx] = packed
