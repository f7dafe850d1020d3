"""World-size-2 gloo test (CPU): the reduction protocol of the sharded exchange build --
ranks compute disjoint round-robin subsets of the work, one all-reduce(sum) gives the full K.
The compute stand-in is the C oracle (the CUDA engine shards by the same rule on the GPU)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %r)
    from tests import cases
    from oracle import cjk
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ob = cases.oracle_diatomic(3, 1, 1.8, (2, 2), 2)
    C = cjk.DiatomicCaches.from_oracle(ob)
    P = cases.random_density(ob.Nbf(), 3, 5, cases.m_blocks(ob.mval, ob.Nrad(), True))
    Pd = C.expand(P)
    na, N = C.Nang, C.Nrad
    pairs = [(j, k) for j in range(na) for k in range(na)]
    mine = pairs[rank::world]                      # round-robin deal, like the engine's task sharding
    blk = C.exchange_blocks(Pd, [p[0] for p in mine], [p[1] for p in mine])
    K = np.zeros((na * N, na * N))
    for b, (j, k) in enumerate(mine):
        K[j * N:(j + 1) * N, k * N:(k + 1) * N] = blk[b]
    Kt = torch.from_numpy(K)
    dist.all_reduce(Kt)                            # the single collective of the sharded build
    pi = C.pure_idx()
    Kfull = Kt.numpy()[np.ix_(pi, pi)]
    err = cases.relerr(Kfull, ob.exchange(P))
    assert err < 1e-13, err
    if rank == 0:
        print("OK", err)
    dist.destroy_process_group()
""")


def test_sharded_exchange_allreduce_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout


WORKER_CAR = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %r)
    from helfem_b200.dist import CompactAllReduce
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()

    class FakeBasis:                      # what CompactAllReduce needs from a basis: the output pattern
        def exchange_output_pattern(self, coulomb=False):
            bf_sector = np.array([0, 0, 1, 2, 2, 2, 1, 0], dtype=np.int32)      # interleaved sectors
            pairs = [(0, 0), (1, 2), (2, 1)] if not coulomb else [(2, 2)]
            return bf_sector, pairs

    n = 8
    for coulomb in (False, True):
        car = CompactAllReduce(FakeBasis(), torch.device("cpu"), coulomb=coulomb)
        rng = np.random.default_rng(10 + rank)
        K = rng.standard_normal((n, n))                       # K[r, c]; the tensor holds the column-major image
        dK = torch.from_numpy(np.ascontiguousarray(K.T)).clone()
        before = dK.clone()
        car(dK)
        # expected: entries of the pattern blocks are summed over ranks, everything else is untouched
        bf, pairs = FakeBasis().exchange_output_pattern(coulomb)
        mask = np.zeros((n, n), dtype=bool)
        for sj, sk in pairs:
            mask[np.ix_(bf == sj, bf == sk)] = True
        tot = sum(np.random.default_rng(10 + r).standard_normal((n, n)) for r in range(world))
        got = dK.numpy().T
        assert np.allclose(got[mask], tot[mask], rtol=0, atol=1e-14)
        assert np.array_equal(got[~mask], before.numpy().T[~mask])
        assert car.nbytes == int(mask.sum()) * 8
    if rank == 0:
        print("OK")
    dist.destroy_process_group()
""")


def test_compact_allreduce_gloo(tmp_path):
    """helfem_b200/dist.py on CPU tensors: only the blocks of the output pattern take part in the collective."""
    script = tmp_path / "worker_car.py"
    script.write_text(WORKER_CAR % ROOT)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout
