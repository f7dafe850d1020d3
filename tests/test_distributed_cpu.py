"""World-size-2 gloo tests (CPU): the protocol of the sharded exchange build (helfem_b200/csrc/engine.cu) --
ownership by hfq_shard_assign (identical on every rank), each rank fills only its own segment of the compact
buffer, ONE all-gather completes it, every rank unpacks the complete matrix.  The compute stand-in is the C
oracle (the CUDA engine follows the same rule with finer units); plus the legacy compact all-reduce."""
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
    from helfem_b200.dist import shard_assign
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ob = cases.oracle_diatomic(3, 1, 1.8, (2, 2), 2)
    C = cjk.DiatomicCaches.from_oracle(ob)
    P = cases.random_density(ob.Nbf(), 3, 5, cases.m_blocks(ob.mval, ob.Nrad(), True))
    Pd = C.expand(P)
    na, N = C.Nang, C.Nrad
    units = [(j, k) for j in range(na) for k in range(na)]           # unit = output block
    cost = [1.0 + abs(int(C.lval[j]) - int(C.lval[k])) + 0.1 * j for j, k in units]
    owner = shard_assign(cost, world)
    assert sorted(set(owner)) == list(range(world))
    # segment layout: units of a rank contiguous, all segments of equal length
    fill = [0] * world
    off = []
    for u, o in enumerate(owner):
        off.append(fill[o]); fill[o] += N * N
    seg = max(fill)
    mine = [u for u in range(len(units)) if owner[u] == rank]
    blk = C.exchange_blocks(Pd, [units[u][0] for u in mine], [units[u][1] for u in mine])
    Kc = torch.zeros(world * seg, dtype=torch.float64)
    for b, u in enumerate(mine):
        Kc[rank * seg + off[u]: rank * seg + off[u] + N * N] = torch.from_numpy(np.ascontiguousarray(blk[b]).ravel())
    parts = list(Kc.view(world, seg).unbind(0))
    dist.all_gather(parts, Kc[rank * seg:(rank + 1) * seg].clone())   # the single collective of the sharded build
    Kc = torch.stack(parts).view(-1).numpy()
    K = np.zeros((na * N, na * N))
    for u, (j, k) in enumerate(units):                                # unpack, on every rank
        o = owner[u] * seg + off[u]
        K[j * N:(j + 1) * N, k * N:(k + 1) * N] = Kc[o:o + N * N].reshape(N, N)
    pi = C.pure_idx()
    err = cases.relerr(K[np.ix_(pi, pi)], ob.exchange(P))
    assert err < 1e-13, err
    # every rank derived the same ownership
    o_all = [torch.zeros(len(owner), dtype=torch.int32) for _ in range(world)]
    dist.all_gather(o_all, torch.from_numpy(owner.astype(np.int32)))
    assert all(torch.equal(o_all[0], x) for x in o_all)
    if rank == 0:
        print("OK", err)
    dist.destroy_process_group()
""")


def test_sharded_exchange_allgather_gloo(tmp_path):
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


def test_shard_assign_is_balanced_and_deterministic(hb):
    import numpy as np
    from helfem_b200.dist import shard_assign
    rng = np.random.default_rng(0)
    cost = rng.uniform(1.0, 10.0, 156)
    for nr in (1, 2, 4, 8):
        own = shard_assign(cost, nr)
        assert np.array_equal(own, shard_assign(cost, nr))
        load = np.array([cost[own == r].sum() for r in range(nr)])
        assert load.max() <= cost.sum() / nr + cost.max()        # LPT bound
        assert load.max() / load.mean() < 1.05


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
