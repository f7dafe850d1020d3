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
