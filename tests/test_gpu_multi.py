"""Multi-GPU build through the library's own communicator (hfq_comm_init: owner-computes sharding + one in-place
ncclAllGather issued from C++).  Needs >= 2 GPUs on the box (gpurun --gpus 2); skipped otherwise."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %r)
    import helfem_b200 as hb
    from tests import cases
    from tests.test_gpu_parity_large import gu_density
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    tabs = hb.Tables.diatomic(7, 7, 2.07, [18, 17], 3)
    single = hb.TablesBasis(tabs, device=local)
    multi = hb.TablesBasis(tabs, device=local).comm_init()
    n = tabs.Nbf
    for P in (gu_density(tabs, 3), cases.random_density(n, 3, 4, cases.m_blocks(tabs.mval, tabs.Nrad, True)),
              cases.random_density(n, 3, 5)):
        dP = torch.from_numpy(np.ascontiguousarray(P.T)).cuda()
        ref = [torch.empty_like(dP) for _ in range(2)]
        out = [torch.empty_like(dP) for _ in range(2)]
        single.coulomb_exchange_device(dP.data_ptr(), ref[0].data_ptr(), ref[1].data_ptr(), 0.5)
        multi.coulomb_exchange_device(dP.data_ptr(), out[0].data_ptr(), out[1].data_ptr(), 0.5)
        torch.cuda.synchronize()
        for a, b, name in zip(out, ref, "JK"):
            err = float((a - b).norm() / b.norm())
            assert err < 1e-13, (name, err, rank)
    dist.barrier()
    if rank == 0:
        print("OK")
    dist.destroy_process_group()
""")


def test_comm_sharded_build_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    nproc = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % nproc, "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "OK" in r.stdout
