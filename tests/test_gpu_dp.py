"""Data-parallel learner on 2 GPUs: gradient exchange over NVLink peer memory (csrc/p2p.cu) and, for comparison, over NCCL.
Each rank trains on half of the golden batch; both exchanges must reproduce the reference's single-batch step, leave the
replicas bit-identical, and agree with each other to summation order.  Skipped on a single-GPU box."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(p2p):
    with tempfile.TemporaryDirectory() as td:
        port = 29600 + os.getpid() % 300 + (50 if p2p else 0)
        procs = []
        for r in range(2):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                       MARL_B200_P2P="1" if p2p else "0")
            procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dp_worker_gpu.py"), td], env=env,
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
        outs = [p.communicate(timeout=300)[0].decode() for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(outs)
        return [dict(np.load(os.path.join(td, "rank%d.npz" % r))) for r in range(2)], outs


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_step_peer_memory_and_nccl():
    g = load_golden("qmix_small")
    steps = int(g["meta.cfg"][-1])
    res = {}
    for p2p in (True, False):
        (r0, r1), outs = _run(p2p)
        assert int(r0["p2p"]) == int(p2p), "\n".join(outs)
        for s in range(steps):
            assert float(r0["s%d.timeout" % s]) == 0.0
            for k in ("loss", "grad_norm", "Q_tot"):
                assert float(r0["s%d.%s" % (s, k)]) == float(r1["s%d.%s" % (s, k)])
                want = float(g["s%d.%s" % (s, k)])
                assert abs(float(r0["s%d.%s" % (s, k)]) - want) <= 1e-4 * abs(want), (p2p, s, k)
        assert np.array_equal(r0["theta"], r1["theta"]) and np.array_equal(r0["theta_tgt"], r1["theta_tgt"])      # replicas bit-identical
        res[p2p] = r0
    assert np.abs(res[True]["theta"] - res[False]["theta"]).max() <= 1e-5
