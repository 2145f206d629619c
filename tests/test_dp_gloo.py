"""Data-parallel host logic on CPU: 2 gloo ranks, each training on half of the golden batch with the emulated kernels,
must reproduce the reference's single-batch step (loss, grad_norm, parameters) -- the all-reduced buffer carries
gradient NUMERATORS and the loss denominators, so sharding the batch is exact up to summation order."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode,golden", [("rnn", "qmix_small"), ("mlp", "mqmix_small")])
def test_two_rank_step_equals_single_batch_reference(emu_engine, mode, golden):
    g = load_golden(golden)
    with tempfile.TemporaryDirectory() as td:
        port = 29500 + (os.getpid() + (7 if mode == "mlp" else 0)) % 1000
        procs = []
        for r in range(2):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
            procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dp_worker.py"), td, mode], env=env,
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
        outs = [p.communicate(timeout=600)[0].decode() for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(outs)
        r0, r1 = np.load(os.path.join(td, "rank0.npz")), np.load(os.path.join(td, "rank1.npz"))
    for k in ("loss", "grad_norm", "Q_tot"):
        assert float(r0[k]) == float(r1[k])
        want = float(g["s0." + k])
        assert abs(float(r0[k]) - want) <= 1e-4 * abs(want), (k, float(r0[k]), want)
    assert np.array_equal(r0["theta"], r1["theta"])          # replicas stay bit-identical
    assert np.array_equal(r0["theta_tgt"], r1["theta_tgt"])
