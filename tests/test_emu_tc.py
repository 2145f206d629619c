"""tcgen05 kernels on the CPU emulator (tests/emu + csrc/mx_tc.cuh's restated primitives): the emulation decodes shared-memory descriptors
with the convention the B200 runs validated (tests/test_gpu_tc.py), so these tests check the kernels' INDEXING -- operand tiles, descriptor
strides, TMEM lanes / columns, barrier phases -- without a GPU.  Async-proxy ordering is outside what an emulator can see."""
import numpy as np
import pytest
import torch

import qmix_checks as qc


def probe(capi, M, N, K, passes, swap):
    lib = capi.lib()
    g = torch.Generator().manual_seed(M * 7 + N)
    X = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) * 0.3
    Y = torch.full((M, N), float("nan"))
    capi.check(lib.mx_tc_linear_probe(capi.ptr(X), capi.ptr(W), capi.ptr(Y), M, N, K, passes, swap, None))
    ref = X.double() @ W.double().t()
    return float((Y.double() - ref).abs().max() / ref.abs().max())


def test_3xtf32_building_block_matches_fp64(emu_engine):
    for (M, N, K) in [(128, 16, 8), (300, 64, 64), (100, 256, 64), (130, 192, 64)]:
        e3 = probe(emu_engine, M, N, K, 3, 0)
        e1 = probe(emu_engine, M, N, K, 1, 0)
        assert e3 < 2e-6, (M, N, K, e3)
        assert 1e-5 < e1 < 5e-3, (M, N, K, e1)
    assert probe(emu_engine, 128, 64, 64, 3, 1) > 1e-2          # the other stride assignment reads the wrong core matrices


@pytest.mark.parametrize("front_tc", [1, 0])
def test_qmix_step_front_paths_match_reference_golden(emu_engine, front_tc):
    lib = emu_engine.lib()
    lib.mx_set_option(b"front_tc", front_tc)
    try:
        qc.check_step_against(None, "qmix_5ag")
    finally:
        lib.mx_set_option(b"front_tc", 1)
