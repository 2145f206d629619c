"""Replay parity on the real sm_100a kernels (through the C-ABI) -- same checks as the emulated module,
plus full-size properties."""
import numpy as np
import pytest
import torch

import replay_checks as rc

pytestmark = pytest.mark.gpu


def test_uniform_golden(gpu_engine):
    rc.check_uniform_golden()


def test_device_rng_stream(gpu_engine):
    rc.check_device_rng_stream()
    rc.check_device_rng_stream(seed=2**32 - 1, E=5000, B=32, rounds=40)


def test_per_golden(gpu_engine):
    rc.check_per_golden()


def test_per_vs_oracle_random(gpu_engine):
    rc.check_per_vs_oracle_random()
    rc.check_per_vs_oracle_random(seed=9, E=700, B=16, steps=8)


@pytest.mark.parametrize("shape", [(3, 30, 9, 48, 5, 11, 4, True), (2, 5, 3, 7, 3, 6, 6, False), (1, 1, 2, 1, 1, 3, 2, True),
                                   (8, 80, 14, 168, 20, 40, 16, True)])
def test_uniform_vs_oracle(gpu_engine, shape):
    rc.check_uniform_vs_oracle_shapes(shape)


def test_reward_norm_running_statistics(gpu_engine):
    rc.check_reward_norm_running_stats()
    rc.check_reward_norm_running_stats(seed=8, E=50, B=16, inserts=130)


def test_full_size_gather_properties(gpu_engine):
    """BASELINE config 4 shapes (8m: N=8, O=80, A=14, S=168, T=120, B=64): every sampled episode's bytes equal the
    stored bytes (round trip through insert -> sample), indices follow np.random.choice, duplicates allowed."""
    N, O, A, S, T, E, B = 8, 80, 14, 168, 120, 300, 64
    buf = rc.make_buffers(N, O, A, S, T, E, rng="device", max_batch=B)
    rs = np.random.RandomState(1)
    kept = {}
    for chunk in range(E // 50):
        n = 50
        ep = [rs.randn(T + 1, n, N, O), np.repeat(rs.randn(T + 1, n, 1, S), N, 2), np.eye(A)[rs.randint(0, A, (T, n, N))],
              rs.randn(T, n, N, 1), np.zeros((T, n, N, 1)), np.zeros((T, n, 1)), (rs.rand(T + 1, n, N, A) < 0.7) * 1.0]
        ep = [x.astype(np.float32) for x in ep]
        slots = buf.insert(n, *[rc.d(x) for x in ep])
        for j, s in enumerate(slots):
            kept[int(s)] = [x[:, j] for x in ep]
    buf.seed_device_rng(4)
    np.random.seed(4)
    for _ in range(3):
        smp = buf.sample(B)
        idx = np.asarray(buf.policy_buffers["policy_0"].sampled_indices(B))
        assert np.array_equal(idx, np.random.choice(E, B))
        obs, share, acts, rew, dones, de, av = [smp[i]["policy_0"] for i in range(7)]
        for b, e in enumerate(idx):
            o, s, a, r, d, den, avl = kept[int(e)]
            assert np.array_equal(obs[:, :, b], o.transpose(1, 0, 2))
            assert np.array_equal(share[:, b], s[:, 0])
            assert np.array_equal(acts[:, :, b], a.transpose(1, 0, 2))
            assert np.array_equal(rew[:, :, b], r.transpose(1, 0, 2))
            assert np.array_equal(de[:, b], den)
            assert np.array_equal(av[:, :, b], avl.transpose(1, 0, 2))


@pytest.mark.parametrize("shape", [(3, 30, 9, 48, 60, 300, 32, False), (8, 80, 14, 168, 120, 200, 64, False), (3, 18, 5, 54, 25, 100, 32, True), (5, 7, 3, 9, 7, 40, 16, True)])
def test_tma_gather_equals_vectorised_gather(gpu_engine, shape):
    """k_gather_tma (tensor-map tile loads / stores, csrc/gather_tma.cu) against k_gather (16-byte vector loads): the whole batch region must be
    bit-identical for the same indices -- full boxes, the clipped last box of every field, and the reward-normalisation transform."""
    N, O, A, S, T, E, B, norm = shape
    lib = gpu_engine.lib()
    outs = []
    for tma in (1, 0):
        lib.mx_set_option(b"gather_tma", tma)
        try:
            buf = rc.make_buffers(N, O, A, S, T, E, norm=norm, rng="numpy", max_batch=max(B, 64))
            rs = np.random.RandomState(5)
            for c in range(0, E, 50):
                n = min(50, E - c)
                ep = [rs.randn(T + 1, n, N, O), np.repeat(rs.randn(T + 1, n, 1, S), N, 2), np.eye(A)[rs.randint(0, A, (T, n, N))],
                      np.repeat(3.0 + rs.randn(T, n, 1, 1), N, 2), np.zeros((T, n, N, 1)), np.zeros((T, n, 1)), (rs.rand(T + 1, n, N, A) < 0.7) * 1.0]
                buf.insert(n, *[rc.d(x.astype(np.float32)) for x in ep])
            pb = buf.policy_buffers["policy_0"]
            pb.gather(np.random.RandomState(6).randint(0, E, B))
            torch.cuda.synchronize()
            lo, hi = int(pb.L.off_b_obs), int(pb.L.off_b_idx)
            outs.append(pb.blob[lo:hi].clone())
        finally:
            lib.mx_set_option(b"gather_tma", 1)
    assert torch.equal(outs[0], outs[1])
    assert float(outs[0].view(torch.float32)[:1024].abs().sum()) > 0.0
