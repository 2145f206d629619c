"""Replay parity checks shared by the emulated (CPU) and the real (GPU) test modules.
Bit-exact against the reference goldens and against oracle/replay.py."""
import numpy as np

from helpers import load_golden, FIELDS
from oracle.mt19937 import LegacyMT19937
from oracle.replay import UniformReplay, PrioritizedReplay

REF_FIELDS = ["obs", "share_obs", "acts", "rewards", "dones", "dones_env", "avail_acts"]


from offpolicy._b200.factory import Discrete, make_rec_buffers as make_buffers, pd as d  # noqa: E402,F401


class Box(object):      # duck-typed gym.spaces.Box for the constructor
    def __init__(self, d):
        self.shape = (d,)


def golden_insert(g, tag, j):
    f = [g["%s.ins%d.%s" % (tag, j, k)] for k in FIELDS]
    return [d(x) for x in f]


def check_uniform_golden():
    g = load_golden("replay_small")
    N, O, A, S, T, E = [int(v) for v in g["meta"]]
    for tag, norm in (("plain", False), ("norm", True)):
        buf = make_buffers(N, O, A, S, T, E, norm=norm)
        for j, n_ep in enumerate((5, 7, 9)):
            r = buf.insert(n_ep, *golden_insert(g, tag, j))
            assert np.array_equal(r, g["%s.idx_range%d" % (tag, j + 1)]), (tag, j)
        assert len(buf) == E
        np.random.seed(123)
        for k in range(3):
            smp = buf.sample(6)
            assert smp[7] is None and smp[8] is None
            idx = np.asarray(buf.policy_buffers["policy_0"].sampled_indices(6))
            assert np.array_equal(idx, g["%s.inds" % tag][k])
            for name, f in zip(FIELDS, smp[:7]):
                got, want = f["policy_0"], g["%s.draw%d.%s" % (tag, k, name)]
                assert got.shape == want.shape and got.dtype == want.dtype, (name, got.shape, want.shape)
                if name == "rew" and norm:
                    assert np.allclose(got, want, rtol=2e-6, atol=2e-6)   # fp32 mean/std summation order differs
                else:
                    assert np.array_equal(got, want), (tag, k, name)


def check_device_rng_stream(seed=321, E=37, B=16, rounds=45):
    """Device MT19937 == np.random.seed(seed); np.random.choice(...) across several state regenerations."""
    N, O, A, S, T = 2, 5, 3, 6, 3
    buf = make_buffers(N, O, A, S, T, E, rng="device", max_batch=B)
    rs = np.random.RandomState(0)
    ep = lambda n: [d(rs.randn(T + 1, n, N, O)), d(rs.randn(T + 1, n, N, S)), d(np.eye(A)[rs.randint(0, A, (T, n, N))]),
                    d(rs.randn(T, n, N, 1)), d(np.zeros((T, n, N, 1))), d(np.zeros((T, n, 1))), d(np.ones((T + 1, n, N, A)))]
    left = E
    while left > 0:
        n = min(B, left)
        buf.insert(n, *ep(n))
        left -= n
    assert len(buf) == E
    buf.seed_device_rng(seed)
    np.random.seed(seed)
    pb = buf.policy_buffers["policy_0"]
    for k in range(rounds):      # 45 x ~20 words > 624: crosses the twist at least once
        buf.sample(B)
        got = np.asarray(pb.sampled_indices(B))
        want = np.random.choice(E, B)
        assert np.array_equal(got, want), k
    # n == 1 consumes nothing; adopt / export of NumPy's own state round-trips
    one = make_buffers(N, O, A, S, T, 4, rng="device", max_batch=4)
    one.insert(1, *ep(1))
    np.random.seed(5)
    np.random.random(size=11)
    one.adopt_numpy_rng()
    one.sample(3)
    assert np.array_equal(np.asarray(one.policy_buffers["policy_0"].sampled_indices(3)), np.zeros(3, np.int64))
    one.policy_buffers["policy_0"].export_rng_to_numpy()
    a = np.random.random(size=4)
    np.random.seed(5)
    np.random.random(size=11)
    assert np.array_equal(a, np.random.random(size=4))


def check_per_golden():
    g = load_golden("replay_small")
    N, O, A, S, T, E = [int(v) for v in g["meta"]]
    buf = make_buffers(N, O, A, S, T, E, per_alpha=0.6)
    z = lambda *s: np.zeros(s, np.float32)
    dummy = [d(z(T + 1, 6, N, O)), d(z(T + 1, 6, N, S)), d(z(T, 6, N, A)), d(z(T, 6, N, 1)), d(z(T, 6, N, 1)), d(z(T, 6, 1)),
             d(z(T + 1, 6, N, A))]
    buf.insert(6, *dummy)
    buf.insert(6, *dummy)
    pb = buf.policy_buffers["policy_0"]
    # leaves handed over pre-powered in fp64 (NumPy's float32 pow is platform SIMD code; exactness of the
    # *tree arithmetic and indices* is what is pinned here; the on-device pow path is checked to 1 ulp below)
    leaves0 = g["per.prio0"] ** 0.6
    pb.update_priorities(np.arange(12), leaves=leaves0.astype(np.float64))
    s, m = pb.tree_values()
    assert np.array_equal(s, g["per.leaves0"]) and np.array_equal(m, g["per.minleaves0"])
    np.random.seed(77)
    smp = buf.sample(5, 0.4, "policy_0")
    assert np.array_equal(np.asarray(smp[8]), g["per.idx0"])
    assert np.allclose(np.asarray(smp[7]), g["per.w0"], rtol=1e-13, atol=0)
    # duplicate indices: last write wins
    pb.update_priorities(g["per.upd_idx"], leaves=(g["per.upd_prio"] ** 0.6).astype(np.float64))
    s, m = pb.tree_values()
    assert np.array_equal(s, g["per.leaves1"]) and np.array_equal(m, g["per.minleaves1"])
    smp = buf.sample(8, 0.7, "policy_0")
    assert np.array_equal(np.asarray(smp[8]), g["per.idx1"])
    assert np.allclose(np.asarray(smp[7]), g["per.w1"], rtol=1e-13, atol=0)
    # on-device fp32 pow path (what the trainer's priorities take): <= 1 ulp of NumPy's float32 power
    buf2 = make_buffers(N, O, A, S, T, E, per_alpha=0.6)
    buf2.insert(6, *dummy)
    buf2.insert(6, *dummy)
    buf2.update_priorities(np.arange(12), g["per.prio0"], "policy_0")
    s2, _ = buf2.policy_buffers["policy_0"].tree_values()
    cap = len(s2) // 2
    got = s2[cap:cap + 12].astype(np.float32)
    want = (g["per.prio0"] ** 0.6).astype(np.float32)
    assert np.all(np.abs(got.view(np.int32) - want.view(np.int32)) <= 1)


def check_per_vs_oracle_random(seed=5, E=50, B=8, steps=6):
    """Insert priming (intent of the reference's broken loop), sampling, write-back with duplicates, vs oracle."""
    N, O, A, S, T = 2, 4, 3, 5, 3
    rs = np.random.RandomState(seed)
    buf = make_buffers(N, O, A, S, T, E, per_alpha=0.6, rng="device", max_batch=16)
    ora = PrioritizedReplay(0.6, E, T, N, O, S, A, rng=LegacyMT19937(99), prime_leaves=True)
    pb = buf.policy_buffers["policy_0"]
    buf.seed_device_rng(99)

    def ep(n):
        f = [rs.randn(T + 1, n, N, O), np.repeat(rs.randn(T + 1, n, 1, S), N, 2), np.eye(A)[rs.randint(0, A, (T, n, N))],
             rs.randn(T, n, N, 1), np.zeros((T, n, N, 1)), np.zeros((T, n, 1)), np.ones((T + 1, n, N, A))]
        return [x.astype(np.float32) for x in f]

    for n in (16, 16, 9):
        e = ep(n)
        buf.insert(n, *[d(x) for x in e])
        ora.insert(n, *e)
    for k in range(steps):
        smp = buf.sample(B, 0.5, "policy_0")
        out, inds = ora.sample(B, 0.5)
        assert np.array_equal(np.asarray(smp[8]), inds), k
        assert np.allclose(np.asarray(smp[7]), out[7], rtol=1e-13)
        for name, ref in zip(REF_FIELDS, out[:7]):
            assert np.array_equal(smp[REF_FIELDS.index(name)]["policy_0"], ref), name
        pr = (rs.rand(B) * 2 + 0.01).astype(np.float32)
        leaves = (pr ** 0.6).astype(np.float64)
        pb.update_priorities(inds, leaves=leaves)
        ora.sum_tree.set(inds, leaves)
        ora.min_tree.set(inds, leaves)
        if k == 2:       # wrap the ring, which also re-primes overwritten slots with max_priority ** alpha
            e = ep(7)
            buf.insert(7, *[d(x) for x in e])
            ora.insert(7, *e)
        s, m = pb.tree_values()
        assert np.array_equal(s, ora.sum_tree.v) and np.array_equal(m, ora.min_tree.v), k


def check_uniform_vs_oracle_shapes(shape, seed=0):
    """Random insert/sample sequence at a given (N,O,A,S,T,E,B) vs oracle, including ring wrap and avail=None."""
    N, O, A, S, T, E, B, use_avail = shape
    rs = np.random.RandomState(seed)
    buf = make_buffers(N, O, A, S, T, E, rng="numpy", max_batch=max(B, 8), avail=use_avail)
    ora = UniformReplay(E, T, N, O, S, A, use_avail=use_avail, rng=None)

    def ep(n):
        f = [rs.randn(T + 1, n, N, O), np.repeat(rs.randn(T + 1, n, 1, S), N, 2), np.eye(A)[rs.randint(0, A, (T, n, N))],
             rs.randn(T, n, N, 1), (rs.rand(T, n, N, 1) < 0.3) * 1.0, (rs.rand(T, n, 1) < 0.3) * 1.0]
        f.append((rs.rand(T + 1, n, N, A) < 0.5) * 1.0 if use_avail else None)
        return [x.astype(np.float32) if x is not None else None for x in f]

    total = 0
    while total < 2 * E + 3:
        n = int(rs.randint(1, min(8, E + 1)))
        e = ep(n)
        r1 = buf.insert(n, *[d(x) for x in e])
        r2 = ora.insert(n, *[x for x in e])
        assert np.array_equal(r1, r2)
        total += n
        assert len(buf) == len(ora)
        assert buf.policy_buffers["policy_0"].current_i == ora.store.cursor
        st = np.random.get_state()
        smp = buf.sample(B)
        np.random.set_state(st)
        out, inds = ora.sample(B)
        for i, name in enumerate(REF_FIELDS):
            got = smp[i]["policy_0"]
            if out[i] is None:
                assert got is None
            else:
                assert np.array_equal(got, out[i]), name


def check_reward_norm_running_stats(seed=3, E=7, B=5, inserts=40):
    """Running reward statistics (updated at insert, evicted episodes subtracted) vs the oracle's full rescan
    (rec_buffer.py:209-223) over many ring wraps: single-episode inserts like the RNN runners, a few multi-episode ones, episodes
    that terminate early (masked steps), sampled rewards compared after every insert."""
    N, O, A, S, T = 3, 4, 3, 5, 6
    rs = np.random.RandomState(seed)
    buf = make_buffers(N, O, A, S, T, E, norm=True, rng="numpy", max_batch=max(B, 8))
    ora = UniformReplay(E, T, N, O, S, A, reward_norm=True, rng=None)

    def ep(n):
        de = np.maximum.accumulate((rs.rand(T, n, 1) < 0.15).astype(np.float32), axis=0)
        f = [rs.randn(T + 1, n, N, O), np.repeat(rs.randn(T + 1, n, 1, S), N, 2), np.eye(A)[rs.randint(0, A, (T, n, N))],
             np.repeat(3.0 + 2.0 * rs.randn(T, n, 1, 1), N, 2), np.repeat(de[:, :, None], N, 2), de, np.ones((T + 1, n, N, A))]
        return [x.astype(np.float32) for x in f]

    for k in range(inserts):
        n = 1 if k % 7 else int(rs.randint(2, 5))
        e = ep(n)
        buf.insert(n, *[d(x) for x in e])
        ora.insert(n, *e)
        st = np.random.get_state()
        smp = buf.sample(B)
        np.random.set_state(st)
        out, inds = ora.sample(B)
        got, want = smp[3]["policy_0"], out[3]
        assert np.abs(got - want).max() <= 2e-5 * max(1.0, np.abs(want).max()), (k, np.abs(got - want).max())
