"""Pin oracle.replay against the reference buffers' outputs (tests/golden/replay_small.npz)."""
import numpy as np

from helpers import load_golden, FIELDS
from oracle.mt19937 import LegacyMT19937
from oracle.replay import UniformReplay, PrioritizedReplay


def _ins(g, tag, j):
    return tuple(g["%s.ins%d.%s" % (tag, j, k)] for k in FIELDS)


def test_uniform_insert_sample_bit_exact():
    g = load_golden("replay_small")
    N, O, A, S, T, E = [int(v) for v in g["meta"]]
    for tag, norm in (("plain", False), ("norm", True)):
        buf = UniformReplay(E, T, N, O, S, A, reward_norm=norm, rng=LegacyMT19937(123))
        for j, n_ep in enumerate((5, 7, 9)):
            r = buf.insert(n_ep, *_ins(g, tag, j))
            assert np.array_equal(r, g["%s.idx_range%d" % (tag, j + 1)])
        assert len(buf) == E
        for d in range(3):
            out, inds = buf.sample(6)
            assert np.array_equal(inds, g["%s.inds" % tag][d])
            for k, f in zip(FIELDS, out[:7]):
                want = g["%s.draw%d.%s" % (tag, d, k)]
                if k == "rew" and norm:
                    assert np.allclose(f, want, rtol=1e-6, atol=1e-6)
                else:
                    assert np.array_equal(f, want), (tag, d, k)


def test_per_sample_weights_and_writeback():
    g = load_golden("replay_small")
    N, O, A, S, T, E = [int(v) for v in g["meta"]]
    per = PrioritizedReplay(0.6, E, T, N, O, S, A, rng=LegacyMT19937(77), prime_leaves=False)
    rs = np.zeros
    dummy = (rs((T + 1, 6, N, O), np.float32), rs((T + 1, 6, N, S), np.float32), rs((T, 6, N, A), np.float32),
             rs((T, 6, N, 1), np.float32), rs((T, 6, N, 1), np.float32), rs((T, 6, 1), np.float32),
             rs((T + 1, 6, N, A), np.float32))
    per.insert(6, *dummy)
    per.insert(6, *dummy)
    per.update_priorities(np.arange(12), g["per.prio0"])
    assert np.array_equal(per.sum_tree.v, g["per.leaves0"])
    assert np.array_equal(per.min_tree.v, g["per.minleaves0"])
    out, inds = per.sample(5, 0.4)
    assert np.array_equal(inds, g["per.idx0"])
    assert np.array_equal(out[7], g["per.w0"])
    per.update_priorities(g["per.upd_idx"], g["per.upd_prio"])
    assert np.array_equal(per.sum_tree.v, g["per.leaves1"])
    assert np.array_equal(per.min_tree.v, g["per.minleaves1"])
    assert per.max_priority == float(g["per.maxprio1"])
    out, inds = per.sample(8, 0.7)
    assert np.array_equal(inds, g["per.idx1"])
    assert np.array_equal(out[7], g["per.w1"])
