"""Pin oracle.mt19937 against the installed NumPy legacy stream."""
import numpy as np
import pytest

from oracle.mt19937 import LegacyMT19937


@pytest.mark.parametrize("seed", [0, 1, 123, 2**32 - 1])
@pytest.mark.parametrize("n", [5000, 37, 4096, 1, 2, 4097])
def test_choice_matches_numpy(seed, n):
    np.random.seed(seed)
    want = np.concatenate([np.random.choice(n, 32), np.random.choice(n, 700)])
    g = LegacyMT19937(seed)
    got = np.concatenate([g.choice(n, 32), g.choice(n, 700)])
    assert got.dtype == np.int64 and np.array_equal(got, want)


def test_random_matches_numpy_and_interleaves():
    np.random.seed(42)
    a = np.random.choice(100, 10)
    r = np.random.random(size=400)
    b = np.random.choice(7, 5)
    g = LegacyMT19937(42)
    assert np.array_equal(g.choice(100, 10), a)
    assert np.array_equal(g.random(400), r)      # bit-exact doubles
    assert np.array_equal(g.choice(7, 5), b)


def test_adopt_numpy_global_state():
    np.random.seed(9)
    np.random.random(size=17)
    g = LegacyMT19937.from_numpy_global()
    assert np.array_equal(g.choice(5000, 64), np.random.choice(5000, 64))
