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


@pytest.mark.parametrize("n", [1, 2, 37, 64, 4096, 5000, 100000])
def test_randint_is_the_same_draw_as_choice(n):
    """The drop-in buffer draws `np.random.randint(0, len, B)` where the reference writes `np.random.choice(len, B)`
    (rec_buffer.py:76): same values, same dtype, same generator state afterwards."""
    np.random.seed(11)
    a = np.random.choice(n, 32)
    sa = np.random.get_state()
    np.random.seed(11)
    b = np.random.randint(0, n, 32)
    sb = np.random.get_state()
    assert a.dtype == b.dtype and np.array_equal(a, b)
    assert sa[2] == sb[2] and np.array_equal(sa[1], sb[1])
