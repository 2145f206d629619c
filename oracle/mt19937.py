"""NumPy *legacy* global-stream generator, restated (oracle; test infrastructure only).

The reference draws replay indices with `np.random.choice(len, B)`
(/root/reference/offpolicy/utils/rec_buffer.py:76) and PER masses with
`np.random.random(size=B)` (rec_buffer.py:274) from the process-global legacy
`RandomState`, seeded by `np.random.seed(seed)` (scripts/train/train_smac.py:121).
That algorithm lives in NumPy (third-party, `numpy==1.18.5`, requirements.txt:77), not in
/root/reference; its published algorithm (Matsumoto & Nishimura MT19937; NumPy
`randomkit`/`_legacy` distributions) is restated here and pinned against the installed
NumPy in tests/test_oracle_rng.py.

 * seed(s):      key[0]=s; key[i] = 1812433253*(key[i-1]^(key[i-1]>>30)) + i   (mod 2^32), pos=624
 * next32():     standard twist every 624 words + tempering
 * randint(0,n): masked rejection on 32-bit words: mask = 2^ceil(log2(n))-1; redraw while (w&mask) > n-1;
                 n == 1 consumes nothing
 * random():     ((w1>>5)*2^26 + (w2>>6)) / 2^53 from two consecutive words
"""
import numpy as np

N, M = 624, 397
_U32 = 0xFFFFFFFF


class LegacyMT19937(object):
    def __init__(self, seed=None):
        self.key = np.zeros(N, dtype=np.uint32)
        self.pos = N
        if seed is not None:
            self.seed(seed)

    # -- state -----------------------------------------------------------------
    def seed(self, s):
        s = int(s) & _U32
        key = [0] * N
        for i in range(N):
            key[i] = s
            s = (1812433253 * (s ^ (s >> 30)) + i + 1) & _U32
        self.key = np.array(key, dtype=np.uint32)
        self.pos = N

    def set_state(self, key, pos):
        self.key = np.array(key, dtype=np.uint32).copy()
        self.pos = int(pos)

    def get_state(self):
        return self.key.copy(), self.pos

    @classmethod
    def from_numpy_global(cls):
        st = np.random.get_state()
        g = cls()
        g.set_state(st[1], st[2])
        return g

    # -- core ------------------------------------------------------------------
    def _twist(self):
        k = [int(v) for v in self.key]
        for i in range(N):
            y = (k[i] & 0x80000000) | (k[(i + 1) % N] & 0x7FFFFFFF)
            v = k[(i + M) % N] ^ (y >> 1)
            if y & 1:
                v ^= 0x9908B0DF
            k[i] = v
        self.key = np.array(k, dtype=np.uint32)
        self.pos = 0

    def next32(self):
        if self.pos >= N:
            self._twist()
        y = int(self.key[self.pos])
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & _U32

    # -- distributions -----------------------------------------------------------
    def randint0(self, n, size):
        """np.random.randint(0, n, size) == np.random.choice(n, size) (legacy, int64)."""
        out = np.zeros(size, dtype=np.int64)
        rng = int(n) - 1
        if rng == 0:
            return out
        mask = rng
        for sh in (1, 2, 4, 8, 16):
            mask |= mask >> sh
        for i in range(size):
            while True:
                v = self.next32() & mask
                if v <= rng:
                    break
            out[i] = v
        return out

    choice = randint0

    def random(self, size):
        out = np.empty(size, dtype=np.float64)
        for i in range(size):
            a = self.next32() >> 5
            b = self.next32() >> 6
            out[i] = (a * 67108864.0 + b) / 9007199254740992.0
        return out
