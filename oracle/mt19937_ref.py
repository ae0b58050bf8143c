"""TEST INFRASTRUCTURE ONLY -- plain-Python restatement of the legacy ``np.random`` calls the reference makes for
its control shifts (``coolpup.py:392-396, 442-445``): ``np.random.randint(minshift, maxshift, n)`` followed by
``np.random.choice([-1, 1], n)`` on the global MT19937 ``RandomState``.

The arithmetic lives in numpy (third-party, not under /root/reference): ``numpy/random/src/mt19937/mt19937.c``
(state update + tempering) and ``numpy/random/src/distributions/distributions.c`` (``random_bounded_uint64_fill`` ->
``buffered_bounded_masked_uint32``: draw 32 bits, mask to the smallest 2^k - 1 >= range, reject values above the
range); ``RandomState.choice(a, n)`` with a uniform ``p`` is ``randint(0, len(a), n)`` then ``a[idx]``.  Pinned
against the installed numpy by ``tests/test_mt19937.py``; the CUDA kernel ``k_mt_shifts`` mirrors this file.
"""
import numpy as np

N, M = 624, 397
MATRIX_A, UPPER, LOWER = 0x9908B0DF, 0x80000000, 0x7FFFFFFF


class MT19937:
    def __init__(self, key, pos):
        self.key = [int(x) for x in key]
        self.pos = int(pos)

    @classmethod
    def from_numpy(cls):
        st = np.random.get_state()
        return cls(st[1], st[2])

    def to_numpy(self):
        st = np.random.get_state()
        np.random.set_state((st[0], np.asarray(self.key, dtype=np.uint32), self.pos, st[3], st[4]))

    def _gen(self):
        k = self.key
        for i in range(N):
            y = (k[i] & UPPER) | (k[(i + 1) % N] & LOWER)
            k[i] = k[(i + M) % N] ^ (y >> 1) ^ (MATRIX_A if (y & 1) else 0)
        self.pos = 0

    def next_uint32(self):
        if self.pos >= N:
            self._gen()
        y = self.key[self.pos]
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def randint(self, low, high, n):
        rng = high - 1 - low
        assert 0 < rng < 0xFFFFFFFF
        mask = rng
        for s in (1, 2, 4, 8, 16):
            mask |= mask >> s
        out = np.empty(n, dtype=np.int64)
        for i in range(n):
            while True:
                v = self.next_uint32() & mask
                if v <= rng:
                    break
            out[i] = low + v
        return out

    def choice_sign(self, n):
        """np.random.choice([-1, 1], n)"""
        return np.array([(-1, 1)[self.next_uint32() & 1] for _ in range(n)], dtype=np.int64)

    def control_shifts(self, minshift, maxshift, resolution, n):
        """One ``_control_regions`` draw: bins to add to the window coordinates."""
        shift = self.randint(minshift, maxshift, n) * self.choice_sign(n)
        return np.round(shift / resolution).astype(np.int64)
