"""``cooltools.numutils`` / ``cooltools.lib.numutils``: LazyToeplitz only.

Documented behaviour: ``LazyToeplitz(c, r)[i0:i1, j0:j1]`` materialises the
block of the Toeplitz matrix whose first column is ``c`` and first row ``r``
(``r = c`` when omitted): ``T[i, j] = r[j - i]`` for ``j >= i`` else
``c[i - j]``.  Slices are clipped to the matrix shape like numpy slicing.
"""
import numpy as np


class LazyToeplitz:
    def __init__(self, c, r=None):
        self._c = np.asarray(c)
        self._r = self._c if r is None else np.asarray(r)

    @property
    def shape(self):
        return (len(self._c), len(self._r))

    def __getitem__(self, key):
        s0, s1 = key
        i0, i1, st0 = s0.indices(self.shape[0])
        j0, j1, st1 = s1.indices(self.shape[1])
        assert st0 == 1 and st1 == 1
        i = np.arange(i0, max(i0, i1))[:, None]
        j = np.arange(j0, max(j0, j1))[None, :]
        d = j - i
        upper = self._r[np.clip(d, 0, None)]
        lower = self._c[np.clip(-d, 0, None)]
        return np.where(d >= 0, upper, lower)


def zoom_array(*args, **kwargs):
    raise NotImplementedError("zoom_array is only used by rescaled pile-ups (out of scope)")
