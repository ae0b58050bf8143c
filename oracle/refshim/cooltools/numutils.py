"""``cooltools.numutils`` / ``cooltools.lib.numutils``: LazyToeplitz and zoom_array.

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


def zoom_array(in_array, final_shape, same_sum=False, zoom_function=None, **zoom_kwargs):
    """``cooltools.lib.numutils.zoom_array`` (used by ``_rescale_snip``, coolpup.py:1223-1233): rescale to
    ``final_shape`` with ``scipy.ndimage.zoom(order=1)``; when an axis shrinks, first zoom to the next multiple of the
    final size, then average blocks.  Restated from the cooltools source as remembered (scipy itself is the real one)."""
    from functools import partial

    from scipy.ndimage import zoom

    if zoom_function is None:
        zoom_function = partial(zoom, order=1)
    in_array = np.asarray(in_array, dtype=np.double)
    in_shape = in_array.shape
    assert len(in_shape) == len(final_shape)
    mults = []
    for i in range(len(in_shape)):
        if final_shape[i] < in_shape[i]:
            mults.append(int(np.ceil(in_shape[i] / final_shape[i])))
        else:
            mults.append(1)
    temp_shape = tuple([i * j for i, j in zip(final_shape, mults)])
    zoom_multipliers = np.array(temp_shape) / np.array(in_shape) + 0.0000001
    assert zoom_multipliers.min() >= 1
    rescaled = zoom_function(in_array, zoom_multipliers, **zoom_kwargs)
    for ind, mult in enumerate(mults):
        if mult != 1:
            sh = list(rescaled.shape)
            assert sh[ind] % mult == 0
            newshape = sh[:ind] + [sh[ind] // mult, mult] + sh[ind + 1 :]
            rescaled.shape = newshape
            rescaled = np.mean(rescaled, axis=ind + 1)
    assert rescaled.shape == tuple(final_shape)
    if same_sum:
        extra_size = np.prod(final_shape) / np.prod(in_shape)
        rescaled /= extra_size
    return rescaled
