"""Small seeded random inputs for kernel-level parity tests."""
import numpy as np


def random_region(nb, density_scale, seed, nan_frac=0.0, with_expected=False, with_cov=False, max_count=50):
    """Symmetric CSR with distance-decaying density, optional weights/expected/coverage."""
    rng = np.random.default_rng(seed)
    i, j = np.triu_indices(nb)
    d = j - i
    p = np.minimum(1.0, density_scale / np.maximum(d, 1))
    keep = rng.random(i.shape[0]) < p
    i, j = i[keep], j[keep]
    c = rng.integers(1, max_count, size=i.shape[0]).astype(np.int32)
    off = i != j
    row = np.concatenate([i, j[off]])
    col = np.concatenate([j, i[off]])
    val = np.concatenate([c, c[off]])
    order = np.lexsort((col, row))
    row, col, val = row[order], col[order], val[order]
    indptr = np.zeros(nb + 1, dtype=np.int64)
    np.cumsum(np.bincount(row, minlength=nb), out=indptr[1:])
    weight = expected = cov = None
    if nan_frac is not None and nan_frac >= 0:
        weight = np.exp(rng.normal(0, 0.2, nb)) * 1e-2
        weight[rng.random(nb) < nan_frac] = np.nan
    if with_expected:
        expected = 5.0 / np.maximum(np.arange(nb), 1.0) * (1 + 0.1 * rng.random(nb))
        expected[:2] = np.nan
        if nb > 40:
            expected[nb // 2] = np.nan
            expected[nb // 3] = 0.0
    if with_cov:
        cov = rng.integers(0, 1000, nb).astype(np.float64)
    return indptr.astype(np.int32), col.astype(np.int32), val.astype(np.int32), weight, expected, cov


def random_windows(nb, W, n, n_slots, seed, near_diag_frac=0.3, oob_frac=0.05):
    rng = np.random.default_rng(seed)
    r0 = rng.integers(0, max(1, nb - W), n)
    c0 = rng.integers(0, max(1, nb - W), n)
    near = rng.random(n) < near_diag_frac
    c0[near] = np.clip(r0[near] + rng.integers(-W, 2 * W, near.sum()), 0, max(0, nb - W))
    oob = rng.random(n) < oob_frac
    r0[oob] += rng.choice([-nb, nb, -3, 3 + nb - W], oob.sum())
    slot = rng.integers(0, n_slots, n)
    return r0.astype(np.int32), c0.astype(np.int32), slot.astype(np.int32)
