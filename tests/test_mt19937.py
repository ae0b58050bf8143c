"""The oracle's restatement of numpy's legacy MT19937 randint / choice stream (what the device-side control-shift
generator mirrors) against the installed numpy."""
import numpy as np
import pytest

from oracle.mt19937_ref import MT19937


@pytest.mark.parametrize("seed", [0, 1, 12345])
def test_stream_matches_numpy(seed):
    np.random.seed(seed)
    np.random.random(7)  # arbitrary position inside a state block
    mt = MT19937.from_numpy()
    for n, lo, hi in [(5, 100_000, 1_000_000), (700, 100_000, 1_000_000), (33, 10, 11_000), (1, 5, 7), (64, 0, 2**20 + 3)]:
        a = np.random.randint(lo, hi, n)
        s = np.random.choice([-1, 1], n)
        assert np.array_equal(mt.randint(lo, hi, n), a)
        assert np.array_equal(mt.choice_sign(n), s)
    # the state the replay ends in is numpy's state
    st = np.random.get_state()
    if mt.pos >= 624:
        mt._gen()
    if st[2] >= 624:
        np.random.randint(0, 2, 1)
        mt.next_uint32()
        st = np.random.get_state()
    assert mt.pos == st[2] and np.array_equal(np.asarray(mt.key, dtype=np.uint32), st[1])


def test_control_shifts_match_product_host_path():
    from coolpuppy_b200._coords import _draw_shifts

    np.random.seed(3)
    mt = MT19937.from_numpy()
    want = _draw_shifts(500, 100_000, 1_000_000, 10_000)
    assert np.array_equal(mt.control_shifts(100_000, 1_000_000, 10_000, 500), want)
    mt.to_numpy()  # nothing to advance: numpy already consumed the same words
