"""Stand-in for ``natsort.natsorted``: strings compared with digit runs as integers."""
import re

_tok = re.compile(r"(\d+)")


def _key(s):
    parts = _tok.split(str(s))
    return tuple((0, int(p)) if p.isdigit() else (1, p) for p in parts if p != "")


def natsorted(seq, key=None):
    if key is None:
        return sorted(seq, key=_key)
    return sorted(seq, key=lambda x: _key(key(x)))
