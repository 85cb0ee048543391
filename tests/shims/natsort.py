"""test shim: natsort is not installed in this image"""
from himut_b200.natsort_compat import _key

_HIMUT_B200_SHIM = True


def natsorted(seq, key=None, reverse=False):
    if key is None:
        return sorted(seq, key=_key, reverse=reverse)
    return sorted(seq, key=lambda v: _key(key(v)), reverse=reverse)
