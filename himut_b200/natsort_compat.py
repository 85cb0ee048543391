"""natsorted() for the worker outputs when the natsort package is not importable.

The reference orders chrom2tsbs_lst with natsort.natsorted (src/himut/caller.py:622) and
contig names with the same call (src/himut/util.py:96).  natsort's default key (ns.INT)
splits strings into text / unsigned-integer runs, puts "" in front of a leading number so
text never meets a number, wraps bare numbers as ("", x) and recurses into tuples; this
module restates exactly that much.
"""
import re

_NUM = re.compile(r"(\d+)")


def _key(x):
    if isinstance(x, str):
        parts = [p for p in _NUM.split(x) if p != ""]
        out = []
        for p in parts:
            if p.isdigit():
                if not out or not isinstance(out[-1], str):
                    out.append("")
                out.append(int(p))
            else:
                out.append(p)
        return tuple(out)
    if isinstance(x, (tuple, list)):
        return tuple(_key(e) for e in x)
    return ("", x)


def natsorted(seq, key=None, reverse=False):
    try:  # the real thing when the environment has it
        import natsort  # noqa: WPS433
        if not getattr(natsort, "_HIMUT_B200_SHIM", False):
            return natsort.natsorted(seq, key=key, reverse=reverse)
    except ImportError:
        pass
    if key is None:
        return sorted(seq, key=_key, reverse=reverse)
    return sorted(seq, key=lambda v: _key(key(v)), reverse=reverse)
