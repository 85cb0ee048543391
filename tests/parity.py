"""Comparison helpers shared by the oracle-vs-golden and CUDA-vs-oracle parity tests."""
import json
import os

import numpy as np

from himut_b200 import abi, records

import cases

# mutlib.tri_lst order (src/himut/mutlib.py:17-50): first base A,C,G,T x centre C,T x last A,C,G,T
TRI_LST = [a + c + b for a in "ACGT" for c in "CT" for b in "ACGT"]


def load_golden(name):
    path = os.path.join(cases.GOLDEN_DIR, name + ".json")
    with open(path) as f:
        return json.load(f)


def golden_rows(fx):
    return [tuple(r) for r in fx["expected"]["tsbs_lst"]]


def rows_equal(a, b):
    """tuple lists equal with exact float equality (np.float64 == float compares exactly)"""
    if len(a) != len(b):
        return False
    return all(len(x) == len(y) and all(u == v for u, v in zip(x, y)) for x, y in zip(a, b))


def rows_digest(rows):
    """SHA-256 over the repr of every field of every row (repr of a float round-trips exactly; numpy scalars are
    taken as the Python numbers they equal)"""
    import hashlib
    h = hashlib.sha256()
    for row in rows:
        plain = [float(v) if isinstance(v, (np.floating, float)) else int(v) if isinstance(v, (np.integer, int)) else v for v in row]
        h.update(("\t".join(repr(v) for v in plain) + "\n").encode())
    return h.hexdigest()


def load_random_sweep():
    """tests/golden/random_sweep.json: what the reference returned for every seed of cases.random_case"""
    with open(os.path.join(cases.GOLDEN_DIR, "random_sweep.json")) as f:
        return json.load(f)


def first_diff(a, b, rec=None):
    """where two tuple lists part; rec: the record array the first came from (its PL-tie count goes into the message)"""
    note = tie_note(rec) if rec is not None else ""
    for i, (x, y) in enumerate(zip(a, b)):
        if not (len(x) == len(y) and all(u == v for u, v in zip(x, y))):
            return "row %d: %r != %r" % (i, x, y) + note
    return "lengths %d vs %d; extra: %r" % (len(a), len(b), (a[len(b):] or b[len(a):])[:3]) + note


def tri_dict_to_bins(d):
    """reference tri2count dict -> 33 bins (32 tri_lst + everything else)"""
    out = np.zeros(abi.TRI_BINS, np.int64)
    for k, v in d.items():
        out[TRI_LST.index(k) if k in TRI_LST else 32] += int(v)
    return out


def sort_records(rec):
    """canonical order for comparing two record arrays"""
    order = np.lexsort((rec["alt"], rec["ref"], rec["tpos"], rec["chunk"]))
    return rec[order]


def tie_note(*recs):
    """how many records of each array carry HM_SITE_PL_TIE (the PL minimum is shared by two genotypes: the reference's
    choice there rests on numpy's argsort being stable) — part of every parity message"""
    return " [PL-tie flagged records: %s]" % ", ".join("%d of %d" % (int(((r["flags"] & abi.SITE_PL_TIE) != 0).sum()), r.size) for r in recs)


def records_equal(a, b):
    a, b = sort_records(a), sort_records(b)
    if a.shape != b.shape:
        return False, "record counts %d vs %d" % (a.size, b.size) + tie_note(a, b)
    for name in abi.SITE_DTYPE.names:
        if name == "pad0":
            continue
        if not np.array_equal(a[name], b[name]):
            bad = np.flatnonzero((a[name] != b[name]).reshape(a.size, -1).any(axis=1))[:3]
            return False, "field %s differs at %s: %r vs %r" % (name, bad, a[bad], b[bad]) + tie_note(a, b)
    return True, ""
