"""reflib.get_chrom_tricount: oracle vs the reference's own output (CPU),
CUDA vs oracle and vs the reference (GPU)."""
import importlib.util
import json
import os

import numpy as np
import pytest

import cases
import parity
from oracle import oracle

_spec = importlib.util.spec_from_file_location("mkref", os.path.join(cases.GOLDEN_DIR, "make_golden_reflib.py"))


def _sequences():
    import random
    rnd = random.Random(5)
    return {"plain": "".join(rnd.choice("ACGT") for _ in range(20011)),
            "dirty": "".join(rnd.choice("ACGTNacgtnRY") if rnd.random() < 0.1 else rnd.choice("ACGT") for _ in range(9973)),
            "tiny": "ACG", "two": "AC", "empty": "", "nrun": "NNNNACGTNNACGTTTNN"}


def _expected():
    return json.load(open(os.path.join(cases.GOLDEN_DIR, "reflib.json")))["expected"]


def test_oracle_matches_reference():
    exp = _expected()
    for name, seq in _sequences().items():
        got = oracle.ref_tricounts(seq.encode())
        assert np.array_equal(got, parity.tri_dict_to_bins(exp[name])), name


@pytest.mark.gpu
def test_cuda_matches_oracle_and_reference(ctx):
    exp = _expected()
    for name, seq in _sequences().items():
        got = ctx.ref_tricounts(seq.encode())
        assert np.array_equal(got, oracle.ref_tricounts(seq.encode())), name
        assert np.array_equal(got, parity.tri_dict_to_bins(exp[name])), name


@pytest.mark.gpu
def test_cuda_64mb_contig(ctx):
    from himut_b200 import synth
    d = synth.generate(64_000_000, seed=20260101, depth=0.01, copy=False)
    ref = bytearray(d.ref)
    ref[1000:1100] = b"N" * 100
    ref[5_000_000:5_000_050] = bytes(ref[5_000_000:5_000_050]).lower()
    ref = bytes(ref)
    got = ctx.ref_tricounts(ref)
    assert np.array_equal(got, oracle.ref_tricounts(ref))
    assert int(got.sum()) == len(ref) - 2 - 100  # every window whose first base is not N


@pytest.mark.gpu
def test_worker_mirror(ctx):
    from himut_b200 import reflib
    exp = _expected()
    out = {}
    reflib.get_chrom_tricount("dirty", _sequences()["dirty"], out)
    for tri in parity.TRI_LST:
        assert out["dirty"][tri] == exp["dirty"].get(tri, 0)
