"""BAM pre-pass: himut_b200.bamlib.get_thresholds against the reference's own
bamlib.get_thresholds output (tests/golden/thresholds.json), through real BAM files and the native decoder."""
import importlib.util
import json
import os

import numpy as np
import pytest

import cases
from himut_b200 import bamdec, bamio, bamlib

_spec = importlib.util.spec_from_file_location("mkthr", os.path.join(cases.GOLDEN_DIR, "make_golden_thresholds.py"))
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)


@pytest.mark.parametrize("name", sorted(mk.CASES))
def test_thresholds_match_the_reference(tmp_path, name):
    exp = json.load(open(os.path.join(cases.GOLDEN_DIR, "thresholds.json")))["expected"][name]
    items = mk.batches(name)
    path = str(tmp_path / "t.bam")
    bamio.write_batches_bam(path, items)
    got = bamlib.get_thresholds(path, [c for c, _, _ in items], {c: n for c, n, _ in items})
    assert [int(v) for v in got] == exp


def test_window_filter_is_mapq_and_tp_only(tmp_path):
    """mapq > 0 and tp:A:P decide; the secondary / supplementary flag does not (bamlib.py:160-164)"""
    path = str(tmp_path / "f.bam")
    w = bamio.BamWriter(path, [("chr1", 10_000)])
    def rec(pos, name, flag=0, mapq=60, tp="P", n=50):
        tags = [("cs", "Z", ":%d" % n)] + ([("tp", "A", tp)] if tp else [])
        w.add(0, pos, name, flag, mapq, [(0, n)], "A" * n, bytes([30] * n), tags)
    rec(100, "a", n=50); rec(110, "mapq0", mapq=0, n=51); rec(120, "secondary_tp_P", flag=0x100, n=52)
    rec(130, "tp_S", tp="S", n=53); rec(140, "no_tp", tp=None, n=54); rec(150, "supp", flag=0x800, n=55)
    rec(5000, "far", n=56)
    w.close()
    nb = bamdec.NativeBam(path, threads=1)
    assert nb.window_qlens("chr1", 0, 1000).tolist() == [50, 52, 55]
    assert nb.window_qlens("chr1", 149, 150).tolist() == [50, 52]   # [100, 150) and [120, 172) overlap, [150, 205) does not
    assert nb.window_qlens("chr1", 4000, 20_000).tolist() == [56]
    assert nb.window_qlens("chr1", 9000, 9500).tolist() == []
