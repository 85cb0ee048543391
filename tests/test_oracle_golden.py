"""The CPU oracle against outputs of the unmodified reference (tests/golden/*.json, generated
by tests/golden/make_golden.py in the build container).  CPU only."""
import numpy as np
import pytest

import cases
import parity
from himut_b200 import records
from oracle import oracle

CALL_CASES = [n for n in cases.CASES if n.startswith("call_")]
NORM_CASES = [n for n in cases.CASES if n.startswith("norm_")]


@pytest.mark.parametrize("name", CALL_CASES)
def test_call_matches_reference(name):
    fx = parity.load_golden(name)
    c = cases.build_case(name)
    assert fx["batch_sha256"] == cases.batch_digest(c["batch"]), "synthetic batch differs from the fixture's"
    rec, log = oracle.call_chunks(c["params"], c["batch"], c["chunk_table"], c["common"], c["pon"], c["phase"])
    rows = records.records_to_tsbs_lst(cases.CHROM, rec)
    gold = parity.golden_rows(fx)
    assert parity.rows_equal(rows, gold), parity.first_diff(rows, gold, rec)
    assert [int(v) for v in log] == fx["expected"]["log"]


@pytest.mark.parametrize("name", NORM_CASES)
def test_normcounts_matches_reference(name):
    fx = parity.load_golden(name)
    c = cases.build_case(name)
    assert fx["batch_sha256"] == cases.batch_digest(c["batch"])
    e = fx["expected"]
    ccs, rt, log, _ = oracle.normcounts_chunks(c["params"], c["batch"], c["ref"].encode(), c["chunk_table"],
                                               c["common"], c["pon"], c["phase"],
                                               alt_order=np.array(e["alt_order"], np.uint8))
    assert np.array_equal(ccs, parity.tri_dict_to_bins(e["ccs_tri2count"]))
    assert np.array_equal(rt, parity.tri_dict_to_bins(e["ref_tri2count"]))
    assert [int(v) for v in log] == e["log"]


def test_every_status_is_covered():
    """the fixtures together exercise every branch of the cascade the data can reach"""
    seen = set()
    for name in CALL_CASES:
        for row in parity.golden_rows(parity.load_golden(name)):
            seen.add(row[4])
    for s in ("PASS", "HetSite", "HomAltSite", "IndelSite", "LowGQ", "LowBQ", "PanelOfNormal", "ComSnp",
              "LowDepth", "HighDepth", "Unphased", "HetAltSite"):
        assert s in seen, s
