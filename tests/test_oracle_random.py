"""The CPU oracle against the unmodified reference on the randomised sweep (tests/golden/random_sweep.json, written by
tests/golden/make_golden_random.py in the build container): the same 60 dirty batches x random parameters that
tests/test_gpu_random.py runs through the CUDA path, so reference == oracle == CUDA on every one of them — plus a
second family of dirty reads several 2048-position tiles long (tests/test_zz_gpu_random_long.py on the GPU).  CPU only."""
import numpy as np
import pytest

import cases
import parity
from himut_b200 import records
from oracle import oracle

SWEEP = parity.load_random_sweep()


@pytest.mark.parametrize("seed", list(cases.RANDOM_CALL_SEEDS) + list(cases.RANDOM_LONG_CALL_SEEDS) + list(cases.RANDOM_DUPNAME_CALL_SEEDS))
def test_call_matches_reference(seed):
    c = cases.random_case("call", seed)
    if c is None:
        assert str(seed) not in SWEEP["call"]
        pytest.skip("no phase set")
    fx = SWEEP["call"][str(seed)]
    assert fx["batch_sha256"] == cases.batch_digest(c["batch"]), "random batch differs from the fixture's"
    rec, log = oracle.call_chunks(c["params"], c["batch"], c["chunk_table"], c["common"], c["pon"], c["phase"])
    rows = records.records_to_tsbs_lst(cases.CHROM, rec)
    assert len(rows) == fx["n_rows"]
    assert parity.rows_digest(rows) == fx["rows_sha256"]
    assert [int(v) for v in log] == fx["log"]


@pytest.mark.parametrize("seed", list(cases.RANDOM_NORM_SEEDS) + list(cases.RANDOM_LONG_NORM_SEEDS))
def test_normcounts_matches_reference(seed):
    c = cases.random_case("norm", seed)
    if c is None:
        assert str(seed) not in SWEEP["norm"]
        pytest.skip("no phase set")
    fx = SWEEP["norm"][str(seed)]
    assert fx["batch_sha256"] == cases.batch_digest(c["batch"])
    ccs, rt, log, _ = oracle.normcounts_chunks(c["params"], c["batch"], c["ref"].encode(), c["chunk_table"], c["common"],
                                               c["pon"], c["phase"], alt_order=np.array(fx["alt_order"], np.uint8))
    assert np.array_equal(ccs, parity.tri_dict_to_bins(fx["ccs_tri2count"]))
    assert np.array_equal(rt, parity.tri_dict_to_bins(fx["ref_tri2count"]))
    assert [int(v) for v in log] == fx["log"]


def test_sweep_is_not_trivial():
    """the sweep reaches every status of the cascade and both kinds of phase outcome"""
    seen = set()
    for fx in SWEEP["call"].values():
        seen.update(fx["statuses"])
    for s in ("PASS", "HetSite", "HetAltSite", "HomAltSite", "IndelSite", "LowGQ", "LowBQ", "PanelOfNormal", "ComSnp",
              "LowDepth", "HighDepth", "Unphased"):
        assert s in seen, s
    assert sum(1 for fx in SWEEP["norm"].values() if fx["log"][-1] > 0) >= 15
