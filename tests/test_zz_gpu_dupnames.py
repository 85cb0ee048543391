"""Duplicate query names at a phase-checked site: three long-read `--phase` seeds on which classifying the re-fetched
records by record instead of by query name (src/himut/caller.py:556-567) changes hap_count / PASS-vs-Unphased.
k_site_reduce walks the re-fetched records and tests name membership when the batch holds shared names; these seeds are
the regression cases (the oracle follows the reference: tests/test_oracle_random.py).  Both device paths.  GPU."""
import pytest

import cases
import parity
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("v1", [False, True], ids=["fused", "first_version"])
@pytest.mark.parametrize("seed", cases.RANDOM_DUPNAME_CALL_SEEDS)
def test_duplicate_names_at_a_phase_checked_site(seed, v1, monkeypatch):
    import himut_b200.lib as lib
    if v1:
        monkeypatch.setenv("HIMUT_B200_CALL_V1", "1")
    with lib.Context(0) as ctx:
        _run(ctx, seed, 1 if v1 else 2)


def _run(ctx, seed, path):
    c = cases.random_case("call", seed)
    ctx.set_params(c["params"])
    ctx.set_site_sets(c["common"], c["pon"])
    ctx.set_phase_sets(c["phase"])
    rec, log = ctx.call_batch(c["batch"], c["chunk_table"])
    o_rec, o_log = oracle.call_chunks(c["params"], c["batch"], c["chunk_table"], c["common"], c["pon"], c["phase"])
    ok, why = parity.records_equal(rec, o_rec)
    assert ok, why
    assert list(log) == list(o_log)
    assert ctx.last_call_path() == path
