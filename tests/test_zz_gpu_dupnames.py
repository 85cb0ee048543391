"""Duplicate query names at a phase-checked site: three long-read `--phase` seeds on which the reference —
and the oracle, which follows it (tests/test_oracle_random.py) — classify the re-fetched records by query name
(src/himut/caller.py:556-567) while the kernels classify by record.  Expected to fail until k_site_reduce tests
qname membership; kept as the regression cases of that fix.  GPU."""
import pytest

import cases
import parity
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.xfail(reason="known deviation: records classified by record, not by query name (DESIGN.md §7)", strict=False)
@pytest.mark.parametrize("seed", cases.RANDOM_DUPNAME_CALL_SEEDS)
def test_duplicate_names_at_a_phase_checked_site(ctx, seed):
    c = cases.random_case("call", seed)
    ctx.set_params(c["params"])
    ctx.set_site_sets(c["common"], c["pon"])
    ctx.set_phase_sets(c["phase"])
    rec, log = ctx.call_batch(c["batch"], c["chunk_table"])
    o_rec, o_log = oracle.call_chunks(c["params"], c["batch"], c["chunk_table"], c["common"], c["pon"], c["phase"])
    ok, why = parity.records_equal(rec, o_rec)
    assert ok, why
    assert list(log) == list(o_log)
