"""Randomised parity sweep, second family: dirty reads of up to 9 kb (several 2048-position normcounts tiles, reads
crossing many chunk edges) on 20-30 kb contigs.  Same checks as tests/test_gpu_random.py; a file of its own, late in
the run, because this family was added after the round's last GPU session.

Randomised parity sweep: many small dirty batches (dense substitutions / indels, N reference
bases, soft clips, secondary records, duplicate names, random BQ), random worker parameters,
random overlapping chunk lists, random phase tables and site sets — CUDA against the CPU oracle,
every record field and every counter — and against what the unmodified reference returned for the same
seeds (tests/golden/random_sweep.json: rows, log vectors, tri-count tables).  GPU."""
import numpy as np
import pytest

import cases
import parity
from himut_b200 import records
from oracle import oracle

pytestmark = pytest.mark.gpu

SWEEP = parity.load_random_sweep()


@pytest.mark.parametrize("seed", cases.RANDOM_LONG_CALL_SEEDS)
def test_random_long_call(ctx, seed):
    c = cases.random_case("call", seed)
    if c is None:
        pytest.skip("no phase set")
    batch, p, table, common, pon, phase = c["batch"], c["params"], c["chunk_table"], c["common"], c["pon"], c["phase"]
    ctx.set_params(p)
    ctx.set_site_sets(common, pon)
    if phase is not None:
        ctx.set_phase_sets(phase)
    rec, log = ctx.call_batch(batch, table)
    o_rec, o_log = oracle.call_chunks(p, batch, table, common, pon, phase)
    ok, why = parity.records_equal(rec, o_rec)
    assert ok, why
    assert list(log) == list(o_log)
    fx = SWEEP["call"][str(seed)]  # the reference itself
    assert fx["batch_sha256"] == cases.batch_digest(batch)
    assert parity.rows_digest(records.records_to_tsbs_lst(cases.CHROM, rec)) == fx["rows_sha256"]
    assert [int(v) for v in log] == fx["log"]


@pytest.mark.parametrize("seed", cases.RANDOM_LONG_NORM_SEEDS)
def test_random_long_normcounts(ctx, seed):
    c = cases.random_case("norm", seed)
    if c is None:
        pytest.skip("no phase set")
    batch, ref, p, table, common, pon, phase = c["batch"], c["ref"], c["params"], c["chunk_table"], c["common"], c["pon"], c["phase"]
    ctx.set_params(p)
    ctx.set_site_sets(common, pon)
    if phase is not None:
        ctx.set_phase_sets(phase)
    ctx.upload(batch)
    g = ctx.normcounts_chunks(ref.encode(), table)
    o = oracle.normcounts_chunks(p, batch, ref.encode(), table, common, pon, phase)
    assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1])
    assert list(g[2]) == list(o[2])
    assert g[3] == o[3]
    fx = SWEEP["norm"][str(seed)]  # the reference itself (where no alt tie was flagged: its set order decides those)
    if g[3] == 0:
        assert np.array_equal(g[0], parity.tri_dict_to_bins(fx["ccs_tri2count"]))
        assert np.array_equal(g[1], parity.tri_dict_to_bins(fx["ref_tri2count"]))
        assert [int(v) for v in g[2]] == fx["log"]
