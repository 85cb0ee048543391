"""Parity of the CUDA `himut call` path (through the C ABI) with the CPU oracle and with the
reference's own outputs (golden fixtures).  Needs a B200: `pytest -m gpu`."""
import os

import numpy as np
import pytest

import cases
import parity
from himut_b200 import abi, gtmodel, lib, records, synth
from oracle import oracle

pytestmark = pytest.mark.gpu
CALL_CASES = [n for n in cases.CASES if n.startswith("call_")]


def run_gpu(ctx, c, resident=False):
    ctx.set_params(c["params"])
    ctx.set_site_sets(c["common"], c["pon"])
    if c["phase"] is not None:
        ctx.set_phase_sets(c["phase"])
    if resident:
        ctx.upload(c["batch"])
        return ctx.call_chunks(c["chunk_table"])
    return ctx.call_batch(c["batch"], c["chunk_table"])


@pytest.mark.parametrize("name", CALL_CASES)
def test_call_matches_oracle_and_reference(ctx, name):
    c = cases.build_case(name)
    rec, log = run_gpu(ctx, c)
    o_rec, o_log = oracle.call_chunks(c["params"], c["batch"], c["chunk_table"], c["common"], c["pon"], c["phase"])
    ok, why = parity.records_equal(rec, o_rec)
    assert ok, why
    assert list(log) == list(o_log)
    fx = parity.load_golden(name)
    rows = records.records_to_tsbs_lst(cases.CHROM, rec)
    gold = parity.golden_rows(fx)
    assert parity.rows_equal(rows, gold), parity.first_diff(rows, gold, rec)
    assert [int(v) for v in log] == fx["expected"]["log"]


def test_read_stats_match_oracle(ctx):
    for name in ("call_basic", "call_adversarial_a"):
        c = cases.build_case(name)
        ctx.set_params(c["params"])
        ctx.upload(c["batch"])
        g = ctx.read_stats(c["batch"])
        o = oracle.read_stats(c["batch"])
        for k in o:
            assert np.array_equal(g[k], o[k]), k


def test_config0_one_megabase(ctx):
    """BASELINE configs[0]: 1 Mb, 30x, five 200 kb chunks — full oracle comparison"""
    d = synth.generate(1_000_000)
    p = gtmodel.make_params(**gtmodel.DEFAULT_CALL_ARGS)
    chunks = d.batch.chunk_table(cases.chunkloci(0, 1_000_000))
    ctx.set_params(p)
    ctx.set_site_sets()
    rec, log = ctx.call_batch(d.batch, chunks)
    o_rec, o_log = oracle.call_chunks(p, d.batch, chunks)
    ok, why = parity.records_equal(rec, o_rec)
    assert ok, why
    assert list(log) == list(o_log)
    assert int((rec["status"] == abi.ST_PASS).sum()) > 50


def test_resident_and_e2e_agree_and_are_idempotent(ctx):
    c = cases.build_case("call_sets")
    a, la = run_gpu(ctx, c, resident=True)
    b, lb = run_gpu(ctx, c, resident=True)
    e, le = run_gpu(ctx, c, resident=False)
    for x, lx in ((b, lb), (e, le)):
        ok, why = parity.records_equal(a, x)
        assert ok, why
        assert list(la) == list(lx)


def test_chunking_invariance(ctx):
    """property: away from chunk borders, how a contig is cut into chunks does not change a site's
    verdict (the pileup at a site only depends on the reads covering it)"""
    d = synth.generate(120_000, seed=77)
    p = gtmodel.make_params(**gtmodel.DEFAULT_CALL_ARGS)
    ctx.set_params(p)
    ctx.set_site_sets()
    one, _ = ctx.call_batch(d.batch, d.batch.chunk_table([(0, 120_000)]))
    cuts = [(0, 40_000), (40_000, 81_000), (81_000, 120_000)]
    three, _ = ctx.call_batch(d.batch, d.batch.chunk_table(cuts))
    border = {40_000, 81_000}
    k = lambda r: {(int(x["tpos"]), int(x["ref"]), int(x["alt"])): (int(x["status"]), int(x["gq"]), tuple(x["counts"]))
                   for x in r if int(x["tpos"]) not in border}
    assert k(one) == k(three)


def test_empty_and_ragged_inputs(ctx):
    p = gtmodel.make_params(**gtmodel.DEFAULT_CALL_ARGS)
    ctx.set_params(p)
    ctx.set_site_sets()
    d = synth.generate(60_000, seed=5)
    # a chunk with no reads, a chunk list of length 0, a one-read batch
    chunks = d.batch.chunk_table([(0, 30_000)])
    chunks["read_hi"] = chunks["read_lo"]
    rec, log = ctx.call_batch(d.batch, chunks)
    assert rec.size == 0 and not log.any()
    rec, log = ctx.call_batch(d.batch, chunks[:0])
    assert rec.size == 0 and not log.any()
    one = d.batch.select([3])
    ch = one.chunk_table([(0, 60_000)])
    rec, log = ctx.call_batch(one, ch)
    o_rec, o_log = oracle.call_chunks(p, one, ch)
    ok, why = parity.records_equal(rec, o_rec)
    assert ok, why
    assert list(log) == list(o_log)


def test_bq_zero_fails_loudly(ctx):
    d = synth.generate(60_000, seed=5)
    d.batch.bq[:] = np.where(d.batch.bq == 93, 0, d.batch.bq)
    p = gtmodel.make_params(**dict(gtmodel.DEFAULT_CALL_ARGS, min_qv=0))
    ctx.set_params(p)
    ctx.set_site_sets()
    with pytest.raises(lib.HimutError) as e:
        ctx.call_batch(d.batch, d.batch.chunk_table([(0, 60_000)]))
    assert e.value.code == abi.HM_ERR_BQ_ZERO


def test_malformed_batch_is_rejected(ctx):
    d = synth.generate(60_000, seed=5)
    p = gtmodel.make_params(**gtmodel.DEFAULT_CALL_ARGS)
    ctx.set_params(p)
    d.batch.tstart[5] = d.batch.tstart[4] - 1  # not coordinate sorted
    d.batch._struct = None
    with pytest.raises(lib.HimutError) as e:
        ctx.upload(d.batch)
    assert e.value.code == abi.HM_ERR_ARG


def test_async_record_copy_equals_synchronous(ctx):
    """hm_call_chunks_async: the record copy overlaps the next call; after records_wait the buffers hold exactly
    what the synchronous call returns, for several calls in flight"""
    from himut_b200 import gtmodel, synth
    d = synth.generate(400_000, seed=71)
    p = gtmodel.make_params(**gtmodel.DEFAULT_CALL_ARGS)
    loci = cases.chunkloci(0, 400_000)
    tables = [d.batch.chunk_table(loci), d.batch.chunk_table(loci[:1]), d.batch.chunk_table(loci[1:])]
    ctx.set_params(p)
    ctx.set_site_sets()
    ctx.upload(d.batch)
    sync = [ctx.call_chunks(t) for t in tables]
    for _ in range(2):
        pend = []
        for t in tables[:2]:
            pend.append(ctx.call_chunks(t, view=True, wait=False))
        ctx.records_wait()
        for (rec, log), (srec, slog) in zip(pend, sync[:2]):
            assert list(log) == list(slog)
            ok, why = parity.records_equal(rec, srec)
            assert ok, why
    rec, log = ctx.call_chunks(tables[2], view=True, wait=False)
    ctx.records_wait()
    ok, why = parity.records_equal(rec, sync[2][0])
    assert ok, why


def test_async_with_dropped_boundary_records(ctx):
    """overlapping regions: records a previous chunk claimed are left out by the copy itself, also when it is deferred"""
    c = cases.build_case("call_adversarial_b")
    ctx.set_params(c["params"])
    ctx.set_site_sets(c["common"], c["pon"])
    ctx.upload(c["batch"])
    srec, slog = ctx.call_chunks(c["chunk_table"])
    o_rec, o_log = oracle.call_chunks(c["params"], c["batch"], c["chunk_table"], c["common"], c["pon"], c["phase"])
    ok, why = parity.records_equal(srec, o_rec)
    assert ok, why
    a = ctx.call_chunks(c["chunk_table"], view=True, wait=False)
    b = ctx.call_chunks(c["chunk_table"], view=True, wait=False)
    ctx.records_wait()
    for rec, log in (a, b):
        ok, why = parity.records_equal(rec, srec)
        assert ok, why
        assert list(log) == list(slog)


def test_async_followed_by_a_call_without_candidates(ctx):
    """the deferred record copy of an asynchronous call must not get lost when the next call has nothing to sort"""
    from himut_b200 import gtmodel, synth
    d = synth.generate(200_000, seed=72)
    ctx.set_params(gtmodel.make_params(**gtmodel.DEFAULT_CALL_ARGS))
    ctx.set_site_sets()
    ctx.upload(d.batch)
    full = d.batch.chunk_table(cases.chunkloci(0, 200_000))
    srec, slog = ctx.call_chunks(full)
    empty = d.batch.chunk_table([(5, 6)])          # a one-base window: no candidate
    a = ctx.call_chunks(full, view=True, wait=False)
    b = ctx.call_chunks(empty, view=True, wait=False)
    c = ctx.call_chunks(full, view=True, wait=False)
    ctx.records_wait()
    assert b[0].size == 0
    for rec, log in (a, c):
        ok, why = parity.records_equal(rec, srec)
        assert ok, why
        assert list(log) == list(slog)


@pytest.mark.parametrize("name", ["call_basic", "call_sets", "call_phase", "call_adversarial_a", "call_adversarial_b", "call_lowdepth"])
def test_omit_restatements(ctx, name):
    """HM_OPT_OMIT_RESTATEMENTS: the same records minus the germline restatements (which the reference never emits),
    the same counters; synchronous, asynchronous, and through the host fallback of the boundary replay"""
    c = cases.build_case(name)
    ctx.set_params(c["params"])
    ctx.set_site_sets(c["common"], c["pon"])
    if c["phase"] is not None:
        ctx.set_phase_sets(c["phase"])
    ctx.upload(c["batch"])
    full, flog = ctx.call_chunks(c["chunk_table"])
    restates = np.isin(full["status"], [abi.ST_GERM_HET, abi.ST_GERM_HETALT, abi.ST_GERM_HOMALT, abi.ST_GERM_HOMREF])
    want = full[~restates]
    ctx.omit_restatements(True)
    try:
        rec, log = ctx.call_chunks(c["chunk_table"])
        assert list(log) == list(flog)
        ok, why = parity.records_equal(rec, want)
        assert ok, why
        a = ctx.call_chunks(c["chunk_table"], view=True, wait=False)
        b = ctx.call_chunks(c["chunk_table"], view=True, wait=False)
        ctx.records_wait()
        for r, l in (a, b):
            assert list(l) == list(flog)
            ok, why = parity.records_equal(r, want)
            assert ok, why
        os.environ["HIMUT_B200_BOUNDARY_CAP"] = "2"
        try:
            rec, log = ctx.call_chunks(c["chunk_table"])
        finally:
            del os.environ["HIMUT_B200_BOUNDARY_CAP"]
        assert list(log) == list(flog)
        ok, why = parity.records_equal(rec, want)
        assert ok, why
    finally:
        ctx.omit_restatements(False)
