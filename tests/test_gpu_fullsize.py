"""Parity at BASELINE.json's full single-GPU size (configs[1]: 64 Mb contig, 30x, 1.92 G aligned
bases, 320 chunks).  The C oracle is fast enough to check `himut call` in full; the callable-base
half of normcounts is checked through additivity over the whole contig plus an oracle comparison
on a random sample of chunks.  GPU, a few minutes."""
import numpy as np
import pytest

import cases
import parity
from himut_b200 import abi, gtmodel, synth
from oracle import oracle

pytestmark = pytest.mark.gpu
CONTIG = 64_000_000


@pytest.fixture(scope="module")
def big():
    d = synth.generate(CONTIG, seed=20260101, copy=False)
    p = gtmodel.make_params(**gtmodel.DEFAULT_CALL_ARGS)
    chunks = d.batch.chunk_table(cases.chunkloci(0, CONTIG))
    return d, p, chunks


def test_call_full_size_equals_oracle(ctx, big):
    d, p, chunks = big
    ctx.set_params(p)
    ctx.set_site_sets()
    rec, log = ctx.call_batch(d.batch, chunks)
    # structural properties that hold at any size
    assert rec.size == int(log[1])
    key = (rec["chunk"].astype(np.int64) << 36) | (rec["tpos"].astype(np.int64) << 4) | (rec["ref"] << 2) | rec["alt"]
    assert np.all(np.diff(np.sort(key)) > 0), "records are distinct in (chunk, tpos, ref, alt); they come back in no particular order"
    assert np.all((rec["tpos"] >= chunks["start"][rec["chunk"]]) & (rec["tpos"] <= chunks["end"][rec["chunk"]]))
    assert int(log[6]) == int(log[8:15].sum())
    assert int(log[1]) == int(log[2:8].sum()) + int((rec["status"] == abi.ST_GERM_HOMREF).sum())
    # and the whole thing against the oracle
    o_rec, o_log = oracle.call_chunks(p, d.batch, chunks, cap=int(rec.size) + 1024)
    ok, why = parity.records_equal(rec, o_rec)
    assert ok, why
    assert list(log) == list(o_log)
    assert int((rec["status"] == abi.ST_PASS).sum()) > 10_000


def test_normcounts_full_size(ctx, big):
    d, p, chunks = big
    ctx.set_params(p)
    ctx.set_site_sets()
    ctx.upload(d.batch)
    whole = ctx.normcounts_chunks(d.ref, chunks)
    # additivity over the contig's quarters (num_ccs, log[0], is a distinct count and not additive)
    parts = [ctx.normcounts_chunks(d.ref, chunks[i:i + 80]) for i in range(0, 320, 80)]
    assert np.array_equal(whole[0], sum(x[0] for x in parts))
    assert np.array_equal(whole[1], sum(x[1] for x in parts))
    assert list(whole[2][1:]) == list(sum(x[2] for x in parts)[1:])
    assert int(whole[2][1]) == int(whole[2][2:7].sum())          # every counted base lands in one category
    assert int(whole[2][6]) == int(whole[2][7:14].sum())         # homref bases split over the filters
    # the whole contig against the oracle: every one of the 320 chunks, the oracle on one thread per host core over
    # runs of consecutive chunks (the tallies are additive over chunks; the distinct-read count is compared per run)
    import os
    from concurrent.futures import ThreadPoolExecutor
    n_thr = max(1, min(32, len(os.sched_getaffinity(0))))
    per = -(-len(chunks) // n_thr)
    groups = [chunks[i:i + per] for i in range(0, len(chunks), per)]
    with ThreadPoolExecutor(n_thr) as ex:
        o_parts = list(ex.map(lambda sub: oracle.normcounts_chunks(p, d.batch, d.ref, sub), groups))
    assert np.array_equal(whole[0], sum(x[0] for x in o_parts)), "ccs tri counts differ from the oracle over the whole contig"
    assert np.array_equal(whole[1], sum(x[1] for x in o_parts)), "ref tri counts differ from the oracle over the whole contig"
    assert list(whole[2][1:]) == list(sum(x[2] for x in o_parts)[1:])
    assert whole[3] == sum(x[3] for x in o_parts)
    for sub, o in zip(groups[:3], o_parts[:3]):  # and the distinct-read counter on the first runs
        g = ctx.normcounts_chunks(d.ref, sub)
        assert list(g[2]) == list(o[2])


def test_call_phase_full_size_equals_oracle(ctx):
    """BASELINE configs[3] at size: `himut call --phase` on the 64 Mb contig with a phased germline table (phase sets
    of ~200 kb, the chunks are their spans: vcflib.py:655-662) and common-SNP + panel-of-normals sets, every record
    against the oracle; both ways of shipping the batch (with and without the 2-bit base stream)"""
    d = synth.generate(CONTIG, seed=20260103, copy=False, phase_block=200_000)
    ph, chunk_spans, sets = cases.phase_case(d, 200_000)
    common, pon = cases.site_sets_from_synth(d, 3)
    p = gtmodel.make_params(**dict(gtmodel.DEFAULT_CALL_ARGS, phase=True))
    chunks = d.batch.chunk_table(chunk_spans, sets)
    assert len(chunk_spans) > 250 and common.size > 50_000 and pon.size > 10_000
    o_rec, o_log = oracle.call_chunks(p, d.batch, chunks, common, pon, ph, cap=int(d.batch.ops.size))
    for batch in (d.batch.without_seq(), d.batch):
        ctx.set_params(p)
        ctx.set_site_sets(common, pon)
        ctx.set_phase_sets(ph)
        rec, log = ctx.call_batch(batch, chunks)
        assert ctx.last_call_path() == 2
        ok, why = parity.records_equal(rec, o_rec)
        assert ok, why
        assert list(log) == list(o_log)
    st = rec["status"]
    assert int((st == abi.ST_PASS).sum()) > 1000
    assert int((st == abi.ST_COMSNP).sum()) > 100 and int((st == abi.ST_PON).sum()) > 100
    assert int((rec["phase_set"] >= 0).sum()) == int((st == abi.ST_PASS).sum())
