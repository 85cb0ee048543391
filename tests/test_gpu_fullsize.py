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
    # oracle on a random sample of chunks
    rng = np.random.default_rng(7)
    idx = np.sort(rng.choice(320, size=24, replace=False))
    sample = chunks[idx]
    g = ctx.normcounts_chunks(d.ref, sample)
    o = oracle.normcounts_chunks(p, d.batch, d.ref, sample)
    assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1])
    assert list(g[2]) == list(o[2]) and g[3] == o[3]
