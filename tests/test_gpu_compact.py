"""Compact quality stream (hm_bq_compact): the host bitmap + exceptions expand on the device to exactly the stream
the plain upload ships, so every downstream result is identical.  GPU."""
import numpy as np
import pytest

import cases
from himut_b200 import abi, bamdec, gtmodel, lib, synth

pytestmark = pytest.mark.gpu


def _same_records(a, b):
    assert a.shape == b.shape
    order = ["chunk", "tpos", "ref", "alt"]  # records come back in no particular order (include/himut_b200.h)
    a, b = np.sort(a, order=order), np.sort(b, order=order)
    for name in a.dtype.names:
        assert np.array_equal(a[name], b[name]), name


def test_sizeof_matches_library():
    import ctypes as C
    assert lib.load().hm_abi_sizeof(4) == C.sizeof(abi.hm_bq_compact)


@pytest.mark.parametrize("seed,n", [(61, 300_000), (62, 1_000_000)])
def test_compact_call_equals_plain_call(ctx, seed, n):
    d = synth.generate(n, seed=seed)
    p = gtmodel.make_params(**gtmodel.DEFAULT_CALL_ARGS)
    chunks = d.batch.chunk_table(cases.chunkloci(0, n))
    ctx.set_params(p)
    ctx.set_site_sets()
    rec, log = ctx.call_batch(d.batch, chunks)
    cq = bamdec.compact_bq(d.batch, threads=3)
    assert cq.nbytes() < 0.3 * d.batch.bq.nbytes
    rec2, log2 = ctx.call_batch_compact(d.batch, cq, chunks)
    _same_records(rec, rec2)
    assert list(log) == list(log2)
    # the expanded stream itself: per-read quality sums and the normcounts path read every byte of it
    ctx.upload(d.batch)
    s1 = ctx.read_stats(d.batch)["bq_total"].copy()
    n1 = ctx.normcounts_chunks(d.ref, chunks)
    ctx.upload_compact(d.batch, cq)
    assert np.array_equal(ctx.read_stats(d.batch)["bq_total"], s1)
    n2 = ctx.normcounts_chunks(d.ref, chunks)
    assert all(np.array_equal(x, y) for x, y in zip(n1[:3], n2[:3])) and n1[3] == n2[3]


@pytest.mark.parametrize("seed", [3, 4, 5])
def test_compact_adversarial_qualities(ctx, seed):
    """random qualities (most bases are exceptions), soft clips, odd read lengths"""
    batch, ref = cases.adversarial_batch(seed, contig_len=4000, n_reads=200, max_len=1500)
    p = gtmodel.make_params(**cases.call_args(min_qv=0, min_mapq=0, qlen_lower_limit=0, qlen_upper_limit=100000, min_bq=1))
    chunks = batch.chunk_table([(0, 2000), (2000, 4000)])
    ctx.set_params(p)
    ctx.set_site_sets()
    rec, log = ctx.call_batch(batch, chunks)
    cq = bamdec.compact_bq(batch, threads=2)
    rec2, log2 = ctx.call_batch_compact(batch, cq, chunks)
    _same_records(rec, rec2)
    assert list(log) == list(log2)
    ctx.upload(batch)
    s1 = ctx.read_stats(batch)["bq_total"].copy()
    ctx.upload_compact(batch, cq)
    assert np.array_equal(ctx.read_stats(batch)["bq_total"], s1)


def test_compact_rejects_inconsistent_offsets(ctx):
    d = synth.generate(60_000, seed=63)
    cq = bamdec.compact_bq(d.batch)
    bad = abi.BqCompact(cq.mask, cq.exc, cq.exc_off[::-1].copy(), cq.modal, cq.n_exc)
    with pytest.raises(lib.HimutError):
        ctx.upload_compact(d.batch, bad)
