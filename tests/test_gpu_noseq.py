"""A batch without a base stream (hm_read_batch.seq == NULL): `call` and the phase edges take the bases of match
runs from the site's reference allele (what a cs match means, src/himut/cslib.py:22-29) and must return exactly
what the batch with its bases returns — the oracle always reads the bases.  The CPU half checks the premise on the
test data itself: inside match runs the packed bases are the reference's."""
import numpy as np
import pytest

import cases
import parity
from himut_b200 import abi, bamdec, gtmodel, synth
from oracle import oracle

CALL_CASES = [n for n in cases.CASES if n.startswith("call_")]


def _match_runs_equal_reference(batch, ref):
    refc = np.array([abi.BASE2CODE.get(ch, 4) for ch in (ref if isinstance(ref, str) else ref.decode())], np.uint8)
    checked = 0
    for r in range(batch.n_reads):
        s0 = int(batch.seq_off[r])
        nb = (int(batch.qlen[r]) + 3) // 4
        packed = batch.seq[s0:s0 + nb]
        bases = ((packed[:, None] >> (2 * np.arange(4, dtype=np.uint8))) & 3).reshape(-1)
        t, q = int(batch.tstart[r]), int(batch.qstart[r])
        o0 = int(batch.op_off[r])
        for w in batch.ops[o0:o0 + int(batch.n_ops[r])]:
            kind, v = int(w) & 3, int(w) >> 2
            if kind == abi.OP_MATCH:
                if not np.array_equal(bases[q:q + v], refc[t:t + v]):
                    return False
                checked += v
                t += v; q += v
            elif kind == abi.OP_SUB:
                t += 1; q += 1
            elif kind == abi.OP_INS:
                q += v
            else:
                t += v
    return checked > 0


@pytest.mark.parametrize("name", ["call_basic", "call_adversarial_a", "call_adversarial_b"])
def test_premise_match_runs_carry_reference_bases(name):
    c = cases.build_case(name)
    assert _match_runs_equal_reference(c["batch"], c["ref"])


def test_without_seq_struct():
    c = cases.build_case("call_basic")
    b = c["batch"].without_seq()
    s = b.as_struct()
    assert s.seq is None and s.seq_off is None and s.seq_bytes == 0
    assert s.bq_bytes == c["batch"].bq.size and s.n_reads == c["batch"].n_reads
    assert b.bq is c["batch"].bq and b.ops is c["batch"].ops


@pytest.mark.gpu
@pytest.mark.parametrize("name", CALL_CASES)
def test_call_without_seq_matches_oracle(ctx, name):
    c = cases.build_case(name)
    ctx.set_params(c["params"])
    ctx.set_site_sets(c["common"], c["pon"])
    if c["phase"] is not None:
        ctx.set_phase_sets(c["phase"])
    rec, log = ctx.call_batch(c["batch"].without_seq(), c["chunk_table"])
    o_rec, o_log = oracle.call_chunks(c["params"], c["batch"], c["chunk_table"], c["common"], c["pon"], c["phase"])
    ok, why = parity.records_equal(rec, o_rec)
    assert ok, why
    assert list(log) == list(o_log)


@pytest.mark.gpu
def test_compact_without_seq_one_megabase(ctx):
    d = synth.generate(1_000_000, seed=71)
    p = gtmodel.make_params(**gtmodel.DEFAULT_CALL_ARGS)
    chunks = d.batch.chunk_table(cases.chunkloci(0, 1_000_000))
    ctx.set_params(p)
    ctx.set_site_sets()
    cq = bamdec.compact_bq(d.batch, threads=3)
    rec, log = ctx.call_batch_compact(d.batch.without_seq(), cq, chunks)
    o_rec, o_log = oracle.call_chunks(p, d.batch, chunks)
    ok, why = parity.records_equal(rec, o_rec)
    assert ok, why
    assert list(log) == list(o_log)
    # back to a batch with bases on the same context
    rec2, log2 = ctx.call_batch(d.batch, chunks)
    ok, why = parity.records_equal(rec2, o_rec)
    assert ok, why


@pytest.mark.gpu
@pytest.mark.parametrize("slots", [None, "64", "96"])
def test_deep_pileup_without_seq(ctx, slots, monkeypatch):
    """400x: the entry table has about twice the batch's mean depth in slots per site (512 at most: here every site is
    deeper than that, and with HIMUT_B200_SITE_SLOTS forced to 64 / 96 far deeper) — k_site_reduce gathers the slots
    beyond the table itself, with the same result"""
    if slots is not None:
        monkeypatch.setenv("HIMUT_B200_SITE_SLOTS", slots)  # read at upload
    d = synth.generate(30_000, seed=34, depth=400.0)
    p = gtmodel.make_params(**cases.call_args(md_threshold=1000))
    chunks = d.batch.chunk_table([(0, 12_345), (12_345, 30_000)])
    ctx.set_params(p)
    ctx.set_site_sets()
    rec, log = ctx.call_batch(d.batch.without_seq(), chunks)
    o_rec, o_log = oracle.call_chunks(p, d.batch, chunks)
    ok, why = parity.records_equal(rec, o_rec)
    assert ok, why
    assert list(log) == list(o_log)


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [60.0, 100.0])
def test_deeper_samples_stay_inside_the_entry_table(ctx, depth):
    """60x / 100x: the slots follow the depth (128 / 224), so no site takes the walk; records equal the oracle's"""
    d = synth.generate(60_000, seed=35, depth=depth)
    p = gtmodel.make_params(**cases.call_args(md_threshold=1000))
    chunks = d.batch.chunk_table([(0, 30_000), (30_000, 60_000)])
    ctx.set_params(p)
    ctx.set_site_sets()
    rec, log = ctx.call_batch(d.batch.without_seq(), chunks)
    assert ctx.last_call_path() == 2
    o_rec, o_log = oracle.call_chunks(p, d.batch, chunks)
    ok, why = parity.records_equal(rec, o_rec)
    assert ok, why
    assert list(log) == list(o_log)


@pytest.mark.gpu
def test_phase_edges_without_seq(ctx):
    d = synth.generate(400_000, seed=72)
    t = synth.phase_table(d.germ, d.spec.phase_block)
    hpos, href = t["hpos"], t["href"]
    assert hpos.size > 50
    exp, _need = oracle.phase_edges(d.batch, hpos, href, 256, 1, 20)
    assert int(exp.sum()) > 1000
    ctx.upload(d.batch.without_seq())
    ctx.phase_edges_begin(hpos, href, 256)
    assert ctx.phase_edges_add(1, 20) == 0
    assert np.array_equal(ctx.phase_edges_end(), exp)


@pytest.mark.gpu
def test_normcounts_refuses_a_batch_without_seq(ctx):
    d = synth.generate(50_000, seed=73)
    p = gtmodel.make_params(**gtmodel.DEFAULT_CALL_ARGS)
    ctx.set_params(p)
    ctx.set_site_sets()
    ctx.upload(d.batch.without_seq())
    with pytest.raises(RuntimeError, match="base stream"):
        ctx.normcounts_chunks(d.ref, d.batch.chunk_table([(0, 50_000)]))
    ctx.upload(d.batch)
    o = oracle.normcounts_chunks(p, d.batch, d.ref, d.batch.chunk_table([(0, 50_000)]))
    g = ctx.normcounts_chunks(d.ref, d.batch.chunk_table([(0, 50_000)]))
    assert np.array_equal(g[0], o[0]) and list(g[2]) == list(o[2])
