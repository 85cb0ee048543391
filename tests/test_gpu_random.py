"""Randomised parity sweep: many small dirty batches (dense substitutions / indels, N reference
bases, soft clips, secondary records, duplicate names, random BQ), random worker parameters,
random overlapping chunk lists, random phase tables and site sets — CUDA against the CPU oracle,
every record field and every counter.  GPU."""
import random

import numpy as np
import pytest

import cases
import parity
from himut_b200 import gtmodel
from oracle import oracle

pytestmark = pytest.mark.gpu


def random_setup(seed):
    rnd = random.Random(seed)
    n = rnd.choice([1500, 3000, 6000])
    batch, ref = cases.adversarial_batch(seed, contig_len=n, n_reads=rnd.choice([40, 120, 300]), max_len=rnd.choice([300, 900, 2000]))
    args = cases.call_args(
        min_qv=rnd.choice([0, 20, 30]), min_mapq=rnd.choice([0, 20, 60]), qlen_lower_limit=rnd.choice([0, 30, 200]),
        qlen_upper_limit=rnd.choice([500, 900, 5000]), min_sequence_identity=rnd.choice([0.0, 0.9, 0.99]),
        min_gq=rnd.choice([0, 5, 20]), min_bq=rnd.choice([1, 30, 93]), min_trim=rnd.choice([0.0, 0.01, 0.1]),
        max_mismatch_count=rnd.choice([0, 0, 1, 3]), mismatch_window=rnd.choice([0, 5, 20, 40]),
        md_threshold=rnd.choice([10, 45, 1000]), min_ref_count=rnd.choice([0, 2, 5]), min_alt_count=rnd.choice([1, 2]),
        min_hap_count=rnd.choice([0, 1, 3]), germline_snv_prior=rnd.choice([1e-3, 1e-2]))
    # chunk list: sometimes the reference's own tiling, sometimes overlapping / unordered windows
    if rnd.random() < 0.5:
        cuts = sorted(rnd.sample(range(1, n), rnd.choice([1, 2, 4])))
        edges = [0] + cuts + [n]
        chunks = [(edges[i], edges[i + 1]) for i in range(len(edges) - 1)]
    else:
        chunks = []
        for _ in range(rnd.choice([1, 3, 5])):
            a = rnd.randrange(0, n - 10)
            chunks.append((a, min(n, a + rnd.randrange(5, n))))
    # site sets drawn from positions that exist
    keys = [((rnd.randrange(1, n) << 4) | (rnd.randrange(4) << 2) | rnd.randrange(4)) for _ in range(n // 4)]
    common = np.unique(np.array(keys[: len(keys) // 2], np.uint64))
    pon = np.unique(np.array(keys[len(keys) // 2:], np.uint64))
    return rnd, batch, ref, args, chunks, common, pon


def random_phase(rnd, ref, chunks):
    """one phase set per chunk: random hetSNPs inside the chunk's window (some with non-matching alleles)"""
    hpos, href, halt, hbit, set_off, new_chunks = [], [], [], [], [0], []
    for (s, e) in chunks:
        cand = sorted(rnd.sample(range(max(s, 1), max(e, s + 2)), min(rnd.choice([2, 6, 20]), max(e - s - 1, 1))))
        cand = [p for p in cand if ref[p - 1] in "ATGC"]
        if len(cand) < 1:
            continue
        for p in cand:
            r = "ATGC".index(ref[p - 1])
            hpos.append(p); href.append(r); halt.append(rnd.choice([x for x in range(4) if x != r])); hbit.append(rnd.randrange(2))
        set_off.append(len(hpos))
        new_chunks.append((cand[0], cand[-1]))  # the reference's phase chunks: (first hpos, last hpos)
    ph = dict(hpos=np.array(hpos, np.int32), href=np.array(href, np.uint8), halt=np.array(halt, np.uint8),
              hbit=np.array(hbit, np.uint8), set_off=np.array(set_off, np.uint64))
    return ph, new_chunks


@pytest.mark.parametrize("seed", range(1000, 1030))
def test_random_call(ctx, seed):
    rnd, batch, ref, args, chunks, common, pon = random_setup(seed)
    phase = None
    sets = None
    if seed % 3 == 0:
        phase, chunks = random_phase(rnd, ref, chunks)
        if not chunks:
            pytest.skip("no phase set")
        args["phase"] = True
        sets = list(range(len(chunks)))
    p = gtmodel.make_params(**args)
    table = batch.chunk_table(chunks, sets)
    ctx.set_params(p)
    ctx.set_site_sets(common, pon)
    if phase is not None:
        ctx.set_phase_sets(phase)
    rec, log = ctx.call_batch(batch, table)
    o_rec, o_log = oracle.call_chunks(p, batch, table, common, pon, phase)
    ok, why = parity.records_equal(rec, o_rec)
    assert ok, why
    assert list(log) == list(o_log)


@pytest.mark.parametrize("seed", range(2000, 2030))
def test_random_normcounts(ctx, seed):
    rnd, batch, ref, args, chunks, common, pon = random_setup(seed)
    ref = "".join(c.lower() if rnd.random() < 0.02 else c for c in ref)
    phase = None
    sets = None
    if seed % 3 == 0:
        phase, chunks = random_phase(rnd, ref.upper(), chunks)
        if not chunks:
            pytest.skip("no phase set")
        args["phase"] = True
        sets = list(range(len(chunks)))
    p = gtmodel.make_params(**args)
    table = batch.chunk_table(chunks, sets)
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
