"""`himut phase` edge counting: the oracle against the reference's own get_edges output
(CPU), the CUDA band tables against the oracle and the worker mirror against the reference (GPU)."""
import importlib.util
import json
import os

import numpy as np
import pytest

import cases
from oracle import oracle

_spec = importlib.util.spec_from_file_location("mkedges", os.path.join(cases.GOLDEN_DIR, "make_golden_edges.py"))
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)
BASE2CODE = {"A": 0, "T": 1, "G": 2, "C": 3}


def _expected(name):
    return json.load(open(os.path.join(cases.GOLDEN_DIR, "edges.json")))["expected"][name]


def _band_to_rows(table):
    a, d = np.nonzero(table.any(axis=2))
    return [[int(x), int(x + y + 1)] + [int(v) for v in table[x, y]] for x, y in zip(a.tolist(), d.tolist())]


@pytest.mark.parametrize("name", sorted(mk.CASES))
def test_oracle_matches_reference(name):
    batch, hetsnp_lst, _h2i, _n, min_bq, min_mapq = mk.inputs(name)
    hpos = np.array([h[0] for h in hetsnp_lst], np.int32)
    href = np.array([BASE2CODE[h[1]] for h in hetsnp_lst], np.uint8)
    table, need = oracle.phase_edges(batch, hpos, href, 256, min_bq, min_mapq)
    assert need == 0
    assert sorted(_band_to_rows(table)) == sorted(_expected(name))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mk.CASES))
def test_cuda_matches_oracle(ctx, name):
    batch, hetsnp_lst, _h2i, _n, min_bq, min_mapq = mk.inputs(name)
    hpos = np.array([h[0] for h in hetsnp_lst], np.int32)
    href = np.array([BASE2CODE[h[1]] for h in hetsnp_lst], np.uint8)
    ctx.upload(batch)
    ctx.phase_edges_begin(hpos, href, 256)
    assert ctx.phase_edges_add(min_bq, min_mapq) == 0
    got = ctx.phase_edges_end()
    exp, _ = oracle.phase_edges(batch, hpos, href, 256, min_bq, min_mapq)
    assert np.array_equal(got, exp)
    # a band that is too narrow is reported, not silently truncated
    ctx.phase_edges_begin(hpos, href, 2)
    need = ctx.phase_edges_add(min_bq, min_mapq)
    assert need == oracle.phase_edges(batch, hpos, href, 2, min_bq, min_mapq)[1] and need > 2
    ctx.phase_edges_end()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["plain", "bq0_counts_deletions"])
def test_worker_mirror_matches_reference(ctx, tmp_path, monkeypatch, name):
    from himut_b200 import bamdec, phaselib, worker
    batch, hetsnp_lst, h2i, n, min_bq, min_mapq = mk.inputs(name)
    path = str(tmp_path / "e.bam")
    bamdec.write_batch_bam(path, "chr1", n, batch)
    monkeypatch.setattr(worker, "GROUP_SPAN", 37_000)  # several decode windows: shared reads must be counted once
    edge_lst, e2c = phaselib.get_edges("chr1", path, min_bq, min_mapq, [h[0] for h in hetsnp_lst], hetsnp_lst, h2i)
    got = [[int(i), int(j)] + [int(v) for v in e2c[(i, j)]] for (i, j) in edge_lst]
    assert got == _expected(name)
    assert all(isinstance(v, np.ndarray) and v.dtype == np.float64 for v in e2c.values())


@pytest.mark.gpu
def test_cuda_matches_oracle_2mb(ctx):
    """a contig-sized run: 4 000 reads, 2 000 hetSNPs, default `himut phase` thresholds"""
    from himut_b200 import synth
    d = synth.generate(2_000_000, seed=57)
    het = d.germ["gt"] < 2
    hpos = d.germ["pos"][het].astype(np.int32)
    href = d.germ["ref"][het].astype(np.uint8)
    ctx.upload(d.batch)
    ctx.phase_edges_begin(hpos, href, 128)
    assert ctx.phase_edges_add(20, 20) == 0
    got = ctx.phase_edges_end()
    exp, need = oracle.phase_edges(d.batch, hpos, href, 128, 20, 20)
    assert need == 0 and np.array_equal(got, exp) and int(got.sum()) > 100_000
