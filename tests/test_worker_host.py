"""Host logic of the drop-in workers on the CPU: himut_b200.caller / normcounts / phaselib run here against a
stand-in context whose device calls are answered by the oracle (test infrastructure only), through real BAM files
and the native decoder, and must reproduce the reference's own outputs (tests/golden/).  What this pins without a
GPU: chunk grouping, the som_seen and distinct-read carry across groups, the log vector, rows and their order, the
tri-count dictionaries.  The same workers with the real context are tests/test_gpu_worker.py."""
import importlib.util
import os

import numpy as np
import pytest

import cases
import parity
from himut_b200 import abi, bamio, caller, normcounts, phaselib, worker
from standin import OracleContext


@pytest.fixture
def octx(monkeypatch):
    ctx = OracleContext()
    monkeypatch.setattr(worker, "context", lambda: ctx)
    plain = worker.RegionSource.batch
    monkeypatch.setattr(worker.RegionSource, "batch",
                        lambda self, chrom, loci, phase_sets=None, seq=True, **kw: plain(self, chrom, loci, phase_sets, seq=True, **kw))
    return ctx


def _inputs(c, tmp_path):
    bam = str(tmp_path / (c["name"] + ".bam"))
    bamio.write_batch_bam(bam, cases.CHROM, c["contig_len"], c["batch"])
    common = pon = None
    if c["common_vcf"].size:
        common = str(tmp_path / "common.vcf.bgz")
        cases.write_sites_vcf(common, cases.CHROM, c["common_vcf"])
    if c["pon_vcf"].size:
        pon = str(tmp_path / "pon.vcf.bgz")
        cases.write_sites_vcf(pon, cases.CHROM, c["pon_vcf"])
    return bam, common, pon


def _run_call(c, tmp_path):
    a = c["args"]
    bam, common, pon = _inputs(c, tmp_path)
    hbit, hpos, hetsnp = cases.phase_dicts(c)
    lst, log = {}, {}
    caller.get_somatic_substitutions(
        cases.CHROM, bam, common, pon, [(cases.CHROM, s, e) for s, e in c["chunks"]], hbit, hpos, hetsnp,
        a["min_qv"], a["min_mapq"], a["qlen_lower_limit"], a["qlen_upper_limit"], a["min_sequence_identity"],
        a["min_gq"], a["min_bq"], a["min_trim"], a["max_mismatch_count"], a["mismatch_window"], a["md_threshold"],
        a["min_ref_count"], a["min_alt_count"], a["min_hap_count"], 1e-6, a["germline_snv_prior"], 1e-4,
        bool(a.get("phase")), bool(a.get("non_human_sample")), bool(a.get("create_panel_of_normals")), lst, log)
    return lst[cases.CHROM], log[cases.CHROM]


@pytest.mark.parametrize("name", ["call_basic", "call_sets", "call_pon_params", "call_phase", "call_phase_indel", "call_adversarial_a"])
def test_call_worker_host_logic(octx, name, tmp_path):
    c = cases.build_case(name)
    fx = parity.load_golden(name)
    rows, log = _run_call(c, tmp_path)
    gold = parity.golden_rows(fx)
    assert parity.rows_equal(rows, gold), parity.first_diff(rows, gold)
    assert log == fx["expected"]["log"]


def test_duplicate_query_names_in_phase_mode(octx, tmp_path):
    """records sharing a query name: the reference classifies the records it re-fetches at a phase-checked site by
    name (caller.py:556-567); the worker (whole contig in one group, and cut into several) reproduces its rows"""
    c = cases.random_case("call", cases.RANDOM_DUPNAME_CALL_SEEDS[0])
    c["name"] = "dupnames"
    fx = parity.load_random_sweep()["call"][str(cases.RANDOM_DUPNAME_CALL_SEEDS[0])]
    rows, log = _run_call(c, tmp_path)
    assert parity.rows_digest(rows) == fx["rows_sha256"] and log == fx["log"]


@pytest.mark.parametrize("span", [700, 1200, 2500])
def test_call_worker_group_carry_on_cpu(octx, tmp_path, monkeypatch, span):
    """a contig fed to the device in several batches: som_seen and the distinct-read count carry across them"""
    monkeypatch.setattr(worker, "GROUP_SPAN", span)
    for name in ("call_adversarial_a", "call_adversarial_b"):
        c = cases.build_case(name)
        fx = parity.load_golden(name)
        before = octx.uploads
        rows, log = _run_call(c, tmp_path)
        assert octx.uploads - before > 1
        gold = parity.golden_rows(fx)
        assert parity.rows_equal(rows, gold), (name, parity.first_diff(rows, gold))
        assert log == fx["expected"]["log"], name


@pytest.mark.parametrize("name", ["norm_basic", "norm_phase", "norm_adversarial"])
def test_normcounts_worker_host_logic(octx, name, tmp_path, monkeypatch):
    if name == "norm_adversarial":
        monkeypatch.setattr(worker, "GROUP_SPAN", 1200)
    c = cases.build_case(name)
    e = parity.load_golden(name)["expected"]
    a = c["args"]
    bam, common, pon = _inputs(c, tmp_path)
    hbit, hpos, hetsnp = cases.phase_dicts(c)
    ccs, rt, log = {}, {}, {}
    ties = normcounts.get_callable_tricounts(
        cases.CHROM, c["ref"], bam, common, pon, [(cases.CHROM, s, e2) for s, e2 in c["chunks"]], hbit, hpos, hetsnp,
        a["min_qv"], a["min_mapq"], a["min_trim"], a["qlen_lower_limit"], a["qlen_upper_limit"],
        a["min_sequence_identity"], a["min_gq"], a["min_bq"], a["mismatch_window"], a["max_mismatch_count"],
        a["min_ref_count"], a["min_alt_count"], a["min_hap_count"], float(a["md_threshold"]), 1e-6,
        a["germline_snv_prior"], 1e-4, bool(a.get("phase")), bool(a.get("non_human_sample")), ccs, rt, log)
    if ties:
        pytest.skip("%d alt ties: the reference's own result depends on PYTHONHASHSEED here" % ties)
    for tri in normcounts.TRI_LST:
        assert ccs[cases.CHROM][tri] == e["ccs_tri2count"].get(tri, 0), tri
        assert rt[cases.CHROM][tri] == e["ref_tri2count"].get(tri, 0), tri
    assert log[cases.CHROM] == e["log"]


_spec = importlib.util.spec_from_file_location("mkedges_host", os.path.join(cases.GOLDEN_DIR, "make_golden_edges.py"))
mk_edges = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk_edges)


@pytest.mark.parametrize("name", ["plain", "bq0_counts_deletions"])
@pytest.mark.parametrize("span", [37_000, 10_000_000])
def test_phase_edge_worker_host_logic(octx, tmp_path, monkeypatch, name, span):
    """phaselib.get_edges over several decode windows (a read shared by two windows counts once) and the band
    retry, against the reference's own get_edges output"""
    import json
    from himut_b200 import bamdec
    batch, hetsnp_lst, h2i, n, min_bq, min_mapq = mk_edges.inputs(name)
    path = str(tmp_path / "e.bam")
    bamdec.write_batch_bam(path, "chr1", n, batch)
    monkeypatch.setattr(worker, "GROUP_SPAN", span)
    edge_lst, e2c = phaselib.get_edges("chr1", path, min_bq, min_mapq, [h[0] for h in hetsnp_lst], hetsnp_lst, h2i)
    got = [[int(i), int(j)] + [int(v) for v in e2c[(i, j)]] for (i, j) in edge_lst]
    assert got == json.load(open(os.path.join(cases.GOLDEN_DIR, "edges.json")))["expected"][name]
    assert all(isinstance(v, np.ndarray) and v.dtype == np.float64 for v in e2c.values())


def test_decoder_buffers_are_page_locked_only_on_request(monkeypatch):
    """PinCache: cudaHostRegister of the decoder's buffers is opt-in (HIMUT_B200_PIN_DECODE=1); each (address, size) is
    locked once, a larger array at the same address replaces the smaller lock, close() unlocks what is held"""
    calls = []

    class Ctx:
        def pin_arrays(self, arrays):
            calls.append(("pin", [a.nbytes for a in arrays]))

        def unpin_arrays(self, arrays):
            calls.append(("unpin", [a.nbytes for a in arrays]))

    big = np.zeros(1 << 23, np.uint8)
    monkeypatch.delenv("HIMUT_B200_PIN_DECODE", raising=False)
    pins = worker.PinCache(Ctx(), enabled=True)
    pins.pin([big])
    pins.close()
    assert calls == []
    monkeypatch.setenv("HIMUT_B200_PIN_DECODE", "1")
    pins = worker.PinCache(Ctx(), enabled=True)
    pins.pin([big, np.zeros(16, np.uint8), None])  # small arrays and missing ones are left alone
    pins.pin([big])                                  # already held
    pins.pin([big[: 1 << 22]])                       # a smaller view at the same address: the lock covers it
    pins.close()
    assert calls == [("pin", [1 << 23]), ("unpin", [1 << 23])]
    assert worker.PinCache(Ctx(), enabled=False).enabled is False


def test_worker_timing_laps(monkeypatch, capsys):
    monkeypatch.delenv("HIMUT_B200_WORKER_TIMING", raising=False)
    lap = worker.Laps("quiet")
    lap("a")
    lap.report()
    assert capsys.readouterr().err == ""
    monkeypatch.setenv("HIMUT_B200_WORKER_TIMING", "1")
    lap = worker.Laps("call_region chrT")
    lap("wait for decode"); lap("upload"); lap("wait for decode")
    lap.report()
    err = capsys.readouterr().err
    assert err.startswith("[call_region chrT]") and "wait for decode" in err and "upload" in err
