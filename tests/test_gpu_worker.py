"""The drop-in workers (himut_b200.caller / himut_b200.normcounts: the reference's signatures, a
BAM path in, the reference's dict entries out) against the reference's own outputs.  GPU."""
import os

import numpy as np
import pytest

import cases
import parity
from himut_b200 import bamio, caller, normcounts, worker

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("name", ["call_basic", "call_sets", "call_pon_params", "call_phase", "call_phase_indel", "call_adversarial_b"])
def test_call_worker_matches_reference(name, tmp_path):
    c = cases.build_case(name)
    fx = parity.load_golden(name)
    rows, log = _run_call(c, tmp_path)
    gold = parity.golden_rows(fx)
    assert parity.rows_equal(rows, gold), parity.first_diff(rows, gold)
    assert log == fx["expected"]["log"]


def test_call_worker_group_carry(tmp_path, monkeypatch):
    """a contig fed to the device in several batches gives the same rows (som_seen and the
    distinct-read count carry across batches)"""
    monkeypatch.setattr(worker, "GROUP_SPAN", 1200)
    for name in ("call_adversarial_a", "call_adversarial_b"):
        c = cases.build_case(name)
        fx = parity.load_golden(name)
        rows, log = _run_call(c, tmp_path)
        gold = parity.golden_rows(fx)
        assert parity.rows_equal(rows, gold), parity.first_diff(rows, gold)
        assert log == fx["expected"]["log"]


@pytest.mark.parametrize("name", ["norm_basic", "norm_phase", "norm_adversarial"])
def test_normcounts_worker_matches_reference(name, tmp_path, monkeypatch):
    if name == "norm_adversarial":
        monkeypatch.setattr(worker, "GROUP_SPAN", 1200)
    c = cases.build_case(name)
    fx = parity.load_golden(name)
    e = fx["expected"]
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
    for tri in normcounts.TRI_LST:  # the keys mutlib.get_cumsum_tricounts reads
        assert ccs[cases.CHROM][tri] == e["ccs_tri2count"].get(tri, 0), tri
        assert rt[cases.CHROM][tri] == e["ref_tri2count"].get(tri, 0), tri
    assert log[cases.CHROM] == e["log"]
