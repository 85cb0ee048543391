#!/usr/bin/env python
"""Generates tests/golden/*.json by running the UNMODIFIED reference workers
(/root/reference/src/himut, through tests/shims) on the deterministic cases of tests/cases.py.

Run in the build container only (the reference is not on the GPU box):
    PYTHONHASHSEED=0 python tests/golden/make_golden.py [case ...]
PYTHONHASHSEED is pinned because normcounts picks its alt allele by iterating a set of
one-character strings (src/himut/normcounts.py:370); the order it saw is stored in the fixture.
"""
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

if os.environ.get("PYTHONHASHSEED") != "0":
    os.environ["PYTHONHASHSEED"] = "0"
    os.execv(sys.executable, [sys.executable] + sys.argv)

import numpy as np  # noqa: E402

import cases  # noqa: E402
import refshim  # noqa: E402
from himut_b200 import abi  # noqa: E402


def _plain(v):
    if isinstance(v, (np.floating, float)):
        return float(v)
    if isinstance(v, (np.integer, int)):
        return int(v)
    return v


write_sites_vcf = cases.write_sites_vcf
phase_dicts = cases.phase_dicts


def reference_outputs(c, himut, tmp):
    """run the reference worker of case dict `c` (cases.build_case / cases.random_case) -> (expected dict, seconds)"""
    import pysam
    name = c["name"]
    a = c["args"]
    bam = os.path.join(tmp, name + ".bam")
    pysam.register(bam, refshim.BatchProvider(cases.CHROM, c["contig_len"], c["batch"]))
    common = pon = None
    if c["common_vcf"].size:
        common = os.path.join(tmp, name + ".common.vcf.bgz")  # served by the tabix shim
        write_sites_vcf(common, cases.CHROM, c["common_vcf"])
    if c["pon_vcf"].size:
        pon = os.path.join(tmp, name + ".pon.vcf.bgz")
        write_sites_vcf(pon, cases.CHROM, c["pon_vcf"])
    chunkloci = [(cases.CHROM, s, e) for s, e in c["chunks"]]
    hbit, hpos, hetsnp = phase_dicts(c)
    phase = bool(a.get("phase", False))
    non_human = bool(a.get("non_human_sample", False))
    create_pon = bool(a.get("create_panel_of_normals", False))
    t0 = time.time()
    exp = {}
    if c["kind"] == "call":
        lst, log = {}, {}
        himut.caller.get_somatic_substitutions(
            cases.CHROM, bam, common, pon, chunkloci, hbit, hpos, hetsnp,
            a["min_qv"], a["min_mapq"], a["qlen_lower_limit"], a["qlen_upper_limit"],
            a["min_sequence_identity"], a["min_gq"], a["min_bq"], a["min_trim"],
            a["max_mismatch_count"], a["mismatch_window"], a["md_threshold"], a["min_ref_count"],
            a["min_alt_count"], a["min_hap_count"], 1e-6, a["germline_snv_prior"], 1e-4,
            phase, non_human, create_pon, lst, log)
        exp["tsbs_lst"] = [[_plain(v) for v in row] for row in lst[cases.CHROM]]
        exp["log"] = [int(v) for v in log[cases.CHROM]]
    else:
        ccs, rt, log = {}, {}, {}
        himut.normcounts.get_callable_tricounts(
            cases.CHROM, c["ref"], bam, common, pon, chunkloci, hbit, hpos, hetsnp,
            a["min_qv"], a["min_mapq"], a["min_trim"], a["qlen_lower_limit"], a["qlen_upper_limit"],
            a["min_sequence_identity"], a["min_gq"], a["min_bq"], a["mismatch_window"],
            a["max_mismatch_count"], a["min_ref_count"], a["min_alt_count"], a["min_hap_count"],
            float(a["md_threshold"]), 1e-6, a["germline_snv_prior"], 1e-4, phase, non_human,
            ccs, rt, log)
        exp["ccs_tri2count"] = {k: int(v) for k, v in ccs[cases.CHROM].items()}
        exp["ref_tri2count"] = {k: int(v) for k, v in rt[cases.CHROM].items()}
        exp["log"] = [int(v) for v in log[cases.CHROM]]
        exp["alt_order"] = [[abi.BASE2CODE[x] for x in list(himut.util.base_set.difference(r))] for r in "ATGC"]
    return exp, time.time() - t0


def run_case(name, himut, tmp):
    c = cases.build_case(name)
    exp, dt = reference_outputs(c, himut, tmp)
    fixture = dict(case=name, kind=c["kind"], batch_sha256=cases.batch_digest(c["batch"]),
                   n_reads=c["batch"].n_reads, aligned_bases=c["batch"].aligned_bases,
                   reference_seconds=round(dt, 2), expected=exp)
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(fixture, f, indent=0, separators=(",", ":"))
    print("%-22s reads=%d bases=%d ref=%.1fs (%.3g bases/s) log=%s" % (
        name, c["batch"].n_reads, c["batch"].aligned_bases, dt, c["batch"].aligned_bases / max(dt, 1e-9), exp["log"]))


def main():
    himut = refshim.import_reference()
    names = sys.argv[1:] or list(cases.CASES)
    with tempfile.TemporaryDirectory() as tmp:
        for n in names:
            run_case(n, himut, tmp)


if __name__ == "__main__":
    main()
