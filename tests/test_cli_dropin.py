"""The drop-in claim end to end, on the CPU: the reference's own command line (`himut call`, `himut normcounts`,
`himut phase`; /root/reference/src, run unmodified through the I/O look-alikes of tests/shims) once with its own workers and once after
`himut_b200.patch.install()`, on the same three-contig BAM and VCF files — option parsing, the BAM pre-pass
(get_thresholds mirror on the real BAM file through the native decoder), chunking, the fork of the worker pool with a
lazily created context, the Manager-dict contract, natsort, header and writers all included.  The output files (VCF,
normcounts TSV, phased VCF) and himut.log / norm.log must be byte-identical.  Without a GPU the device calls of the drop-in workers are answered by the oracle
(tests/standin.py); what the device computes is pinned separately (tests/test_gpu_*.py).

Build container only: the reference is not on the GPU box."""
import gzip
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
import refshim
from himut_b200 import bamio, synth

pytestmark = pytest.mark.skipif(not refshim.have_reference(), reason="reference sources are not present")

HERE = os.path.dirname(os.path.abspath(__file__))
VCF_HEAD = "##fileformat=VCFv4.2\n%s#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tsynth\n"


def _write_sites(path, per_contig, compress):
    """per_contig: [(chrom, keys)] -> single-sample VCF of PASS SNVs; compress: BGZF + a real tabix index, which the
    drop-in's loaders use (the reference side goes through the pytabix look-alike, which scans the file)"""
    import tbi
    head = VCF_HEAD % ""
    rows = []
    for chrom, keys in per_contig:
        for k in keys.tolist():
            ref = "ATGC"[(k >> 2) & 3]
            rows.append((chrom, k >> 4, ref, "%s\t%d\t.\t%s\t%s\t.\tPASS\t.\tGT\t0/1" % (chrom, k >> 4, ref, "ATGC"[k & 3])))
    if compress:
        tbi.write_vcf_bgz_tbi(path, head, rows)
    else:
        with open(path, "w") as f:
            f.write(head + "".join(r[3] + "\n" for r in rows))


def _write_phased(path, data, phase_block):
    text = [VCF_HEAD % '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n##FORMAT=<ID=PS,Number=1,Type=Integer,Description="Phase set">\n']
    for chrom, _n, d in data:
        ph = synth.phase_table(d.germ, phase_block)
        for s in range(ph["set_off"].size - 1):
            a, b = int(ph["set_off"][s]), int(ph["set_off"][s + 1])
            ps = int(ph["hpos"][a])
            for i in range(a, b):
                bit = int(ph["hbit"][i])
                text.append("%s\t%d\t.\t%s\t%s\t.\tPASS\t.\tGT:PS\t%s:%d\n" % (
                    chrom, int(ph["hpos"][i]), "ATGC"[int(ph["href"][i])], "ATGC"[int(ph["halt"][i])], "0|1" if bit == 0 else "1|0", ps))
    with open(path, "w") as f:
        f.write("".join(text))


def _inputs(tmp, phase_block=None):
    data = cases.cli_dataset(phase_block)
    bam = os.path.join(tmp, "synth.bam")
    bamio.write_batches_bam(bam, [(c, n, d.batch) for c, n, d in data])
    sets = [(c,) + cases.site_sets_from_synth(d, 40 + i) for i, (c, _n, d) in enumerate(data)]
    return data, bam, sets


def _start(mode, tmp, phase_block, argv, out_name):
    work = os.path.join(tmp, mode)
    os.makedirs(work)
    out = os.path.join(work, out_name)
    cmd = [sys.executable, os.path.join(HERE, "cli_runner.py"), mode, work, str(phase_block or 0)] + argv + ["-o", out]
    env = dict(os.environ, PYTHONHASHSEED="0")
    return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env), work, out


def _finish(started, log_name):
    p, work, out = started
    try:
        stdout, _ = p.communicate(timeout=1500)
    except subprocess.TimeoutExpired:
        p.kill()
        raise
    assert p.returncode == 0, stdout[-4000:]
    assert os.path.exists(out), stdout[-4000:]
    strip = lambda text: [l for l in text.split("\n") if not l.startswith("##fileDate")]
    vcf = strip(open(out).read())
    log = open(os.path.join(work, log_name)).read()
    return vcf, log, stdout


def _run(mode, tmp, phase_block, argv, out_name="out.vcf", log_name="himut.log"):
    return _finish(_start(mode, tmp, phase_block, argv, out_name), log_name)


def _run_both(tmp, phase_block, argv, out_name="out.vcf", log_name="himut.log"):
    """the unmodified run and the drop-in run side by side -> ((vcf, log, stdout) of the reference, of the drop-in)"""
    a = _start("reference", tmp, phase_block, argv, out_name)
    b = _start("dropin", tmp, phase_block, argv, out_name)
    return _finish(a, log_name), _finish(b, log_name)


def _compare(tmp, phase_block, argv, expect_status, contigs=("chr1", "chr2", "chr10")):
    (ref_vcf, ref_log, ref_out), (our_vcf, our_log, our_out) = _run_both(tmp, phase_block, argv)
    # the header names the output path: the two runs write to different directories
    fix = lambda lines, mode: [l.replace(os.path.join(tmp, mode), "<work>") for l in lines]
    ref_vcf, our_vcf = fix(ref_vcf, "reference"), fix(our_vcf, "dropin")
    rows = [l for l in ref_vcf if l and not l.startswith("#")]
    assert len(rows) > 50, "the reference emitted almost nothing: not a test"
    for s in expect_status:
        assert any(("\t%s\t" % s) in l for l in rows), "no %s row in the reference's VCF" % s
    assert our_vcf == ref_vcf, next((a, b) for a, b in zip(our_vcf + [None], ref_vcf + [None]) if a != b)
    assert our_log == ref_log
    chroms = [l.split("\t")[0] for l in rows]
    first = [chroms.index(c) for c in contigs]
    assert first == sorted(first)  # natural order (chr1, chr2, chr10), not lexicographic
    return rows


def test_call_cli_is_a_drop_in(tmp_path):
    """`himut call -i … --common_snps … --panel_of_normals … -t 3`: human-sample defaults, bgzip-style site files"""
    tmp = str(tmp_path)
    data, bam, sets = _inputs(tmp)
    common, pon = os.path.join(tmp, "common.vcf.bgz"), os.path.join(tmp, "pon.vcf.bgz")
    _write_sites(common, [(c, k) for c, k, _ in sets], compress=True)
    _write_sites(pon, [(c, k) for c, _, k in sets], compress=True)
    _compare(tmp, None, ["call", "-i", bam, "--common_snps", common, "--panel_of_normals", pon, "-t", "3"],
             ["PASS", "ComSnp", "PanelOfNormal"])


def test_call_cli_phase_is_a_drop_in(tmp_path):
    """`himut call --phase --phased_vcf …` with plain-text site files (the `.vcf` common-SNP loader keeps the *other*
    contigs' records, vcflib.py:434 — the drop-in must do the same) and non-default thresholds"""
    tmp = str(tmp_path)
    block = 50_000
    data, bam, sets = _inputs(tmp, block)
    common, pon, phased = os.path.join(tmp, "common.vcf"), os.path.join(tmp, "pon.vcf"), os.path.join(tmp, "germline.phased.vcf")
    _write_sites(common, [(c, k) for c, k, _ in sets], compress=False)
    _write_sites(pon, [(c, k) for c, _, k in sets], compress=False)
    _write_phased(phased, data, block)
    regions = os.path.join(tmp, "regions.txt")
    open(regions, "w").write("chr2\nchr10\n")  # the reference needs 8 us per aligned base in --phase mode; all three contigs in
    _compare(tmp, block, ["call", "-i", bam, "--phase", "--phased_vcf", phased, "--common_snps", common, "--panel_of_normals", pon,
                          "--region_list", regions,   # --phase mode are tests/test_zz_pool.py against tests/golden/cli_call.json
                          "--min_gq", "15", "--min_bq", "60", "--min_trim", "0.02", "--mismatch_window_size", "30", "-t", "2"],
             ["PASS"], contigs=("chr2", "chr10"))


def test_normcounts_cli_is_a_drop_in(tmp_path):
    """`himut normcounts --bam … --ref … --sbs <the VCF of himut call> --region_list …`: thresholds read back from the call
    VCF's header, the callable-base workers, the reference tri-count workers (reflib.get_genome_tricounts' pool) and the
    reference's own normalisation arithmetic and writers on top"""
    tmp = str(tmp_path)
    data, bam, sets = _inputs(tmp)
    common, pon = os.path.join(tmp, "common.vcf.bgz"), os.path.join(tmp, "pon.vcf.bgz")
    _write_sites(common, [(c, k) for c, k, _ in sets], compress=True)
    _write_sites(pon, [(c, k) for c, _, k in sets], compress=True)
    fasta = os.path.join(tmp, "ref.fa")
    with open(fasta, "w") as f:
        for c, _n, d in data:
            seq = d.ref.decode()
            f.write(">%s\n" % c + "\n".join(seq[i:i + 60] for i in range(0, len(seq), 60)) + "\n")
    # the --sbs input: the VCF `himut call` writes (identical from either side, test_call_cli_is_a_drop_in)
    os.makedirs(os.path.join(tmp, "call"))
    call_vcf, _log, _out = _run("dropin", os.path.join(tmp, "call"), None,
                                ["call", "-i", bam, "--common_snps", common, "--panel_of_normals", pon, "-t", "3"])
    sbs = os.path.join(tmp, "call", "dropin", "out.vcf")
    regions = os.path.join(tmp, "regions.txt")
    open(regions, "w").write("chr2\nchr10\n")  # the two short contigs: the reference needs 7 us per aligned base here
    argv = ["normcounts", "--bam", bam, "--ref", fasta, "--sbs", sbs, "--common_snps", common, "--panel_of_normals", pon,
            "--region_list", regions, "-t", "2"]
    (ref_tsv, ref_log, _), (our_tsv, our_log, _) = _run_both(tmp, None, argv, "out.normcounts.tsv", "norm.log")
    fix = lambda lines, mode: [l.replace(os.path.join(tmp, mode), "<work>") for l in lines]
    assert fix(our_tsv, "dropin") == fix(ref_tsv, "reference")
    assert our_log == ref_log
    body = [l for l in ref_tsv if l and not l.startswith("#")]
    assert len(body) >= 96


def _write_germline(path, data):
    text = [VCF_HEAD % ""]
    for chrom, _n, d in data:
        g = d.germ
        for p, r, a, gt in zip(g["pos"].tolist(), g["ref"].tolist(), g["alt"].tolist(), g["gt"].tolist()):
            text.append("%s\t%d\t.\t%s\t%s\t50\tPASS\t.\tGT\t%s\n" % (chrom, p, "ATGC"[r], "ATGC"[a], "0/1" if gt < 2 else "1/1"))
        # a few germline indels: with none the reference's get_truncated_float(0.0) raises (util.py:539-544)
        for i, p in enumerate((1000, 2000, 3000, 4000)):
            text.append("%s\t%d\t.\tA\t%s\t50\tPASS\t.\tGT\t%s\n" % (chrom, p, "AT" if i % 2 else "ATT", "0/1" if i < 3 else "1/1"))
    open(path, "w").write("".join(text))


def _write_fasta(path, data):
    with open(path, "w") as f:
        for c, _n, d in data:
            seq = d.ref.decode()
            f.write(">%s\n" % c + "\n".join(seq[i:i + 60] for i in range(0, len(seq), 60)) + "\n")


def test_call_cli_non_human_sample_is_a_drop_in(tmp_path):
    """`himut call --non_human_sample --ref … --vcf …`: the germline priors come from the sample's own VCF (truncated
    floats, vcflib.get_germline_priors), no common-SNP / panel-of-normals sets are loaded, a --region_list restricts
    the run to two contigs"""
    tmp = str(tmp_path)
    data, bam, _sets = _inputs(tmp)
    germline, fasta, regions = os.path.join(tmp, "germline.vcf"), os.path.join(tmp, "ref.fa"), os.path.join(tmp, "regions.txt")
    _write_germline(germline, data)
    _write_fasta(fasta, data)
    open(regions, "w").write("chr10\nchr2\n")
    (ref_vcf, ref_log, _), (our_vcf, our_log, _) = _run_both(tmp, None, ["call", "-i", bam, "--non_human_sample", "--ref", fasta, "--vcf", germline,
                                                                        "--region_list", regions, "-t", "2"])
    fix = lambda lines, mode: [l.replace(os.path.join(tmp, mode), "<work>") for l in lines]
    assert fix(our_vcf, "dropin") == fix(ref_vcf, "reference")
    assert our_log == ref_log
    cmd = next(l for l in ref_vcf if l.startswith("##himut_command"))
    assert "--germline_snv_prior 0.001 " not in cmd  # a prior measured from the VCF, not the default
    assert sum(1 for l in ref_vcf if l and not l.startswith("#")) > 30


def test_call_cli_create_panel_of_normal_is_a_drop_in(tmp_path):
    """`himut call --create_panel_of_normal --region_list <chrom start end>`: the preset of util.load_pon_params replaces
    the thresholds; arbitrary, overlapping windows (the som_seen carry between them) through the whole CLI"""
    tmp = str(tmp_path)
    data, bam, _sets = _inputs(tmp)
    regions = os.path.join(tmp, "regions.txt")
    open(regions, "w").write("chr2\t1000\t60000\nchr2\t55000\t90000\nchr10\t0\t40000\n")  # windows, two of them overlapping
    argv = ["call", "-i", bam, "--create_panel_of_normal", "--region_list", regions, "-t", "2"]
    (ref_vcf, ref_log, _), (our_vcf, our_log, _) = _run_both(tmp, None, argv)
    fix = lambda lines, mode: [l.replace(os.path.join(tmp, mode), "<work>") for l in lines]
    assert fix(our_vcf, "dropin") == fix(ref_vcf, "reference")
    assert our_log == ref_log
    assert sum(1 for l in ref_vcf if l and not l.startswith("#")) > 30


def test_phase_cli_is_a_drop_in(tmp_path):
    """`himut phase --bam … --vcf <germline> -o x.phased.vcf`: the hetSNP pair tables come from the phase-edge mirror
    (phaselib.get_edges), the binomial tests, the graph search and the writer stay the reference's"""
    tmp = str(tmp_path)
    data, bam, _sets = _inputs(tmp)
    germline = os.path.join(tmp, "germline.vcf")
    text = [VCF_HEAD % ""]
    for chrom, _n, d in data:
        g = d.germ
        for p, r, a, gt in zip(g["pos"].tolist(), g["ref"].tolist(), g["alt"].tolist(), g["gt"].tolist()):
            text.append("%s\t%d\t.\t%s\t%s\t50\tPASS\t.\tGT\t%s\n" % (chrom, p, "ATGC"[r], "ATGC"[a], "0/1" if gt < 2 else "1/1"))
    open(germline, "w").write("".join(text))
    argv = ["phase", "--bam", bam, "--vcf", germline, "-t", "2"]
    (ref_vcf, _l, ref_out), (our_vcf, _l2, our_out) = _run_both(tmp, None, argv, "out.phased.vcf", "out.phased.vcf")
    fix = lambda lines, mode: [l.replace(os.path.join(tmp, mode), "<work>") for l in lines]
    assert fix(our_vcf, "dropin") == fix(ref_vcf, "reference")
    phased = [l for l in ref_vcf if l and not l.startswith("#")]
    assert len(phased) > 100 and any("0|1" in l for l in phased) and any("1|0" in l for l in phased)
