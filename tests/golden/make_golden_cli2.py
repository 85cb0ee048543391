#!/usr/bin/env python
"""tests/golden/cli_norm_phase.json: what the reference's own `himut normcounts` command line and its own
phaselib.get_edges produce for the three-contig data set of cases.cli_dataset(), so that the drop-in workers of those two
sub-commands can be held against the reference on the GPU box (forked pool, CUDA library; tests/test_zz_pool.py), where
the reference does not exist.

  normcounts  `himut normcounts --bam .. --ref .. --sbs <call VCF> --common_snps .. --panel_of_normals .. --region_list`
              run unmodified through tests/cli_runner.py: per trinucleotide the ref_callable_tri_count and
              ccs_callable_tri_count columns of the TSV (mutlib.dump_normcounts, the workers' tallies summed over
              contigs), norm.log's per-contig vectors, the thresholds normcounts read back from the call VCF's header
  phase       himut.phaselib.get_edges (the starmap target inside `himut phase`, phaselib.py:236-248) per contig with
              the sub-command's defaults (min_bq 20, min_mapq 20): every edge with its four counts

    python tests/golden/make_golden_cli2.py         (build container only, a few minutes)
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
import test_cli_dropin as t  # noqa: E402

NORM_CONTIGS = ["chr2", "chr10"]  # the reference needs 7 us per aligned base here


def parse_norm_log(text):
    lines = [l.split() for l in text.strip().split("\n")]
    chroms = lines[0][:-1]
    return {c: [int(float(row[1 + i])) for row in lines[1:]] for i, c in enumerate(chroms)}


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        data, bam, sets = t._inputs(tmp)
        common, pon = os.path.join(tmp, "common.vcf.bgz"), os.path.join(tmp, "pon.vcf.bgz")
        t._write_sites(common, [(c, k) for c, k, _ in sets], compress=True)
        t._write_sites(pon, [(c, k) for c, _, k in sets], compress=True)
        fasta = os.path.join(tmp, "ref.fa")
        t._write_fasta(fasta, data)
        os.makedirs(os.path.join(tmp, "call"))
        call_vcf, _log, _o = t._run("reference", os.path.join(tmp, "call"), None,
                                    ["call", "-i", bam, "--common_snps", common, "--panel_of_normals", pon, "-t", "3"])
        sbs = os.path.join(tmp, "call", "reference", "out.vcf")
        regions = os.path.join(tmp, "regions.txt")
        open(regions, "w").write("".join(c + "\n" for c in NORM_CONTIGS))
        tsv, log, _o = t._run("reference", tmp, None, ["normcounts", "--bam", bam, "--ref", fasta, "--sbs", sbs, "--common_snps", common,
                                                      "--panel_of_normals", pon, "--region_list", regions, "-t", "2"],
                              "out.normcounts.tsv", "norm.log")
        rows = [l.split("\t") for l in tsv if l and not l.startswith("#") and not l.startswith("sub\t") and len(l.split("\t")) == 10]
        tri = {}
        for r in rows:
            tri[r[1]] = [int(float(r[8])), int(float(r[9]))]  # ref_callable_tri_count, ccs_callable_tri_count
        import re
        cmd = next(l for l in call_vcf if l.startswith("##himut_command"))
        get = lambda k: int(float(re.search(r"--%s (\S+)" % k, cmd).group(1)))
        depth = next(l for l in call_vcf if l.startswith("##FILTER=<ID=HighDepth"))
        md = float(depth.strip().split()[-1].replace('">', ""))
        out["normcounts"] = {"contigs": NORM_CONTIGS, "tri": tri, "log": parse_norm_log(log),
                             "thresholds": [get("qlen_lower_limit"), get("qlen_upper_limit"), md]}
        print("normcounts:", len(tri), "trinucleotides", {c: v[-1] for c, v in out["normcounts"]["log"].items()})
    # phase edges: the reference's get_edges per contig
    import scipy.stats
    if not hasattr(scipy.stats, "binom_test"):
        scipy.stats.binom_test = lambda k, n, p=0.5, alternative="two-sided": scipy.stats.binomtest(int(k), int(n), p, alternative=alternative).pvalue
    import refshim
    refshim.import_reference()
    import himut.phaselib
    import pysam
    edges = {}
    for c, n, d in cases.cli_dataset():
        het = d.germ["gt"] < 2
        hetsnp_lst = [(int(p), "ATGC"[r], "ATGC"[a]) for p, r, a in zip(d.germ["pos"][het], d.germ["ref"][het], d.germ["alt"][het])]
        h2i = {h: i for i, h in enumerate(hetsnp_lst)}
        pysam.register("cli_%s.bam" % c, refshim.BatchProvider(c, n, d.batch))
        edge_lst, e2c = himut.phaselib.get_edges(c, "cli_%s.bam" % c, 20, 20, [h[0] for h in hetsnp_lst], hetsnp_lst, h2i)
        edges[c] = [[int(i), int(j)] + [int(v) for v in e2c[(i, j)]] for (i, j) in edge_lst]
        print("phase edges:", c, len(hetsnp_lst), "hetSNPs", len(edge_lst), "edges")
    out["phase_edges"] = edges
    with open(os.path.join(HERE, "cli_norm_phase.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))


if __name__ == "__main__":
    main()
