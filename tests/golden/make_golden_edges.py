#!/usr/bin/env python
"""tests/golden/edges.json: the reference's own phaselib.get_edges (/root/reference/src/himut/phaselib.py:16-67)
on deterministic synthetic contigs, served through the pysam shim.  Build container only."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refshim  # noqa: E402

# name -> (contig_len, seed, min_bq, min_mapq, synth overrides)
CASES = {
    "plain": (120_000, 51, 20, 20, {}),
    "bq0_counts_deletions": (60_000, 52, 0, 0, {"indel_rate": 2e-3, "het_rate": 4e-3}),
    "strict": (80_000, 53, 93, 60, {"het_rate": 3e-3}),
}


def inputs(name):
    """-> (batch, hetsnp_lst [(pos, ref, alt)], hetsnp2hidx, contig_len, min_bq, min_mapq)"""
    from himut_b200 import synth
    n, seed, min_bq, min_mapq, over = CASES[name]
    d = synth.generate(n, seed=seed, **over)
    het = d.germ["gt"] < 2
    hetsnp_lst = [(int(p), "ATGC"[r], "ATGC"[a]) for p, r, a in zip(d.germ["pos"][het], d.germ["ref"][het], d.germ["alt"][het])]
    return d.batch, hetsnp_lst, {h: i for i, h in enumerate(hetsnp_lst)}, n, min_bq, min_mapq


def main():
    import scipy.stats
    if not hasattr(scipy.stats, "binom_test"):  # removed in scipy 1.12; phaselib imports it at module level
        scipy.stats.binom_test = lambda k, n, p=0.5, alternative="two-sided": scipy.stats.binomtest(int(k), int(n), p, alternative=alternative).pvalue
    refshim.import_reference()
    import himut.phaselib
    import pysam
    exp = {}
    for name in CASES:
        batch, hetsnp_lst, h2i, n, min_bq, min_mapq = inputs(name)
        pysam.register("edges.bam", refshim.BatchProvider("chr1", n, batch))
        edge_lst, e2c = himut.phaselib.get_edges("chr1", "edges.bam", min_bq, min_mapq, [h[0] for h in hetsnp_lst], hetsnp_lst, h2i)
        exp[name] = [[int(i), int(j)] + [int(v) for v in e2c[(i, j)]] for (i, j) in edge_lst]
        print(name, len(hetsnp_lst), "hetSNPs", len(edge_lst), "edges")
    with open(os.path.join(HERE, "edges.json"), "w") as f:
        json.dump({"expected": exp}, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
