#!/usr/bin/env python
"""tests/golden/thresholds.json: the reference's own bamlib.get_thresholds
(/root/reference/src/himut/bamlib.py:137-178) on deterministic synthetic contigs, served to it
through the pysam shim.  Build container only."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refshim  # noqa: E402

CASES = {"two_contigs": [("chr1", 400_000, 41), ("chr2", 250_000, 42)], "one_contig": [("chrA", 300_000, 43)]}


def batches(name):
    from himut_b200 import synth
    return [(chrom, n, synth.generate(n, seed=seed).batch) for chrom, n, seed in CASES[name]]


class MultiProvider:
    def __init__(self, items):
        self.by_chrom = {chrom: refshim.BatchProvider(chrom, n, b) for chrom, n, b in items}
        self.header_text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (c, n) for c, n, _ in items) + "@RG\tID:rg\tSM:synth\n"

    def fetch_records(self, chrom=None, start=None, end=None):
        return self.by_chrom[chrom].fetch_records(chrom, start, end)


def main():
    himut = refshim.import_reference()
    import himut.bamlib
    import pysam
    exp = {}
    for name in CASES:
        items = batches(name)
        pysam.register("golden.bam", MultiProvider(items))
        chrom_lst = [c for c, _, _ in items]
        chrom2len = {c: n for c, n, _ in items}
        exp[name] = [int(v) for v in himut.bamlib.get_thresholds("golden.bam", chrom_lst, chrom2len)]
        print(name, exp[name])
    with open(os.path.join(HERE, "thresholds.json"), "w") as f:
        json.dump({"expected": exp}, f)


if __name__ == "__main__":
    main()
