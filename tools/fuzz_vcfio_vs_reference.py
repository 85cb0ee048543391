#!/usr/bin/env python
"""Fuzzing session, build container only: the site-file loaders of himut_b200.vcfio against the reference's own
(vcflib.load_common_snp / load_bgz_common_snp / load_pon / load_bgz_pon, through the pytabix look-alike) on random VCF
records: multi-allelic, non-PASS, indels, MNVs, symbolic and N alleles, several contigs; plain `.vcf` (with the
reference's `chrom != arr[0]` quirk) and `.vcf.bgz` + real tabix index.
    python tools/fuzz_vcfio_vs_reference.py 0 200
(400 seeds on 2026-10-18: no mismatch.)"""
import os
import random
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import refshim  # noqa: E402
import tbi  # noqa: E402
from himut_b200 import abi, vcfio  # noqa: E402

HEAD = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS\n"


def keys_of(sbs_set):
    out = []
    for pos, ref, alt in sbs_set:
        if ref in abi.BASE2CODE and alt in abi.BASE2CODE:
            out.append((int(pos) << 4) | (abi.BASE2CODE[ref] << 2) | abi.BASE2CODE[alt])
    return np.unique(np.array(sorted(out), np.uint64)) if out else np.zeros(0, np.uint64)


def main():
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    himut = refshim.import_reference()
    import himut.vcflib as ref
    bad = 0
    alleles = ["A", "T", "G", "C", "N", "AT", "GCA", "<DEL>", "*", "a"]
    with tempfile.TemporaryDirectory() as tmp:
        for seed in range(lo, hi):
            rnd = random.Random(seed)
            contigs = ["chr1", "chr2", "chrX"][: rnd.choice([1, 2, 3])]
            rows = []
            for c in contigs:
                n = rnd.choice([2000, 40000])
                for pos in sorted(rnd.choices(range(1, n), k=rnd.choice([5, 60, 400]))):
                    r = rnd.choice(alleles[:7])
                    a = ",".join(rnd.choice(alleles) for _ in range(rnd.choice([1, 1, 1, 2, 3])))
                    flt = rnd.choice(["PASS", "PASS", "PASS", "q10", "."])
                    rows.append((c, pos, r, "%s\t%d\t.\t%s\t%s\t%s\t%s\t.\tGT\t%s" % (c, pos, r, a, rnd.choice([".", "30"]), flt, rnd.choice(["0/1", "1/1", "0|1"]))))
            plain, bgz = os.path.join(tmp, "s%d.vcf" % seed), os.path.join(tmp, "s%d.vcf.bgz" % seed)
            open(plain, "w").write(HEAD + "".join(r[3] + "\n" for r in rows))
            tbi.write_vcf_bgz_tbi(bgz, HEAD, rows, block_bytes=rnd.choice([700, 3000, 60000]))
            for c in contigs + ["chr9"]:
                checks = [("common .vcf", keys_of(ref.load_common_snp(c, plain)), vcfio.load_common_snps(c, plain)),
                          ("pon .vcf", keys_of(ref.load_pon(c, plain)), vcfio.load_pon(c, plain))]
                if c != "chr9":  # the reference's tabix query raises for a contig the file does not name
                    checks += [("common .bgz", keys_of(ref.load_bgz_common_snp((c, 0, 50000), bgz)), vcfio.load_common_snps(c, bgz)),
                               ("pon .bgz", keys_of(ref.load_bgz_pon((c, 0, 50000), bgz)), vcfio.load_pon(c, bgz))]
                else:
                    assert vcfio.load_common_snps(c, bgz).size == 0 and vcfio.load_pon(c, bgz).size == 0
                for what, want, got in checks:
                    if not np.array_equal(want, got):
                        bad += 1
                        print("seed", seed, c, what, want.size, got.size)
            for p in (plain, bgz, bgz + ".tbi"):
                os.unlink(p)
    print("seeds", lo, hi, "mismatches", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
