#!/usr/bin/env python
"""Fuzzing session, build container only: the reference's phaselib.get_edges against the oracle's phase-edge tables on
adversarial batches (tests/cases.py:adversarial_batch) with random hetSNP lists and thresholds.
    python tools/fuzz_edges_vs_reference.py 0 150      # seeds; prints mismatches, exits 0 when there is none
(760 seeds on 2026-10-18: no mismatch.)"""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import numpy as np
import cases, refshim
from oracle import oracle
import scipy.stats
if not hasattr(scipy.stats, "binom_test"):
    scipy.stats.binom_test = lambda k, n, p=0.5, alternative="two-sided": scipy.stats.binomtest(int(k), int(n), p, alternative=alternative).pvalue
refshim.import_reference()
import himut.phaselib, pysam
B = {"A":0,"T":1,"G":2,"C":3}
bad = 0
lo, hi = int(sys.argv[1]), int(sys.argv[2])
for seed in range(lo, hi):
    rnd = random.Random(seed * 7 + 1)
    n = rnd.choice([1500, 3000])
    batch, ref = cases.adversarial_batch(seed, contig_len=n, n_reads=rnd.choice([40, 120, 300]), max_len=rnd.choice([300, 900, 2000]))
    pos = sorted(rnd.sample(range(1, n + 1), rnd.choice([5, 30, 120])))
    hets = [(p, ref[p-1], rnd.choice([b for b in "ATGC" if b != ref[p-1]])) for p in pos if ref[p-1] in "ATGC"]
    if len(hets) < 2: continue
    min_bq, min_mapq = rnd.choice([0, 1, 20, 93]), rnd.choice([0, 20, 60])
    pysam.register("e.bam", refshim.BatchProvider("chr1", n, batch))
    try:
        edge_lst, e2c = himut.phaselib.get_edges("chr1", "e.bam", min_bq, min_mapq, [h[0] for h in hets], hets, {h: i for i, h in enumerate(hets)})
    except Exception as ex:
        print(seed, "reference raised", repr(ex)); continue
    exp = sorted([int(i), int(j)] + [int(v) for v in e2c[(i, j)]] for (i, j) in edge_lst)
    hpos = np.array([h[0] for h in hets], np.int32); href = np.array([B[h[1]] for h in hets], np.uint8)
    table, need = oracle.phase_edges(batch, hpos, href, 256, min_bq, min_mapq)
    a, d = np.nonzero(table.any(axis=2))
    got = sorted([int(x), int(x + y + 1)] + [int(v) for v in table[x, y]] for x, y in zip(a.tolist(), d.tolist()))
    ok = need == 0 and got == exp
    if not ok:
        bad += 1
        print(seed, "MISMATCH need", need, len(got), len(exp), [g for g in got if g not in exp][:3], [e for e in exp if e not in got][:3])
print("seeds", lo, hi, "bad", bad)
sys.exit(1 if bad else 0)
