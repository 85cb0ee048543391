#!/usr/bin/env python
"""BAM (page cache) -> records through the `call` worker mirror with HIMUT_B200_WORKER_TIMING=1: where a contig's host
time goes (decode wait, page-locking, upload, submit / collect).  python tools/worker_trace.py [--contig-mb 64]"""
import argparse
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["HIMUT_B200_WORKER_TIMING"] = "1"
ap = argparse.ArgumentParser()
ap.add_argument("--contig-mb", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()

import cases  # noqa: E402
from himut_b200 import bamdec, caller, gtmodel, synth  # noqa: E402

n = a.contig_mb * 1_000_000
d = synth.generate(n, seed=5, copy=False)
tmp = tempfile.mkdtemp()
bam = os.path.join(tmp, "c.bam")
bamdec.write_batch_bam(bam, "chr1", n, d.batch)
g = gtmodel.DEFAULT_CALL_ARGS
loci = [("chr1", s, e) for s, e in cases.chunkloci(0, n)]
nb = bamdec.NativeBam(bam)
t0 = time.perf_counter()
nb.read_batch("chr1", 0, n, copy=False, seq=False, compact=True)
print("whole-contig decode %.3f s (%d threads)" % (time.perf_counter() - t0, bamdec.default_threads()), file=sys.stderr)
nb.close()
for _ in range(a.reps):
    lst, log = {}, {}
    t0 = time.perf_counter()
    caller.get_somatic_substitutions(
        "chr1", bam, None, None, loci, {}, {}, {}, g["min_qv"], g["min_mapq"], g["qlen_lower_limit"], g["qlen_upper_limit"],
        g["min_sequence_identity"], g["min_gq"], g["min_bq"], g["min_trim"], g["max_mismatch_count"], g["mismatch_window"],
        g["md_threshold"], g["min_ref_count"], g["min_alt_count"], g["min_hap_count"], 1e-6, g["germline_snv_prior"], 1e-4,
        False, True, False, lst, log)
    print("worker %.3f s, %d rows" % (time.perf_counter() - t0, len(lst["chr1"])), file=sys.stderr)
