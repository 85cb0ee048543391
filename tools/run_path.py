#!/usr/bin/env python
"""Run one path (call | normcounts) a few times on a synthetic contig: the small driver used for
ncu captures (`ncu ... python tools/run_path.py normcounts --contig-mb 8`)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("path", choices=["call", "normcounts"])
ap.add_argument("--contig-mb", type=int, default=8)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--depth", type=float, default=30.0)
ap.add_argument("--with-seq", action="store_true", help="call: upload the 2-bit base stream too (the worker mirror does not)")
ap.add_argument("--compact", action="store_true", help="upload the qualities as bitmap + exceptions, as the workers do")
a = ap.parse_args()

import cases  # noqa: E402
from himut_b200 import gtmodel, lib, synth  # noqa: E402

n = a.contig_mb * 1_000_000
d = synth.generate(n, seed=5, copy=False, depth=a.depth)
params = gtmodel.make_params(**dict(gtmodel.DEFAULT_CALL_ARGS, md_threshold=max(gtmodel.DEFAULT_CALL_ARGS["md_threshold"], 4 * a.depth)))
chunks = d.batch.chunk_table(cases.chunkloci(0, n))
with lib.Context(0) as ctx:
    ctx.set_params(params)
    ctx.set_site_sets()
    batch = d.batch if (a.path == "normcounts" or a.with_seq) else d.batch.without_seq()
    if a.compact:
        from himut_b200 import bamdec
        ctx.upload_compact(batch, bamdec.compact_bq(d.batch))
    else:
        ctx.upload(batch)
    for _ in range(a.reps):
        if a.path == "call":
            rec, log = ctx.call_chunks(chunks, view=True)
        else:
            ccs, rt, log, ties = ctx.normcounts_chunks(d.ref, chunks)
        print(a.path, [int(v) for v in log][:6], ctx.last_kernel_times())
