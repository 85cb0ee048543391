#!/usr/bin/env python
"""Where a resident `himut call` step goes when two contexts alternate (bench.py's headline loop): host wall time of
submit / collect per step, the CUDA-event time of every kernel group of every call, and the whole loop by wall clock
and by events.  Diagnostics only: python tools/step_trace.py [--contig-mb 64] [--steps 12] [--mode alt|sync|sync_wait]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--contig-mb", type=int, default=64)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--mode", default="alt,sync,sync_wait", help="alt | sync | sync_wait | e2e (alternating pair fed from page-locked host buffers)")
ap.add_argument("--no-ktiming", action="store_true")
ap.add_argument("--own-streams", action="store_true", help="the contexts keep the streams the library created for them")
ap.add_argument("--start", type=int, default=0, help="index of the context that takes the first step")
a = ap.parse_args()

import numpy as np  # noqa: E402
import torch  # noqa: E402
import cases  # noqa: E402
from himut_b200 import gtmodel, lib, synth  # noqa: E402

n = a.contig_mb * 1_000_000
d = synth.generate(n, seed=5, copy=False)
params = gtmodel.make_params(**gtmodel.DEFAULT_CALL_ARGS)
chunks = d.batch.chunk_table(cases.chunkloci(0, n))
batch = d.batch.without_seq()
torch.cuda.set_device(0)
ctxs = [lib.Context(0), lib.Context(0)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for c, st in zip(ctxs, streams):
    if not a.own_streams:
        c.set_stream(st.cuda_stream)
    c.set_params(params)
    c.set_site_sets()
    c.omit_restatements(True)
    c.kernel_timing(not a.no_ktiming)
    c.upload(batch)


cq = small = None
if "e2e" in a.mode:
    from himut_b200 import bamdec
    cq = bamdec.compact_bq(d.batch)
    small = [getattr(batch, n_) for n_, _ in batch._FIELDS if n_ not in ("seq", "bq", "ops", "seq_off")]
    ctxs[0].pin_arrays([cq.mask, cq.exc, cq.exc_off, batch.ops] + small)


def fmt(kt):
    return " ".join("%s %.3f" % (k.replace("k_", ""), v) for k, v in kt)


for mode in a.mode.split(","):
    state = {"k": a.start, "pending": None}
    rows, ups = [], []
    marks, w_ref = None, [0.0]

    def step():
        t0 = time.perf_counter()
        if mode in ("alt", "e2e"):
            c = ctxs[state["k"] % 2]
            state["k"] += 1
            if mode == "e2e":
                c.upload_compact(batch, cq, wait=False)  # as the worker: enqueue, then wait for the previous upload
                if state["pending"] is not None:
                    state["pending"].upload_wait()
                tu = time.perf_counter()
                ups.append(1e3 * (tu - t0))
                t0 = tu
                if marks is not None:  # device-time marks on the context's stream: behind k_bq_expand, behind the call's kernels
                    st = streams[ctxs.index(c)]
                    ea = torch.cuda.Event(enable_timing=True); ea.record(st)
            c.call_chunks_submit(chunks)
            if mode == "e2e" and marks is not None:
                eb = torch.cuda.Event(enable_timing=True); eb.record(st)
                marks.append((1e3 * (tu - w_ref[0]), ea, eb))
            t1 = time.perf_counter()
            kt = None
            if state["pending"] is not None:
                state["pending"].call_chunks_collect(view=True)
                kt = state["pending"].last_kernel_times()
            state["pending"] = c
        else:
            ctxs[0].call_chunks(chunks, view=True, wait=(mode == "sync_wait"))
            t1 = time.perf_counter()
            kt = ctxs[0].last_kernel_times()
        t2 = time.perf_counter()
        rows.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), kt))

    def drain():
        if state["pending"] is not None:
            state["pending"].upload_wait()
            state["pending"].call_chunks_collect(view=True)
            state["pending"] = None
        for c in ctxs:
            c.records_wait()

    for _ in range(8):
        step()
    drain()
    rows.clear()
    torch.cuda.synchronize()
    marks = [] if (mode == "e2e" and not a.own_streams) else None
    ref = torch.cuda.Event(enable_timing=True)
    ref.record(streams[0])
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    w_ref = [w0]
    for _ in range(a.steps):
        step()
    drain()
    torch.cuda.synchronize()
    wall = 1e3 * (time.perf_counter() - w0)
    print("== mode %s: %.3f ms per step by wall clock (%d steps)" % (mode, wall / a.steps, a.steps))
    for i, (s, c, kt) in enumerate(rows):
        up = ("upload %.3f " % ups[len(ups) - len(rows) + i]) if ups else ""
        print("  step %2d  host %ssubmit %.3f collect %.3f ms | %s" % (i, up, s, c, fmt(kt) if kt else ""))
    if marks:
        for i, (host_ms, ea, eb) in enumerate(marks):
            print("  call %2d  upload returned at host %.3f ms | device: expand done %.3f, call done %.3f ms after the loop began"
                  % (i, host_ms, ref.elapsed_time(ea), ref.elapsed_time(eb)))
for c in ctxs:
    c.close()
