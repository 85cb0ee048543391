#!/usr/bin/env python
"""Fuzzing session on a GPU box: CUDA against the CPU oracle on the random dirty batches of tests/cases.py:random_case
for seeds the committed sweep does not hold — `call` (plain upload with and without the base stream, compact upload,
forced 64-slot entry tables) and `normcounts` (plain and compact uploads), every record field and every counter.
python tools/fuzz_gpu_vs_oracle.py [--first 20000] [--count 300]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--first", type=int, default=20000)
ap.add_argument("--count", type=int, default=300)
a = ap.parse_args()

import numpy as np  # noqa: E402
import cases  # noqa: E402
import parity  # noqa: E402
from himut_b200 import bamdec, lib  # noqa: E402
from oracle import oracle  # noqa: E402

bad = done = 0
with lib.Context(0) as ctx:
    for seed in range(a.first, a.first + a.count):
        s = seed  # seeds >= cases.LONG_SEED_BASE draw long reads (several 2048-position tiles), smaller ones short reads
        for fam in ("call", "norm"):
            c = cases.random_case(fam, s)
            if c is None:
                continue
            batch, p, table, common, pon, phase = c["batch"], c["params"], c["chunk_table"], c["common"], c["pon"], c["phase"]
            ctx.set_params(p)
            ctx.set_site_sets(common, pon)
            if phase is not None:
                ctx.set_phase_sets(phase)
            cq = bamdec.compact_bq(batch)
            if fam == "call":
                o_rec, o_log = oracle.call_chunks(p, batch, table, common, pon, phase)
                forms = [("plain", lambda: ctx.upload(batch)), ("no seq", lambda: ctx.upload(batch.without_seq())),
                         ("compact", lambda: ctx.upload_compact(batch.without_seq(), cq))]
                for slots in (None, "64"):
                    if slots:
                        os.environ["HIMUT_B200_SITE_SLOTS"] = slots
                    else:
                        os.environ.pop("HIMUT_B200_SITE_SLOTS", None)
                    for name, up in (forms if slots is None else forms[2:]):
                        up()
                        rec, log = ctx.call_chunks(table)
                        ok, why = parity.records_equal(rec, o_rec)
                        done += 1
                        if not ok or list(log) != list(o_log):
                            bad += 1
                            print("MISMATCH call seed %d (%s, slots %s): %s" % (s, name, slots, why or "log %r != %r" % (list(log), list(o_log))))
                os.environ.pop("HIMUT_B200_SITE_SLOTS", None)
            else:
                ref = c["ref"].encode()
                o = oracle.normcounts_chunks(p, batch, ref, table, common, pon, phase)
                for name, up in (("plain", lambda: ctx.upload(batch)), ("compact", lambda: ctx.upload_compact(batch, cq))):
                    up()
                    g = ctx.normcounts_chunks(ref, table)
                    done += 1
                    if not (np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]) and list(g[2]) == list(o[2]) and g[3] == o[3]):
                        bad += 1
                        print("MISMATCH normcounts seed %d (%s)" % (s, name))
print("%d comparisons over %d seeds from %d, %d mismatches" % (done, a.count, a.first, bad))
sys.exit(1 if bad else 0)
