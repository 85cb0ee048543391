#!/usr/bin/env python
"""Runs the UNMODIFIED reference workers on the randomised sweep of tests/cases.py (the very cases
tests/test_gpu_random.py feeds to the CUDA path) and writes tests/golden/random_sweep.json:
per seed the batch digest, the reference's log vector and either a SHA-256 of its row list
(`call`) or its two tri-count tables (`normcounts`).

Exact PL ties: the reference's pinned numpy (1.24.4, poetry.lock) sorts a 10-vector by insertion sort, i.e. stably, so
the lowest gt_lst index wins a tie; this container's numpy 2.3 dispatches np.argsort to an unstable SIMD sort
(SURVEY.md A.7).  The sweep is full of depth-1 sites, where het and hom-alt tie exactly, so the reference is run with
gtlib's np.argsort made stable — the reference's behaviour under its own dependency pin — unless --native-argsort.

Build container only:
    python tests/golden/make_golden_random.py            # the committed sweep
    python tests/golden/make_golden_random.py --extra 3000 3200   # more seeds, compared with the oracle on the spot,
                                                                  # nothing written (a fuzzing session)
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden  # noqa: E402  (re-executes with PYTHONHASHSEED=0, sets sys.path)

import numpy as np  # noqa: E402

import cases  # noqa: E402
import parity  # noqa: E402
import refshim  # noqa: E402
from himut_b200 import records  # noqa: E402
from oracle import oracle  # noqa: E402


rows_digest = parity.rows_digest


def oracle_outputs(c, alt_order=None):
    if c["kind"] == "call":
        rec, log = oracle.call_chunks(c["params"], c["batch"], c["chunk_table"], c["common"], c["pon"], c["phase"])
        rows = records.records_to_tsbs_lst(cases.CHROM, rec)
        return dict(rows=rows, log=[int(v) for v in log])
    ccs, rt, log, ties = oracle.normcounts_chunks(c["params"], c["batch"], c["ref"].encode(), c["chunk_table"], c["common"],
                                                  c["pon"], c["phase"], alt_order=alt_order)
    return dict(ccs=ccs, ref=rt, log=[int(v) for v in log], ties=int(ties))


def one(kind, seed, himut, tmp):
    """-> (fixture entry or None, mismatch description or None)"""
    c = cases.random_case(kind, seed)
    if c is None:
        return None, None
    exp, dt = make_golden.reference_outputs(c, himut, tmp)
    entry = dict(batch_sha256=cases.batch_digest(c["batch"]), log=exp["log"], reference_seconds=round(dt, 2))
    why = None
    if kind == "call":
        ref_rows = [tuple(r) for r in exp["tsbs_lst"]]
        entry["n_rows"] = len(ref_rows)
        entry["rows_sha256"] = rows_digest(ref_rows)
        entry["statuses"] = sorted({r[4] for r in ref_rows})
        o = oracle_outputs(c)
        if not parity.rows_equal(o["rows"], ref_rows):
            why = "rows: " + parity.first_diff(o["rows"], ref_rows)
        elif rows_digest(o["rows"]) != entry["rows_sha256"]:
            why = "digest differs although rows compare equal"
    else:
        entry["ccs_tri2count"], entry["ref_tri2count"], entry["alt_order"] = exp["ccs_tri2count"], exp["ref_tri2count"], exp["alt_order"]
        o = oracle_outputs(c, np.array(exp["alt_order"], np.uint8))
        entry["alt_ties"] = o["ties"]
        if not np.array_equal(o["ccs"], parity.tri_dict_to_bins(exp["ccs_tri2count"])):
            why = "ccs tri counts"
        elif not np.array_equal(o["ref"], parity.tri_dict_to_bins(exp["ref_tri2count"])):
            why = "ref tri counts"
    if why is None and o["log"] != exp["log"]:
        why = "log %s != %s" % (o["log"], exp["log"])
    return entry, why


def main():
    himut = refshim.import_reference()
    if "--native-argsort" in sys.argv:
        sys.argv.remove("--native-argsort")
    else:
        himut.gtlib.np = refshim.StableArgsortNumpy()
    extra = None
    if len(sys.argv) >= 4 and sys.argv[1] == "--extra":
        extra = range(int(sys.argv[2]), int(sys.argv[3]))
    sweep = {"call": {}, "norm": {}}
    bad = []
    with tempfile.TemporaryDirectory() as tmp:
        for kind, seeds in (("call", extra or list(cases.RANDOM_CALL_SEEDS) + list(cases.RANDOM_LONG_CALL_SEEDS) + list(cases.RANDOM_DUPNAME_CALL_SEEDS)),
                            ("norm", extra or list(cases.RANDOM_NORM_SEEDS) + list(cases.RANDOM_LONG_NORM_SEEDS))):
            for seed in seeds:
                try:
                    entry, why = one(kind, seed, himut, tmp)
                except Exception as ex:  # the reference crashes on some inputs: not a case then
                    print("%s %d: reference raised %r" % (kind, seed, ex))
                    bad.append((kind, seed, "raised %r" % (ex,)))
                    continue
                if entry is None:
                    print("%s %d: no phase set" % (kind, seed))
                    continue
                print("%s %d: log=%s %s%s" % (kind, seed, entry["log"], entry.get("statuses", ""), "" if why is None else "  MISMATCH " + why))
                if why is not None:
                    bad.append((kind, seed, why))
                sweep[kind][str(seed)] = entry
    if extra is None:
        with open(os.path.join(HERE, "random_sweep.json"), "w") as f:
            json.dump(sweep, f, indent=0, separators=(",", ":"))
    print("%d cases, %d mismatches" % (len(sweep["call"]) + len(sweep["norm"]), len(bad)))
    for b in bad:
        print("  ", b)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
