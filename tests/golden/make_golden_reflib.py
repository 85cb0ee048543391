#!/usr/bin/env python
"""tests/golden/reflib.json: the reference's own reflib.get_chrom_tricount
(/root/reference/src/himut/reflib.py:11-33) on deterministic sequences.  Build container only."""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refshim  # noqa: E402


def sequences():
    rnd = random.Random(5)
    out = {"plain": "".join(rnd.choice("ACGT") for _ in range(20011)),
           "dirty": "".join(rnd.choice("ACGTNacgtnRY") if rnd.random() < 0.1 else rnd.choice("ACGT") for _ in range(9973)),
           "tiny": "ACG", "two": "AC", "empty": "", "nrun": "NNNNACGTNNACGTTTNN"}
    return out


def main():
    refshim.import_reference()
    import himut.reflib
    exp = {}
    for name, seq in sequences().items():
        d = {}
        himut.reflib.get_chrom_tricount(name, seq, d)
        exp[name] = {k: int(v) for k, v in d[name].items()}
    with open(os.path.join(HERE, "reflib.json"), "w") as f:
        json.dump({"expected": exp}, f, separators=(",", ":"))
    print({k: sum(v.values()) for k, v in exp.items()})


if __name__ == "__main__":
    main()
