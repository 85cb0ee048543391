#!/usr/bin/env python
"""tests/golden/vcf_text.json: sha256 of the text the reference's own writers
(/root/reference/src/himut/vcflib.py:820-1021, dump_sbs and dump_phased_sbs) produce from the 12-tuple
rows of the committed `call_*` fixtures (rows the reference's worker itself emitted).  Build container only."""
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

HEADER = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tsynth"


def case_names():
    return sorted(f[:-5] for f in os.listdir(HERE) if f.startswith("call_") and f.endswith(".json"))


def rows_of(name):
    fx = json.load(open(os.path.join(HERE, name + ".json")))
    return [tuple(r) for r in fx["expected"]["tsbs_lst"]]


def digest(writer, rows):
    """-> {"main": sha256, "single": sha256, "main_lines": n, "single_lines": n} of what `writer` writes"""
    chrom_lst = sorted({r[0] for r in rows}) or ["chr1"]
    by_chrom = {c: [r for r in rows if r[0] == c] for c in chrom_lst}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "out.vcf")
        writer(path, HEADER, chrom_lst, by_chrom)
        out = {}
        for key, p in (("main", path), ("single", path.replace(".vcf", ".single_molecule_mutations.vcf"))):
            data = open(p, "rb").read()
            out[key] = hashlib.sha256(data).hexdigest()
            out[key + "_lines"] = data.count(b"\n")
    return out


def main():
    import refshim
    refshim.import_reference()
    import himut.vcflib as ref
    exp = {}
    for name in case_names():
        rows = rows_of(name)
        exp[name] = {"unphased": digest(ref.dump_sbs, rows), "phased": digest(ref.dump_phased_sbs, rows)}
    with open(os.path.join(HERE, "vcf_text.json"), "w") as f:
        json.dump({"source": "himut.vcflib.dump_sbs / dump_phased_sbs of /root/reference", "expected": exp}, f, indent=1, sort_keys=True)
    print({k: (v["unphased"]["main_lines"], v["unphased"]["single_lines"]) for k, v in exp.items()})


if __name__ == "__main__":
    main()
