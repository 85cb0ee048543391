#!/usr/bin/env python
"""tests/golden/cli_call.json: what the reference's own command line (`himut call`, run unmodified in the build
container through tests/cli_runner.py) wrote for the three-contig data set of cases.cli_dataset(): the VCF body, the
thresholds its BAM pre-pass put in the header and himut.log.  tests/test_zz_gpu_pool.py holds the drop-in workers —
forked pool, real CUDA context — against it on the GPU box, where the reference does not exist.

    python tests/golden/make_golden_cli.py
"""
import json
import os
import re
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import test_cli_dropin as t  # noqa: E402

PHASE_BLOCK = 50_000
PHASE_OVERRIDES = {"min_gq": "15", "min_bq": "60", "min_trim": "0.02", "mismatch_window": "30"}


def parse_log(text):
    """himut.log -> {chrom: [15 counters]}"""
    lines = [l.split() for l in text.strip().split("\n")]
    chroms = lines[0][:-1]
    return {c: [int(float(row[1 + i])) for row in lines[1:]] for i, c in enumerate(chroms)}


def thresholds(vcf_lines):
    cmd = next(l for l in vcf_lines if l.startswith("##himut_command"))
    get = lambda k: int(float(re.search(r"--%s (\S+)" % k, cmd).group(1)))
    depth = next(l for l in vcf_lines if l.startswith("##FILTER=<ID=HighDepth"))
    return [get("qlen_lower_limit"), get("qlen_upper_limit"), int(float(depth.strip().split()[-1].replace('">', "")))]


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, block in (("plain", None), ("phase", PHASE_BLOCK)):
            d = os.path.join(tmp, name)
            os.makedirs(d)
            data, bam, sets = t._inputs(d, block)
            common, pon = os.path.join(d, "common.vcf.bgz"), os.path.join(d, "pon.vcf.bgz")
            t._write_sites(common, [(c, k) for c, k, _ in sets], compress=True)
            t._write_sites(pon, [(c, k) for c, _, k in sets], compress=True)
            argv = ["call", "-i", bam, "--common_snps", common, "--panel_of_normals", pon, "-t", "3"]
            if block:
                phased = os.path.join(d, "germline.phased.vcf")
                t._write_phased(phased, data, block)
                argv += ["--phase", "--phased_vcf", phased, "--min_gq", PHASE_OVERRIDES["min_gq"], "--min_bq", PHASE_OVERRIDES["min_bq"],
                         "--min_trim", PHASE_OVERRIDES["min_trim"], "--mismatch_window_size", PHASE_OVERRIDES["mismatch_window"]]
            vcf, log, _ = t._run("reference", d, block, argv)
            body = [l for l in vcf if l and not l.startswith("#")]
            out[name] = {"thresholds": thresholds(vcf), "log": parse_log(log), "body": body}
            print(name, len(body), "rows", out[name]["thresholds"], {c: v[-1] for c, v in out[name]["log"].items()})
    with open(os.path.join(HERE, "cli_call.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))


if __name__ == "__main__":
    main()
