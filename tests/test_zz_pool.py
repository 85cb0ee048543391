"""The deployment model of the drop-in: the reference driver forks `--threads` pool workers, one task per contig, each
creating its context lazily in the child (src/himut/caller.py:766-810; INTEGRATION.md "Process model").
tests/pool_runner.py reproduces that model around himut_b200.caller.get_somatic_substitutions in a fresh process; its
output must equal what the reference's own `himut call` wrote for the same three-contig BAM and site files
(tests/golden/cli_call.json, from tests/golden/make_golden_cli.py): the BAM pre-pass thresholds, the log vector of
every contig and every VCF body line.  On the GPU box the workers use the CUDA library (three processes sharing the
device); on the CPU the oracle stand-in pins the host side.  Runs last (file name) because it spawns processes."""
import gzip
import json
import os
import subprocess
import sys

import pytest

import cases
from himut_b200 import bamio

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(cases.GOLDEN_DIR, "cli_call.json")))
PHASE_BLOCK = 50_000
PHASE_OVERRIDES = ["min_gq=15", "min_bq=60", "min_trim=0.02", "mismatch_window=30"]


def _write_sites(path, per_contig):
    text = ["##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tsynth\n"]
    for chrom, keys in per_contig:
        for k in keys.tolist():
            text.append("%s\t%d\t.\t%s\t%s\t.\tPASS\t.\tGT\t0/1\n" % (chrom, k >> 4, "ATGC"[(k >> 2) & 3], "ATGC"[k & 3]))
    with open(path, "wb") as f:
        f.write(gzip.compress("".join(text).encode()))


def _run_group(cmd, env, timeout):
    """subprocess.run, but the runner and the pool workers it forks form one process group that is killed as a whole
    when the time is up: nothing is left behind holding the device"""
    import signal
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, start_new_session=True)
    try:
        out, err = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        os.killpg(p.pid, signal.SIGKILL)
        out, err = p.communicate()
        pytest.fail("pool runner did not finish within %d s\n%s\n%s" % (timeout, out[-2000:], err[-4000:]))
    return subprocess.CompletedProcess(cmd, p.returncode, out, err)


def _run(tmp, name, real):
    block = PHASE_BLOCK if name == "phase" else 0
    data = cases.cli_dataset(block or None)
    bam = os.path.join(tmp, "synth.bam")
    bamio.write_batches_bam(bam, [(c, n, d.batch) for c, n, d in data])
    sets = [(c,) + cases.site_sets_from_synth(d, 40 + i) for i, (c, _n, d) in enumerate(data)]
    common, pon = os.path.join(tmp, "common.vcf.bgz"), os.path.join(tmp, "pon.vcf.bgz")
    _write_sites(common, [(c, k) for c, k, _ in sets])
    _write_sites(pon, [(c, k) for c, _, k in sets])
    cmd = [sys.executable, os.path.join(HERE, "pool_runner.py"), tmp, bam, common, pon, "3", str(block)] + (PHASE_OVERRIDES if block else [])
    env = dict(os.environ)
    env.pop("HIMUT_B200_DEVICE", None)
    env.pop("LOCAL_RANK", None)  # device = pool worker index modulo visible GPUs, as in a plain `himut call`
    env["HIMUT_B200_CLI_REAL_CONTEXT"] = "1" if real else "0"
    r = _run_group(cmd, env, timeout=600)
    if real and r.returncode != 0 and "hm_create(device=" in (r.stdout + r.stderr):
        # the workers could not open the device although this process has it open: exclusive-process compute mode
        pytest.skip("a second process cannot open the GPU while pytest holds a context (exclusive-process mode)")
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    return json.loads(r.stdout.strip().split("\n")[-1])


def _check(got, name):
    exp = GOLD[name]
    assert got["thresholds"] == exp["thresholds"]
    assert got["log"] == exp["log"]
    assert len(got["body"]) == len(exp["body"])
    assert got["body"] == exp["body"], next((a, b) for a, b in zip(got["body"], exp["body"]) if a != b)


@pytest.mark.parametrize("name", ["plain", "phase"])
def test_forked_pool_matches_reference_cli_cpu(tmp_path, name):
    _check(_run(str(tmp_path), name, real=False), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["plain", "phase"])
def test_forked_pool_matches_reference_cli_gpu(tmp_path, name):
    _check(_run(str(tmp_path), name, real=True), name)


# ---- `himut normcounts` and `himut phase` workers in the same process model -------------------------------------
GOLD2 = json.load(open(os.path.join(cases.GOLDEN_DIR, "cli_norm_phase.json")))


def _run_mode(tmp, real, extra):
    data = cases.cli_dataset(None)
    bam = os.path.join(tmp, "synth.bam")
    bamio.write_batches_bam(bam, [(c, n, d.batch) for c, n, d in data])
    sets = [(c,) + cases.site_sets_from_synth(d, 40 + i) for i, (c, _n, d) in enumerate(data)]
    common, pon = os.path.join(tmp, "common.vcf.bgz"), os.path.join(tmp, "pon.vcf.bgz")
    _write_sites(common, [(c, k) for c, k, _ in sets])
    _write_sites(pon, [(c, k) for c, _, k in sets])
    cmd = [sys.executable, os.path.join(HERE, "pool_runner.py"), tmp, bam, common, pon, "2", "0"] + extra
    env = dict(os.environ)
    env.pop("HIMUT_B200_DEVICE", None)
    env.pop("LOCAL_RANK", None)
    env["HIMUT_B200_CLI_REAL_CONTEXT"] = "1" if real else "0"
    r = _run_group(cmd, env, timeout=900)
    if real and r.returncode != 0 and "hm_create(device=" in (r.stdout + r.stderr):
        pytest.skip("a second process cannot open the GPU while pytest holds a context (exclusive-process mode)")
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    return json.loads(r.stdout.strip().split("\n")[-1])


def _check_normcounts(got):
    exp = GOLD2["normcounts"]
    assert got["log"] == exp["log"]
    assert got["tri"] == {k: v for k, v in exp["tri"].items()}


def _norm_args():
    qlo, qhi, md = GOLD2["normcounts"]["thresholds"]
    return ["mode=normcounts", "qlo=%d" % qlo, "qhi=%d" % qhi, "md=%r" % md, "contigs=" + ",".join(GOLD2["normcounts"]["contigs"])]


def test_normcounts_pool_matches_reference_cli_cpu(tmp_path):
    """the callable-base workers in `himut normcounts`' process model == the TSV columns and norm.log the reference's
    own command line wrote (tests/golden/make_golden_cli2.py)"""
    _check_normcounts(_run_mode(str(tmp_path), False, _norm_args()))


@pytest.mark.gpu
def test_normcounts_pool_matches_reference_cli_gpu(tmp_path):
    _check_normcounts(_run_mode(str(tmp_path), True, _norm_args()))


def test_phase_edge_pool_matches_reference_cpu(tmp_path):
    """the phase-edge workers in `himut phase`'s process model == the reference's own get_edges per contig"""
    assert _run_mode(str(tmp_path), False, ["mode=phase_edges"])["edges"] == GOLD2["phase_edges"]


@pytest.mark.gpu
def test_phase_edge_pool_matches_reference_gpu(tmp_path):
    assert _run_mode(str(tmp_path), True, ["mode=phase_edges"])["edges"] == GOLD2["phase_edges"]
