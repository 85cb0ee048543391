"""configs[2] in small through the product path on hardware: a multi-contig BAM called by himut_b200.genome on one GPU
and, under torchrun, on two (four) GPUs — chunk runs of a contig on different GPUs — must give identical rows and log
vectors, and both must equal the oracle's (the stand-in context answers the same host code).  Needs >= 2 GPUs for the
sharded part (skipped otherwise; `gpurun --gpus 2`).  GPU."""
import json
import os
import subprocess
import sys

import pytest

import parity
from himut_b200 import bamio, genome, gtmodel, synth, worker

pytestmark = pytest.mark.gpu
CONTIGS = [("chr1", 1_400_000, 61), ("chr2", 700_000, 62), ("chr10", 300_000, 63), ("chrX", 150_000, 64)]
HERE = os.path.dirname(os.path.abspath(__file__))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.fixture(scope="module")
def job(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("genome")
    bam = str(tmp / "g.bam")
    parts = [(c, n, synth.generate(n, seed=s, copy=False).batch) for c, n, s in CONTIGS]
    bamio.write_batches_bam(bam, parts)
    cj = str(tmp / "contigs.json")
    json.dump([[c, n] for c, n, _ in CONTIGS], open(cj, "w"))
    return tmp, bam, cj


def _run(tmp, bam, cj, n):
    out = str(tmp / ("out%d.json" % n))
    env = dict(os.environ)
    env.pop("HIMUT_B200_DEVICE", None)
    if n == 1:
        cmd = [sys.executable, os.path.join(HERE, "genome_runner.py"), bam, out, cj]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
               "--master-port", str(29600 + n), os.path.join(HERE, "genome_runner.py"), bam, out, cj]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    return json.load(open(out))


def test_one_gpu_equals_the_oracle(job):
    tmp, bam, cj = job
    one = _run(tmp, bam, cj, 1)
    # the same host code with the device calls answered by the oracle
    from standin import OracleContext
    plain = worker.RegionSource.batch
    worker.RegionSource.batch = lambda self, chrom, loci, phase_sets=None, seq=True, **kw: plain(self, chrom, loci, phase_sets, seq=True, **kw)
    try:
        args = dict(gtmodel.DEFAULT_CALL_ARGS, non_human_sample=True)
        loci = {c: genome.chunkloci(c, n) for c, n, _ in CONTIGS}
        lst, log, _ = genome.call_genome(bam, loci, args, ctx=OracleContext())
    finally:
        worker.RegionSource.batch = plain
    assert one["digest"] == {c: parity.rows_digest(lst[c]) for c in lst}
    assert one["log"] == log
    assert sum(one["rows"].values()) > 1000


@pytest.mark.parametrize("n", [2, 4])
def test_n_gpus_equal_one_gpu(job, n):
    if _n_gpus() < n:
        pytest.skip("%d GPUs needed, %d visible" % (n, _n_gpus()))
    tmp, bam, cj = job
    one = _run(tmp, bam, cj, 1)
    many = _run(tmp, bam, cj, n)
    assert many["stats"]["world"] == n and many["stats"]["split_contigs"] >= 1
    assert many["digest"] == one["digest"] and many["log"] == one["log"] and many["rows"] == one["rows"]
