"""himut_b200.genome on the CPU: a three-contig BAM called by one process and by two gloo ranks (chunk runs of a contig
on different ranks), device calls answered by the oracle stand-in.  Sharded == unsharded == the worker mirror contig by
contig (which tests/test_worker_host.py pins to the reference's outputs): rows, log vectors, distinct-read counts
across a split contig, the som_seen carry across the cut."""
import os

import numpy as np
import pytest
import torch.multiprocessing as mp

import cases
from himut_b200 import bamio, caller, genome, gtmodel, synth, worker

CONTIGS = [("chr1", 720_000, 51), ("chr2", 260_000, 52), ("chr10", 150_000, 53)]
ARGS = dict(gtmodel.DEFAULT_CALL_ARGS, md_threshold=30, non_human_sample=True)


def _make_bam(path):
    parts = []
    for name, n, seed in CONTIGS:
        d = synth.generate(n, seed=seed, depth=10.0)
        parts.append((name, n, d.batch))
    bamio.write_batches_bam(path, parts)


def _standin():
    from standin import OracleContext
    plain = worker.RegionSource.batch
    worker.RegionSource.batch = lambda self, chrom, loci, phase_sets=None, seq=True, **kw: plain(self, chrom, loci, phase_sets, seq=True, **kw)
    return OracleContext()


def _loci():
    return {c: genome.chunkloci(c, n) for c, n, _ in CONTIGS}


def _rank_main(rank, world, port, bam, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = _standin()
    lst, log, stats = genome.call_genome(bam, _loci(), ARGS, ctx=ctx)
    q.put((rank, lst, log, stats))
    dist.barrier()
    dist.destroy_process_group()


def test_plan_is_even_and_keeps_chunks_together():
    hs = [248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 58, 64, 46, 50, 156, 57]
    loci = {"chr%d" % (i + 1): genome.chunkloci("chr%d" % (i + 1), n * 125_000) for i, n in enumerate(hs)}
    for how in ("contiguous", "lpt"):
        for world in (1, 2, 4, 8):
            runs = genome.plan_runs(loci, world, how=how)
            assert runs == runs and [repr(r) for r in runs] == [repr(r) for r in genome.plan_runs(loci, world, how=how)]
            assert genome.imbalance(runs, world) < (1.01 if how == "contiguous" else 1.10)
            for c in loci:
                rs = sorted([r for r in runs if r.chrom == c], key=lambda r: r.lo)
                assert rs[0].lo == 0 and rs[-1].hi == len(loci[c]) and all(a.hi == b.lo for a, b in zip(rs, rs[1:]))
            assert {r.rank for r in runs} == set(range(world))


def test_two_ranks_equal_one_process_equal_the_worker(tmp_path, monkeypatch):
    bam = str(tmp_path / "g.bam")
    _make_bam(bam)
    # one process
    ctx = _standin()
    monkeypatch.setattr(worker, "context", lambda: ctx)
    one_lst, one_log, one_stats = genome.call_genome(bam, _loci(), ARGS, ctx=ctx)
    assert one_stats["world"] == 1 and one_stats["split_contigs"] == 0
    # the worker mirror, contig by contig
    for chrom, loci in _loci().items():
        lst, log = {}, {}
        a = ARGS
        caller.get_somatic_substitutions(
            chrom, bam, None, None, loci, {}, {}, {}, a["min_qv"], a["min_mapq"], a["qlen_lower_limit"], a["qlen_upper_limit"],
            a["min_sequence_identity"], a["min_gq"], a["min_bq"], a["min_trim"], a["max_mismatch_count"], a["mismatch_window"],
            a["md_threshold"], a["min_ref_count"], a["min_alt_count"], a["min_hap_count"], 1e-6, a["germline_snv_prior"], 1e-4,
            False, True, False, lst, log)
        assert lst[chrom] == one_lst[chrom] and log[chrom] == one_log[chrom]
        assert len(lst[chrom]) > 20
    # two gloo ranks: chr1 is cut between them
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [mpc.Process(target=_rank_main, args=(r, 2, port, bam, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in procs:
        rank, lst, log, stats = q.get(timeout=600)
        got[rank] = (lst, log, stats)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got[1][0] is None
    lst2, log2, stats2 = got[0]
    assert stats2["world"] == 2 and stats2["split_contigs"] >= 1 and stats2["imbalance_max_over_mean"] < 1.3
    assert lst2 == one_lst
    assert log2 == one_log


def test_som_seen_carries_across_a_cut():
    """two runs of one contig whose chunks overlap: the later run's candidates at positions the earlier one claimed go"""
    from himut_b200 import abi
    a = np.zeros(3, abi.SITE_DTYPE)
    a["tpos"] = [100, 200, 300]
    a["status"] = [abi.ST_PASS, abi.ST_GERM_HET, abi.ST_LOW_GQ]
    b = np.zeros(3, abi.SITE_DTYPE)
    b["tpos"] = [200, 300, 400]
    b["status"] = [abi.ST_PASS, abi.ST_PASS, abi.ST_PASS]
    out = caller.carry_som_seen([a, b], [150, None])
    assert out[0]["tpos"].tolist() == [100, 200, 300]
    assert out[1]["tpos"].tolist() == [200, 400]  # 300 was claimed (LowGQ is not a restatement), 200 was only restated
