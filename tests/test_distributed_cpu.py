"""world_size-2 gloo test of the sharding + final reduction (host logic of the N > 1 path).
The per-contig compute stand-in is the CPU oracle: this file checks that sharded == unsharded."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

import cases
from himut_b200 import gtmodel, records, shard, synth
from oracle import oracle

CONTIGS = {"chr1": (90_000, 41), "chr2": (60_000, 42), "chr3": (40_000, 43), "chr4": (40_000, 44)}


def _contig_result(name):
    n, seed = CONTIGS[name]
    d = synth.generate(n, seed=seed, depth=12.0)
    p = gtmodel.make_params(**dict(gtmodel.DEFAULT_CALL_ARGS, md_threshold=30))
    rec, log = oracle.call_chunks(p, d.batch, d.batch.chunk_table([(0, n)]))
    return records.records_to_tsbs_lst(name, rec), np.asarray(log, np.int64)


def _rank_main(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    weights = {c: v[0] for c, v in CONTIGS.items()}
    mine = shard.my_contigs(weights, rank, world)
    lst, log = {}, np.zeros(15, np.int64)
    for c in mine:
        rows, l = _contig_result(c)
        lst[c] = rows
        log += l
    total = shard.all_reduce_sum(log)
    merged = shard.gather_dicts(lst, dst=0)
    if rank == 0:
        q.put((sorted(mine), total.tolist(), {c: len(v) for c, v in merged.items()}, merged["chr3"][:3]))
    else:
        q.put((sorted(mine), total.tolist(), None, None))
    dist.barrier()
    dist.destroy_process_group()


def test_lpt_assign_is_balanced_and_deterministic():
    w = {"a": 10, "b": 9, "c": 5, "d": 4, "e": 1}
    a = shard.lpt_assign(w, 2)
    assert a == shard.lpt_assign(dict(reversed(list(w.items()))), 2)
    loads = [sum(w[k] for k in w if a[k] == r) for r in range(2)]
    assert max(loads) - min(loads) <= 1
    assert sorted(sum((shard.my_contigs(w, r, 3) for r in range(3)), [])) == sorted(w)


def test_two_rank_gloo_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    owned = sorted(sum((g[0] for g in got), []))
    assert owned == sorted(CONTIGS)
    single_log = np.zeros(15, np.int64)
    single_rows = {}
    for c in CONTIGS:
        rows, l = _contig_result(c)
        single_rows[c] = rows
        single_log += l
    for g in got:
        assert g[1] == single_log.tolist()
    root = [g for g in got if g[2] is not None][0]
    assert root[2] == {c: len(v) for c, v in single_rows.items()}
    assert [tuple(r) for r in root[3]] == [tuple(r) for r in single_rows["chr3"][:3]]
