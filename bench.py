#!/usr/bin/env python
"""bench.py — `himut call` hot-path throughput on synthetic 30x CCS data (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--contig-mb M]

A step is one pass of the whole `himut call` device path (k_read_scan -> k_candidates -> sort ->
k_site_range|entries|reduce -> host som_seen replay) over one contig's packed read batch.

  value   aligned CCS bases/s with the batch already resident in HBM (hm_call_chunks),
          CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks
  e2e     the same metric through the C-ABI call a worker makes with HOST buffers
          (hm_call_batch: pinned host -> device copies, kernels, records back)
  roofline  dominant kernel (k_read_scan): algorithmic bytes / its CUDA-event duration vs the
          measured HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle port (oracle/himut_oracle.c, one thread per host core, each like a reference worker on a
          single contig) on a bounded sample of the same workload, rank 0 only

N > 1 (torchrun, one rank per GPU): every rank owns one contig of the same size (different
seed) — the genome shards by contig with no data-path collective; the 15 log counters are
all-reduced over NCCL and record counts gathered at the end.  Weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def chunkloci(end):
    out = [(1, 200000)]
    starts = list(range(200000, end, 200000))
    for i, s in enumerate(starts[:-1]):
        out.append((s, starts[i + 1]))
    out.append((starts[-1], end - 2))
    return out


def make_workload(contig_len, seed):
    from himut_b200 import gtmodel, synth
    d = synth.generate(contig_len, seed=seed, copy=False)
    params = gtmodel.make_params(**gtmodel.DEFAULT_CALL_ARGS)
    chunks = d.batch.chunk_table(chunkloci(contig_len))
    return d, params, chunks


def kernel_alg_bytes(batch):
    """algorithmic bytes of one k_read_scan launch: every quality byte once,
    every op word once, per-read metadata once; outputs: two u32 prefix words + one mismatch
    position per op at most, 29 B of per-read results"""
    n_base = int(batch.qlen.astype(np.int64).sum())
    n_op = int(batch.ops.size)
    n_read = int(batch.n_reads)
    return 1.0 * n_base + 4.0 * n_op + 33.0 * n_read + 12.0 * n_op + 29.0 * n_read, n_base, n_op, n_read


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(gpu_index):
    """one process per GPU: run it on the cores next to that GPU (NVML's ideal CPU set), so eight ranks do not
    share cores or cross sockets for their pinned-memory copies"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0)
        if cpus & allowed:
            os.sched_setaffinity(0, cpus & allowed)
    except Exception:
        pass


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(alg_bytes):
    """DRAM bytes per k_read_scan launch: dram/algorithmic ratio of the committed ncu capture
    (profiles/traffic.json, taken on a 16 Mb contig) applied to this launch's algorithmic bytes"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return float(json.load(open(p))["k_read_scan_dram_over_alg"]) * alg_bytes
    except Exception:
        return None


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(d, params, chunks, n_chunks, threads=1):
    """oracle port on `threads` host threads, each over its own run of n_chunks consecutive chunks of the contig
    (the reference's own parallelism is one worker per contig, `caller.py:593-618`: a thread here stands for
    one such worker; ctypes drops the GIL for the duration of the C call)
    -> (bases/s, bases, seconds, threads used)"""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    b = d.batch
    threads = max(1, min(threads, len(chunks) // max(1, n_chunks)))
    groups = [chunks[t * n_chunks:(t + 1) * n_chunks] for t in range(threads)]
    kind, val = b.ops & 3, (b.ops >> 2).astype(np.int64)
    qspan = np.where(kind == 0, val, 0) + (kind == 1) + np.where(kind == 2, val, 0)
    per_read = np.add.reduceat(qspan, b.op_off.astype(np.int64))
    bases = 0
    for sub in groups:
        # distinct reads fetched by the group's chunks = reads [lo, hi) that overlap [start0, endN)
        lo, hi = int(sub["read_lo"].min()), int(sub["read_hi"].max())
        ov = (b.tstart[lo:hi] < int(sub["end"][-1])) & (b.tend[lo:hi] > int(sub["start"][0]))
        bases += int(per_read[lo:hi][ov].sum())
    cap = max(4096, int(b.ops.size) // max(1, len(chunks)) * n_chunks * 4)
    t0 = time.perf_counter()
    if threads == 1:
        oracle.call_chunks(params, b, groups[0], cap=cap)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda sub: oracle.call_chunks(params, b, sub, cap=cap), groups))
    dt = time.perf_counter() - t0
    return bases / dt, bases, dt, threads


def bam_to_records(ctx, params, contig_mb, seed):
    """what a worker does per contig, from a BAM file in the page cache: native decode (threaded BGZF inflate +
    record parse + cs -> ops, csrc/bamdec.c) -> pinned-less host batch -> hm_call_batch.  Host bound by design."""
    import shutil
    from himut_b200 import bamdec, synth
    n = contig_mb * 1_000_000
    d = synth.generate(n, seed=seed + 1, copy=False)
    tmp = tempfile.mkdtemp(prefix="himut_b200_bench_")
    try:
        path = os.path.join(tmp, "synth.bam")
        t0 = time.perf_counter()
        bamdec.write_batch_bam(path, "chr1", n, d.batch)
        t_write = time.perf_counter() - t0
        threads = bamdec.default_threads()
        bam = bamdec.NativeBam(path, threads=threads)
        chunks = None
        best, best_dec = None, None
        for _ in range(3):
            t0 = time.perf_counter()
            batch = bam.read_batch("chr1", 0, n, copy=False, seq=False)  # as the worker mirror does (caller.py):
            t1 = time.perf_counter()                                    # `call` needs no base stream
            if chunks is None:
                chunks = batch.chunk_table(chunkloci(n))
            rec, log = ctx.call_batch(batch, chunks, view=True)
            t2 = time.perf_counter()
            if best is None or t2 - t0 < best:
                best, best_dec = t2 - t0, t1 - t0
        bam.close()
        out = {"value": d.aligned_bases / best, "unit": "bases/s", "contig_mb": contig_mb, "decode_threads": threads,
               "seconds": best, "decode_seconds": best_dec, "bam_bytes": os.path.getsize(path), "bam_write_seconds": t_write,
               "site_records": int(rec.size),
               "note": "BAM (page cache) -> records: bounded by host BGZF inflate + record parse, not by the GPU"}
        try:
            out["to_vcf"] = bam_to_vcf(path, n, d.aligned_bases, tmp)
        except Exception as ex:  # a sub-measurement: never costs the line
            out["to_vcf"] = {"error": repr(ex)}
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def bam_to_vcf(bam_path, contig_len, aligned_bases, tmp):
    """SURVEY 8(d) timing (ii): BAM on disk -> VCF on disk through the worker mirror itself
    (himut_b200.caller.get_somatic_substitutions: native decode, upload, device path, 12-tuples, natsort) and the
    mirror of the reference's writer (vcfio.dump_sbs).  Its own hm_ctx, like a worker process."""
    from himut_b200 import caller, gtmodel, vcfio
    a = gtmodel.DEFAULT_CALL_ARGS
    loci = [("chr1", s, e) for s, e in chunkloci(contig_len)]
    header = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tsynth"
    best = None
    for _ in range(3):
        lst, log = {}, {}
        t0 = time.perf_counter()
        caller.get_somatic_substitutions(
            "chr1", bam_path, None, None, loci, {}, {}, {}, a["min_qv"], a["min_mapq"], a["qlen_lower_limit"],
            a["qlen_upper_limit"], a["min_sequence_identity"], a["min_gq"], a["min_bq"], a["min_trim"], a["max_mismatch_count"],
            a["mismatch_window"], a["md_threshold"], a["min_ref_count"], a["min_alt_count"], a["min_hap_count"], 1e-6,
            a["germline_snv_prior"], 1e-4, False, True, False, lst, log)
        t1 = time.perf_counter()
        vcf = os.path.join(tmp, "out.vcf")
        vcfio.dump_sbs(vcf, header, ["chr1"], lst)
        t2 = time.perf_counter()
        if best is None or t2 - t0 < best[0]:
            best = (t2 - t0, t1 - t0, t2 - t1)
    return {"value": aligned_bases / best[0], "unit": "bases/s", "seconds": best[0], "worker_seconds": best[1],
            "writer_seconds": best[2], "rows": len(lst["chr1"]), "vcf_bytes": os.path.getsize(vcf),
            "log": [int(v) for v in log["chr1"]],
            "note": "non_human_sample = True (no common-SNP / PoN files), otherwise the defaults of the resident workload"}


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path (the oracle port: the reference is
    pure Python and is not on this box), one thread per host core, on a bounded sample of the workload."""
    if rank != 0:
        return
    contig_len = args.contig_mb * 1_000_000
    d, params, chunks = make_workload(contig_len, args.seed)
    threads = host_threads()
    n = max(1, min(len(chunks) // threads, args.cpu_chunks))
    for _ in range(args.warmup):
        cpu_baseline(d, params, chunks, n, threads)
    tot, el = 0, 0.0
    for _ in range(args.steps):
        v, bases, dt, used = cpu_baseline(d, params, chunks, n, threads)  # dt: the oracle calls only
        tot += bases
        el += dt
    value = tot / el
    sample = "%d threads x %d consecutive chunks = %d of %d chunks (%.1f Mb, %d aligned bases) of the %d Mb 30x contig per step" % (
        used, n, used * n, len(chunks), used * n * 0.2, bases, args.contig_mb)
    print(json.dumps({
        "impl": "reference", "metric": "aligned CCS bases/sec (himut call, 30x synthetic)", "value": value, "unit": "bases/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "bases/s", "cores": used, "kind": "port", "sample": sample,
                         "note": "C restatement of the pure-Python reference (oracle/himut_oracle.c); the reference itself ran "
                                 "at 3e5-6e5 bases/s/core in the build container (tests/golden/*.json reference_seconds)"},
        "e2e": {"value": value, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, world):
    return {"workload": "himut call on a %d Mb synthetic contig at 30x CCS (15 kb reads, cs tags), %d x 200 kb chunks%s"
                        % (args.contig_mb, len(chunkloci(args.contig_mb * 1_000_000)), " per GPU" if world > 1 else ""),
            "baseline_config": "configs[1]" if args.contig_mb == 64 else "custom",
            "contig_mb": args.contig_mb, "depth": 30, "parallelism": "contig-sharded x%d" % world,
            "l2": "inputs (>= 2 GB per step) are larger than the 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--contig-mb", type=int, default=64)
    ap.add_argument("--seed", type=int, default=20260101)
    ap.add_argument("--cpu-chunks", type=int, default=25)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-normcounts", action="store_true")
    ap.add_argument("--no-bam-leg", action="store_true")
    ap.add_argument("--bam-mb", type=int, default=8)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import __graft_entry__ as g
    g.build()

    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from himut_b200 import lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=dev)

    contig_len = args.contig_mb * 1_000_000
    d, params, chunks = make_workload(contig_len, args.seed + 7919 * rank)
    batch = d.batch
    ctx = lib.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_params(params)
    ctx.set_site_sets()
    # records of candidates that merely restate the germline genotype are counted on the device and not copied back:
    # the reference drops them too (caller.py:338-345)
    ctx.omit_restatements(True)
    ctx.pin(batch)
    aligned = d.aligned_bases

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident (HBM) timing ----------------
    # resident as the worker mirror leaves it (caller.py): no base stream — `call` takes the bases of match runs from
    # the site's reference allele; the end-to-end legs below check those records byte for byte against a batch with bases
    ctx.upload(batch.without_seq())
    for _ in range(args.warmup):
        rec, log = ctx.call_chunks(chunks, view=True)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:  # one sampler for the job: eight concurrent NVML pollers slow every rank's driver calls down
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_ms = {}
    e0.record(stream)
    for _ in range(args.steps):
        # the record copy of a step (20 MB, device -> pinned host) overlaps the next step's kernels
        rec, log = ctx.call_chunks(chunks, view=True, wait=False)
        for name, ms in ctx.last_kernel_times():
            k_ms.setdefault(name, []).append(ms)
    ctx.records_wait()  # every step's records are in host memory before the clock stops
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)

    # ---------------- end to end (host buffers through the C ABI) ----------------
    # e2e: hm_call_batch_compact — the batch in pinned host memory with its quality stream as modal-value bitmap +
    # exceptions (lossless, built once when the batch is produced, expanded on the device inside the timed region);
    # e2e_plain: hm_call_batch with one quality byte per base.
    from himut_b200 import bamdec
    cq = bamdec.compact_bq(batch)
    small = [getattr(batch, n) for n, _ in batch._FIELDS if n not in ("seq", "bq", "ops")]
    ctx.pin_arrays([cq.mask, cq.exc, cq.exc_off] + small)
    for _ in range(2):
        ctx.call_batch(batch, chunks, view=True)
        ctx.call_batch_compact(batch, cq, chunks, view=True)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(args.steps):
        rec_e, log_e = ctx.call_batch_compact(batch, cq, chunks, view=True)
    f1.record(stream)
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    assert list(log_e) == list(log)
    # the same call for a batch without its base stream (hm_read_batch.seq = NULL): the bases of match runs are the
    # site's reference allele by the meaning of a cs match, so a worker need not ship them.  Counted as the
    # end-to-end figure only if its records are byte for byte those of the call with bases.
    ms_e2e_ns, noseq_note = float("inf"), None
    try:
        rec_with = rec_e.copy()
        batch_ns = batch.without_seq()
        for _ in range(2):
            ctx.call_batch_compact(batch_ns, cq, chunks, view=True)
        barrier()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record(stream)
        for _ in range(args.steps):
            rec_n, log_n = ctx.call_batch_compact(batch_ns, cq, chunks, view=True)
        n1.record(stream)
        barrier()
        srt = lambda a: np.sort(a, order=["chunk", "tpos", "ref", "alt"]).tobytes()  # records come back in no particular order
        if list(log_n) == list(log) and srt(rec_n) == srt(rec_with):
            ms_e2e_ns = n0.elapsed_time(n1)
        else:
            noseq_note = "records differ from the call with bases: not counted"
        del rec_with
    except Exception as ex:  # keep the line: the leg with bases stands
        noseq_note = "failed: %r" % (ex,)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record(stream)
    for _ in range(args.steps):
        rec_p, log_p = ctx.call_batch(batch, chunks, view=True)
    h1.record(stream)
    barrier()
    ms_e2e_plain = h0.elapsed_time(h1)
    assert list(log_p) == list(log)
    clocks = sampler.stop()  # sampled from the first timed resident step to the last end-to-end step
    h2d_compact = int(batch.nbytes() - batch.bq.nbytes + cq.nbytes() + chunks.nbytes)

    # ---------------- callable-base half of `himut normcounts` on the same resident batch ----------------
    norm = None
    if not args.no_normcounts:
        ctx.upload(batch)
        ctx.normcounts_chunks(d.ref, chunks)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_norm = max(2, args.steps // 3)
        nk = {}
        g0.record(stream)
        for _ in range(n_norm):
            ccs_tri, ref_tri, nlog, nties = ctx.normcounts_chunks(d.ref, chunks)
            for name, ms in ctx.last_kernel_times():
                nk.setdefault(name, []).append(ms)
        g1.record(stream)
        barrier()
        ms_norm = g0.elapsed_time(g1) / n_norm
        norm = {"value": aligned / (ms_norm * 1e-3), "unit": "bases/s", "ms_per_step": ms_norm, "steps": n_norm,
                "kernel_ms_per_step": {k: float(np.mean(v)) for k, v in nk.items()},
                "callable_bases": int(nlog[13]), "callable_positions": int(ref_tri.sum()), "alt_ties_flagged": int(nties),
                "positions_evaluated_exactly": ctx.last_norm_exact_sites(), "positions": int(contig_len),
                "bound": "instruction issue (integer pass over every aligned base) + the exact fp64 pass over the listed positions, see DESIGN.md"}

    # ---------------- BAM on disk -> site records (decode included), rank 0, a separate 8 Mb contig ----------------
    bam_leg = None
    if rank == 0 and not args.no_bam_leg:
        try:
            bam_leg = bam_to_records(ctx, params, args.bam_mb, args.seed)
        except Exception as ex:  # a side measurement: never costs the line
            bam_leg = {"error": repr(ex)}

    # ---------------- reduce over ranks ----------------
    t = torch.tensor([ms_total, ms_e2e, ms_e2e_plain, ms_e2e_ns], device=dev, dtype=torch.float64)
    tot = torch.tensor([float(aligned), float(rec.size)], device=dev, dtype=torch.float64)
    logt = torch.tensor(np.asarray(log, np.int64), device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(logt, op=dist.ReduceOp.SUM)  # the only cross-GPU step of the path: 15 counters
    ms_total, ms_e2e, ms_e2e_plain, ms_e2e_ns = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    all_bases, all_recs = float(tot[0]), int(tot[1])

    if rank == 0:
        value = all_bases * args.steps / (ms_total * 1e-3)
        e2e = all_bases * args.steps / (ms_e2e * 1e-3)
        e2e_with = {"value": e2e, "unit": "bases/s", "h2d_bytes_per_step": h2d_compact,
                    "d2h_bytes_per_step": int(rec.nbytes + 32), "ms_per_step": ms_e2e / args.steps,
                    "call": "hm_call_batch_compact (quality stream as modal bitmap + exceptions, expanded on the device)"}
        if np.isfinite(ms_e2e_ns):
            e2e_best = {"value": all_bases * args.steps / (ms_e2e_ns * 1e-3), "unit": "bases/s",
                        "h2d_bytes_per_step": int(h2d_compact - batch.seq.nbytes - batch.seq_off.nbytes),
                        "d2h_bytes_per_step": int(rec.nbytes + 32), "ms_per_step": ms_e2e_ns / args.steps,
                        "call": "hm_call_batch_compact, batch without a base stream (seq = NULL: match runs carry the reference's "
                                "bases by the meaning of cs; substituted bases are in the ops) and the quality stream as modal "
                                "bitmap + exceptions; records byte-identical to the call with bases"}
        else:
            e2e_best = dict(e2e_with, note_without_bases=noseq_note or "not counted on some rank")
        alg, n_base, n_op, n_read = kernel_alg_bytes(batch)
        scan_name = "k_call_scan" if "k_call_scan" in k_ms else "k_read_scan"
        scan_ms = float(np.mean(k_ms[scan_name]))
        peak, peak_src = measured_peak()
        achieved = alg / (scan_ms * 1e-3) / 1e9
        step_ms = {k: float(np.mean(v)) for k, v in k_ms.items()}
        dev_ms = sum(step_ms.values())
        # whole-step figure with SURVEY.md §8(d)'s formula (1.25 B per shipped base + ops + reads + records)
        survey_bytes = 1.25 * n_base + 4.0 * n_op + 40.0 * n_read + 48.0 * rec.size
        out = {
            "metric": "aligned CCS bases/sec (himut call, 30x synthetic)", "value": value, "unit": "bases/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clocks,
            "e2e": e2e_best,
            "e2e_with_bases": e2e_with,
            "e2e_plain": {"value": all_bases * args.steps / (ms_e2e_plain * 1e-3), "unit": "bases/s",
                          "h2d_bytes_per_step": int(batch.nbytes() + chunks.nbytes), "d2h_bytes_per_step": int(rec.nbytes + 32),
                          "ms_per_step": ms_e2e_plain / args.steps, "call": "hm_call_batch (one quality byte per base)"},
            # own kernels per resident step: k_read_scan, k_candidates, k_expand_keys, k_site_range, k_chunk_key_ranges,
            # k_site_entries_by_read, k_site_reduce, k_keep_flags, k_compact_records, k_gather_u32, k_count_flags,
            # 3 x k_publish, 3 x k_publish_items (the cub sort / unique / scan launches are library code, not counted)
            "gpu_launches": int(args.steps * 17),
            "dominant_kernel": max(step_ms, key=step_ms.get),
            "library_launches_per_step": "cub::DeviceRadixSort (candidate keys)",
            "roofline": {"bound": "hbm", "kernel": scan_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "frac_of_nominal_8000_gbs": achieved / 8000.0,  # north_star's ~8 TB/s
                         "traffic": ncu_traffic(alg), "peak_source": peak_src,
                         "alg_bytes_per_launch": alg, "kernel_ms": scan_ms,
                         "share_of_step_device_time": scan_ms / dev_ms if dev_ms else None},
            "roofline_step": {"formula": "SURVEY 8(d): 1.25*N_base + 4*N_op + 40*N_read + 48*N_cand", "bytes": survey_bytes,
                              "note": "the formula's 0.25 B per base is the 2-bit base stream, which this resident batch does not carry: "
                                      "by bytes actually needed the fraction is lower by 1.0/1.25 on that term",
                              "device_ms": dev_ms, "achieved_gbs": survey_bytes / (dev_ms * 1e-3) / 1e9,
                              "frac_of_peak": survey_bytes / (dev_ms * 1e-3) / 1e9 / peak},
            "kernel_ms_per_step": step_ms,
            "aligned_bases_per_step": int(all_bases), "site_records_per_step": all_recs,
            "records": "every record the reference emits (PASS + filtered sites) reaches host memory each step; germline "
                       "restatements are counted in the log on the device (HM_OPT_OMIT_RESTATEMENTS)",
            "log_counters_sum": [int(v) for v in logt.tolist()],
        }
        if norm is not None:
            out["normcounts"] = norm
        if bam_leg is not None:
            out["bam_to_records"] = bam_leg
        if not args.no_cpu_baseline and world == 1:  # the CPU arm beside the one-GPU line only
            threads = host_threads()
            n = max(1, min(len(chunks) // threads, args.cpu_chunks))
            v1, bases1, dt1, _ = cpu_baseline(d, params, chunks, n, 1)
            v, bases, dt, used = cpu_baseline(d, params, chunks, n, threads)
            out["cpu_baseline"] = {"value": v, "unit": "bases/s", "cores": used, "kind": "port",
                                   "sample": "%d threads x %d consecutive chunks of %d (%d aligned bases, %.1f s) of the same contig"
                                             % (used, n, len(chunks), bases, dt),
                                   "one_core": {"value": v1, "bases": bases1, "seconds": dt1}}
        print(json.dumps(out))
    ctx.unpin(batch)
    ctx.unpin_arrays([cq.mask, cq.exc, cq.exc_off] + small)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
