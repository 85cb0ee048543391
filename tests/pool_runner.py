#!/usr/bin/env python
"""Test infrastructure: the reference driver's process model around the drop-in `call` worker, without the reference —
what `himut call` does between option parsing and the writers (src/himut/caller.py:676-830): thresholds from the BAM
pre-pass, one chunk list per contig, `multiprocessing.Pool(threads)` (fork) with Manager dicts, one starmap task per
contig, then the writer.  Run in a fresh process (CUDA must not exist in the parent before the fork):

    python tests/pool_runner.py <workdir> <bam> <common or -> <pon or -> <threads> <phase_block or 0> [key=value ...]

Prints one JSON line: thresholds, per-contig log vectors, the body lines of the VCF.  Workers use the CUDA library
when HIMUT_B200_CLI_REAL_CONTEXT=1, else the oracle stand-in (tests/standin.py) so the script itself is pinned on the CPU.
"""
import json
import multiprocessing as mp
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

HEADER = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tsynth"


def main():
    work, bam, common, pon, threads, block = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
    over = dict(kv.split("=") for kv in sys.argv[7:])
    common, pon = (None if common == "-" else common), (None if pon == "-" else pon)
    import __graft_entry__ as g
    g.build()
    import cases
    from himut_b200 import bamlib, caller, gtmodel, natsort_compat, synth, vcfio, worker
    if os.environ.get("HIMUT_B200_CLI_REAL_CONTEXT") != "1":
        import standin
        ctx = standin.OracleContext()
        worker.context = lambda: ctx
        plain = worker.RegionSource.batch
        worker.RegionSource.batch = lambda self, chrom, loci, phase_sets=None, seq=True, **kw: plain(self, chrom, loci, phase_sets, seq=True, **kw)

    mode = over.pop("mode", "call")
    if mode == "normcounts":
        return run_normcounts(work, bam, common, pon, threads, over)
    if mode == "phase_edges":
        return run_phase_edges(bam, threads)
    a = dict(gtmodel.DEFAULT_CALL_ARGS)
    for k, v in over.items():
        a[k] = type(a[k])(v)
    chrom2len = {c: n for c, n, _ in cases.CLI_CONTIGS}
    chrom_lst = natsort_compat.natsorted(list(chrom2len))
    qlo, qhi, md = bamlib.get_thresholds(bam, chrom_lst, chrom2len)  # in the parent, like caller.py:690: host only
    hbit, hpos, hetsnp, loci = {}, {}, {}, {}
    for c, n, seed in cases.CLI_CONTIGS:
        if block:
            d = synth.generate(n, seed=seed, somatic_rate=2e-5, depth=cases.CLI_DEPTH, phase_block=block)
            ph = synth.phase_table(d.germ, block)
            hbit[c], hpos[c], hetsnp[c] = cases.phase_dicts({"phase": ph})
            loci[c] = [(c, v[0], v[-1]) for v in hpos[c].values()]  # vcflib.py:655-662
        else:
            hbit[c], hpos[c], hetsnp[c] = {}, {}, {}
            loci[c] = [(c, s, e) for s, e in cases.chunkloci(0, n)]

    pool = mp.Pool(threads)
    manager = mp.Manager()
    lst, log = manager.dict(), manager.dict()
    pool.starmap(caller.get_somatic_substitutions, [
        (c, bam, common, pon, loci[c], hbit[c], hpos[c], hetsnp[c], a["min_qv"], a["min_mapq"], qlo, qhi, a["min_sequence_identity"],
         a["min_gq"], a["min_bq"], a["min_trim"], a["max_mismatch_count"], a["mismatch_window"], md, a["min_ref_count"],
         a["min_alt_count"], a["min_hap_count"], 1 / (10 ** 6), a["germline_snv_prior"], 1 / (10 ** 4), bool(block), False, False, lst, log)
        for c in chrom_lst])
    pool.close()
    pool.join()
    out = os.path.join(work, "out.vcf")
    (vcfio.dump_phased_sbs if block else vcfio.dump_sbs)(out, HEADER, chrom_lst, lst)
    body = [l for l in open(out).read().split("\n") if l and not l.startswith("#")]
    print(json.dumps({"thresholds": [int(qlo), int(qhi), int(md)], "log": {c: [int(v) for v in log[c]] for c in chrom_lst}, "body": body}))


def run_normcounts(work, bam, common, pon, threads, over):
    """the process model of `himut normcounts` around the callable-base worker (normcounts.py:494-538): one starmap task
    per contig of the region list, the contig's sequence pickled into the task, Manager dicts for the results; the
    thresholds are the ones the reference read back from the call VCF's header (given as qlo= qhi= md=)"""
    import cases
    from himut_b200 import gtmodel, normcounts
    a = dict(gtmodel.DEFAULT_CALL_ARGS)
    qlo, qhi, md = int(over["qlo"]), int(over["qhi"]), float(over["md"])
    contigs = over["contigs"].split(",")
    data = {c: (n, d) for c, n, d in cases.cli_dataset()}
    pool = mp.Pool(threads)
    manager = mp.Manager()
    ccs, rt, log = manager.dict(), manager.dict(), manager.dict()
    pool.starmap(normcounts.get_callable_tricounts, [
        (c, data[c][1].ref.decode(), bam, common, pon, [(c, s, e) for s, e in cases.chunkloci(0, data[c][0])], {}, {}, {},
         a["min_qv"], a["min_mapq"], a["min_trim"], qlo, qhi, a["min_sequence_identity"], a["min_gq"], a["min_bq"], a["mismatch_window"],
         a["max_mismatch_count"], a["min_ref_count"], a["min_alt_count"], a["min_hap_count"], md, 1 / (10 ** 6), a["germline_snv_prior"],
         1 / (10 ** 4), False, False, ccs, rt, log) for c in contigs])
    pool.close()
    pool.join()
    tri = {}
    for t in normcounts.TRI_LST:  # mutlib.get_cumsum_tricounts: summed over the contigs
        tri[t] = [sum(int(rt[c][t]) for c in contigs), sum(int(ccs[c][t]) for c in contigs)]
    print(json.dumps({"tri": tri, "log": {c: [int(v) for v in log[c]] for c in contigs}}))


def _edges_task(c, n, seed, bam):
    import cases
    from himut_b200 import phaselib, synth
    d = synth.generate(n, seed=seed, somatic_rate=2e-5, depth=cases.CLI_DEPTH)
    het = d.germ["gt"] < 2
    hetsnp_lst = [(int(p), "ATGC"[r], "ATGC"[al]) for p, r, al in zip(d.germ["pos"][het], d.germ["ref"][het], d.germ["alt"][het])]
    h2i = {h: i for i, h in enumerate(hetsnp_lst)}
    edge_lst, e2c = phaselib.get_edges(c, bam, 20, 20, [h[0] for h in hetsnp_lst], hetsnp_lst, h2i)
    return c, [[int(i), int(j)] + [int(v) for v in e2c[(i, j)]] for (i, j) in edge_lst]


def run_phase_edges(bam, threads):
    """`himut phase` forks one task per contig (phaselib.py:280-300), each calling get_edges: the same around the mirror"""
    import cases
    pool = mp.Pool(threads)
    res = pool.starmap(_edges_task, [(c, n, seed, bam) for c, n, seed in cases.CLI_CONTIGS])
    pool.close()
    pool.join()
    print(json.dumps({"edges": dict(res)}))


if __name__ == "__main__":
    main()
