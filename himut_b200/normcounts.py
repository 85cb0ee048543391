"""GPU worker of `himut normcounts` (callable-base half): drop-in for
himut.normcounts.get_callable_tricounts (src/himut/normcounts.py:206-420).

Fills chrom2ccs_callable_tri2count[chrom], chrom2ref_callable_tri2count[chrom] (keys in
mutlib.tri_lst order; contexts containing N are tallied under "NNN", which the reference's
consumers never read, mutlib.py:2384-2393) and chrom2norm_log[chrom] (14 counters).
"""
import numpy as np

from . import abi, gtmodel, vcfio, worker

# mutlib.tri_lst (src/himut/mutlib.py:17-50)
TRI_LST = [a + c + b for a in "ACGT" for c in "CT" for b in "ACGT"]


def get_callable_tricounts(
    chrom, seq, bam_file, common_snps, panel_of_normals, chunkloci_lst, phase_set2hbit_lst, phase_set2hpos_lst,
    phase_set2hetsnp_lst, min_qv, min_mapq, min_trim, qlen_lower_limit, qlen_upper_limit, min_sequence_identity,
    min_gq, min_bq, mismatch_window, max_mismatch_count, min_ref_count, min_alt_count, min_hap_count, md_threshold,
    somatic_snv_prior, germline_snv_prior, germline_indel_prior, phase, non_human_sample,
    chrom2ccs_callable_tri2count, chrom2ref_callable_tri2count, chrom2norm_log,
):
    ctx = worker.context()
    params = gtmodel.make_params(
        min_qv=min_qv, min_mapq=min_mapq, qlen_lower_limit=qlen_lower_limit, qlen_upper_limit=qlen_upper_limit,
        min_sequence_identity=min_sequence_identity, min_gq=min_gq, min_bq=min_bq, min_trim=min_trim,
        max_mismatch_count=max_mismatch_count, mismatch_window=mismatch_window, md_threshold=md_threshold,
        min_ref_count=min_ref_count, min_alt_count=min_alt_count, min_hap_count=min_hap_count,
        germline_snv_prior=germline_snv_prior, phase=phase, non_human_sample=non_human_sample)
    ctx.set_params(params)
    # normcounts loads both sets whenever a file is given (normcounts.py:248-289)
    ctx.set_site_sets(vcfio.load_common_snps(chrom, common_snps), vcfio.load_pon(chrom, panel_of_normals))
    chunk_sets = None
    if phase:
        table, chunk_sets = vcfio.phase_tables(chunkloci_lst, phase_set2hbit_lst, phase_set2hpos_lst, phase_set2hetsnp_lst)
        ctx.set_phase_sets(table)
    refseq = seq.encode() if isinstance(seq, str) else bytes(seq)

    src = worker.RegionSource(bam_file)
    tally = worker.QnameTally()
    ccs = np.zeros(abi.TRI_BINS, np.int64)
    ref = np.zeros(abi.TRI_BINS, np.int64)
    log = np.zeros(abi.NORM_LOG_LEN, np.int64)
    ties = 0
    groups = worker.group_chunks(chunkloci_lst)
    pins = worker.PinCache(ctx, enabled=len(groups) > 1)
    try:
        # qualities in the decoder's compact form, bases as a stream (the tile kernel reads them), decode one group ahead
        for _idx, batch, cq, table, release in worker.pipelined_groups(src, chrom, chunkloci_lst, groups, chunk_sets, seq=True):
            if batch.n_reads == 0:
                release()
                continue
            pins.pin([cq.mask, cq.exc, batch.ops, batch.seq])
            ctx.upload_compact(batch, cq)
            release()
            c, r, l, t = ctx.normcounts_chunks(refseq, table)
            tally.add(ctx.qname_seen())
            ccs += c; ref += r; log += l; ties += t
    finally:
        pins.close()
    src.close()
    log[0] = tally.count()
    ccs_d = {tri: int(ccs[i]) for i, tri in enumerate(TRI_LST)}
    ref_d = {tri: int(ref[i]) for i, tri in enumerate(TRI_LST)}
    if ccs[32] or ref[32]:
        ccs_d["NNN"], ref_d["NNN"] = int(ccs[32]), int(ref[32])
    chrom2ccs_callable_tri2count[chrom] = ccs_d
    chrom2ref_callable_tri2count[chrom] = ref_d
    chrom2norm_log[chrom] = [int(v) for v in log]
    return ties
