"""Known-answer vectors of SURVEY.md Appendix B (produced from the reference's own functions),
checked against the host-side restatements the product uses.  CPU only."""
import math

import numpy as np

import cases
from himut_b200 import abi, gtmodel, pack
from oracle import oracle


def test_priors_and_bq_terms():
    pri = gtmodel.germline_priors(1e-3)
    assert pri == [1 - ((1.5 * 1e-3) + (1e-3 * 1e-3)), 1e-3, 2e-06, 0.0005]
    assert abs(pri[0] - 0.998499) < 1e-12
    hom, het, err = gtmodel.bq_tables()
    assert (hom[1], het[1], err[1]) == (-0.6868253243801155, -0.9878553200440967, -0.03333333333333334)
    assert (hom[20], het[20], err[20]) == (-0.004364805402450088, -0.3053948010664313, -0.6666666666666667)
    assert (hom[40], het[40], err[40]) == (-4.3431619807505604e-05, -0.3010734272837887, -1.3333333333333335)
    assert (hom[93], het[93], err[93]) == (-2.1766283665300684e-10, -0.30102999588164403, -3.1)


def test_cs_kat():
    ops, rspan, qspan = pack.parse_cs(":4*ag:2+tt:1-ca:3", "NNACGTACGTTTGACCA", 2)
    mk, sub = abi.make_op, abi.make_sub
    assert ops == [mk(0, 4), sub(0, 2), mk(0, 2), mk(2, 2), mk(0, 1), mk(3, 2), mk(0, 3)]
    assert (rspan, qspan) == (13, 13)
    # long form and upper/lower case are accepted, mismatching long-form bases are not
    assert pack.parse_cs("=ACGT*ag", "ACGTG", 0)[0] == [mk(0, 4), sub(0, 2)]
    try:
        pack.parse_cs("=ACGA", "ACGT", 0)
        assert False
    except pack.BatchFormatError:
        pass
    try:
        pack.parse_cs(":4~gt12ag:3", "ACGTACG", 0)
        assert False
    except pack.BatchFormatError:
        pass


def _one_read_batch(tstart, qstart, qseq, bq, cs):
    bb = pack.BatchBuilder()
    bb.add(tstart=tstart, tend=None, qstart=qstart, qend=None, qseq=qseq, bq=bytes(bq), mapq=60,
           is_secondary=False, qname="q", cs=cs)
    return bb.finish()


def test_read_stats_kat():
    b = _one_read_batch(100, 2, "NNACGTACGTTTGACCA", [93] * 17, ":4*ag:2+tt:1-ca:3")
    s = oracle.read_stats(b)
    assert int(s["n_match"][0]) == 10 and int(s["n_sub"][0]) == 1 and int(s["ins_len"][0]) == 2 and int(s["del_len"][0]) == 2
    assert s["n_match"][0] / float(s["n_match"][0] + 5) == 0.6666666666666666
    assert int(s["n_mismatch"][0]) == 3 and int(s["bq_total"][0]) == 93 * 17
    assert int(b.tend[0]) == 113


def test_chunkloci_kat():
    assert cases.chunkloci(0, 1000000) == [(1, 200000), (200000, 400000), (400000, 600000), (600000, 800000), (800000, 999998)]
    assert cases.chunkloci(0, 150000) == [(0, 150000)]
    assert cases.chunkloci(0, 450001) == [(1, 200000), (200000, 400000), (400000, 449999)]


def _site_batch(ref_base, lists):
    """reads of length 1 at position 10, one per (allele, bq) entry, in order"""
    bb = pack.BatchBuilder()
    i = 0
    for allele, bqs in lists:
        for q in bqs:
            cs = ":1" if allele == ref_base else "*%s%s" % (ref_base.lower(), allele.lower())
            bb.add(tstart=10, tend=None, qstart=0, qend=None, qseq=allele, bq=bytes([q]), mapq=60,
                   is_secondary=False, qname="q%d" % i, cs=cs)
            i += 1
    return bb.finish()


def _eval(b, alt_ok=True):
    args = cases.call_args(qlen_lower_limit=0, qlen_upper_limit=10, min_trim=0.0, min_bq=1, min_gq=0, min_ref_count=0,
                           min_sequence_identity=0.0, min_qv=0)
    p = gtmodel.make_params(**args)
    rec, log = oracle.call_chunks(p, b, b.chunk_table([(0, 100)]))
    return rec


def test_genotype_kats():
    # KAT1: A x29 @93, T x1 @93 -> ("AA", 89, homref); call-form germ_gq 89
    rec = _eval(_site_batch("A", [("A", [93] * 29), ("T", [93])]))
    assert rec.size == 1 and int(rec["gq"][0]) == 89 and int(rec["germ_state"][0]) == 0
    assert list(rec["germ_gt"][0]) == [0, 0]
    # KAT3: A [93,93,93], C [93] -> ("AA", 11, homref)
    rec = _eval(_site_batch("A", [("A", [93] * 3), ("C", [93])]))
    assert int(rec["gq"][0]) == 11 and int(rec["germ_state"][0]) == 0
    # KAT2: A [93,40,30]x5, G [93,20,50]x5 -> ("AG", 99, het): restates the germline genotype
    rec = _eval(_site_batch("A", [("A", [93, 40, 30] * 5), ("G", [93, 20, 50] * 5)]))
    assert int(rec["status"][0]) == abi.ST_GERM_HET and int(rec["gq"][0]) == 99
    assert list(rec["germ_gt"][0]) == [0, 2]


def test_trim_and_window_kats():
    # get_trimmed_range(15000, 0.01) = (150, 14850); (17, 0.01) = (0, 17)
    assert (math.floor(0.01 * 15000), math.ceil((1 - 0.01) * 15000)) == (150, 14850)
    assert (math.floor(0.01 * 17), math.ceil((1 - 0.01) * 17)) == (0, 17)
    assert math.ceil(30 + 4 * math.sqrt(30)) == 52  # get_md_threshold(30)
