"""hm_site_record[] -> the 12-tuples the reference worker emits (caller.py:349-621).

The library returns integers only; strings, means and fractions are derived here with the
reference's own expressions so vcflib.dump_sbs / dump_phased_sbs format them unchanged.
"""
import numpy as np

from . import abi
from .natsort_compat import natsorted

_EMITTED = set(abi.STATUS_NAME)


def record_to_tuple(chrom, r):
    """one emitted row; field types follow the reference (np.float64 counts, int gq)"""
    ref = abi.CODE2BASE[int(r["ref"])]
    alt = abi.CODE2BASE[int(r["alt"])]
    counts = r["counts"].astype(np.float64)
    ins_count = counts[4]
    read_depth = sum(counts) - ins_count                      # bamlib.get_read_depth
    ref_count = counts[int(r["ref"])]
    alt_count = counts[int(r["alt"])]
    alt_vaf = alt_count / float(read_depth)                   # bamlib.get_alt_counts
    alt_bq = int(r["bq_sum"][int(r["alt"])]) / float(alt_count) if alt_count != 0 else 0.0
    status = int(r["status"])
    phase_set = "."
    if status == abi.ST_HETALT_SITE:                          # caller.get_hetalt_counts
        a1, a2 = int(r["germ_gt"][0]), int(r["germ_gt"][1])
        alt = "{},{}".format(abi.CODE2BASE[a1], abi.CODE2BASE[a2])
        p_count, q_count = counts[a1], counts[a2]
        pbq = int(r["bq_sum"][a1]) / float(p_count)
        qbq = int(r["bq_sum"][a2]) / float(q_count)
        alt_bq = "{:0.1f},{:0.1f}".format(pbq, qbq)
        alt_count = "{:0.0f},{:0.0f}".format(p_count, q_count)
        alt_vaf = "{:.2f},{:.2f}".format(p_count / float(read_depth), q_count / float(read_depth))
    elif status == abi.ST_PASS and int(r["phase_set"]) >= 0:
        phase_set = str(int(r["phase_set"]))
    return (chrom, int(r["tpos"]), ref, alt, abi.STATUS_NAME[status], int(r["gq"]), alt_bq,
            read_depth, ref_count, alt_count, alt_vaf, phase_set)


def records_to_tsbs_lst(chrom, records):
    """chrom2tsbs_lst[chrom] = natsorted(list(set(pass + filtered)))  (caller.py:622-624)"""
    rows = [record_to_tuple(chrom, r) for r in records if int(r["status"]) in _EMITTED]
    return natsorted(list(set(rows)))
