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


_EMITTED_CODES = np.array(sorted(_EMITTED), np.uint8)
_NAME_BY_CODE = [abi.STATUS_NAME.get(c, "") for c in range(256)]


def _sorted_rows(rows):
    """natsorted(rows) for rows of one contig.  All rows share element 0, position keys are ("", int) and the
    ref / alt strings hold no digits, so natsort's order is the plain order of (pos, ref, alt) whenever those
    three are distinct from row to row; otherwise (the later, mixed-type fields decide) the general routine runs."""
    out = sorted(rows, key=lambda r: (r[1], r[2], r[3]))
    for a, b in zip(out, out[1:]):
        if a[1] == b[1] and a[2] == b[2] and a[3] == b[3]:
            return natsorted(rows)
    return out


def records_to_tsbs_lst(chrom, records):
    """chrom2tsbs_lst[chrom] = natsorted(list(set(pass + filtered)))  (caller.py:622-624).
    Column arithmetic is done on whole arrays (the expressions of record_to_tuple, same operand types: int32
    counts as float64, int quality sums divided by float64 counts); HetAltSite rows, which carry formatted
    strings, go through record_to_tuple one by one."""
    rec = records[np.isin(records["status"], _EMITTED_CODES)]
    n = rec.size
    if n == 0:
        return []
    # rows in (pos, ref letter, alt letter) order up front: when no two records share a site the rows are distinct and
    # already in natsort's order (see _sorted_rows), so neither the set nor the sort has anything left to do
    letter = np.frombuffer(abi.CODE2BASE.encode(), np.uint8)
    order = np.lexsort((letter[rec["alt"]], letter[rec["ref"]], rec["tpos"]))
    rec = rec[order]
    same_site = (rec["tpos"][1:] == rec["tpos"][:-1]) & (rec["ref"][1:] == rec["ref"][:-1]) & (rec["alt"][1:] == rec["alt"][:-1])
    distinct = not bool(same_site.any()) and not bool((rec["status"] == abi.ST_HETALT_SITE).any())
    idx = np.arange(n)
    ref_c, alt_c = rec["ref"].astype(np.intp), rec["alt"].astype(np.intp)
    counts = rec["counts"].astype(np.float64)
    depth = counts.sum(axis=1) - counts[:, 4]                 # bamlib.get_read_depth
    n_ref, n_alt = counts[idx, ref_c], counts[idx, alt_c]
    with np.errstate(divide="ignore", invalid="ignore"):
        vaf = n_alt / depth                                   # bamlib.get_alt_counts
        alt_bq = np.where(n_alt != 0, rec["bq_sum"][idx, alt_c].astype(np.int64) / n_alt, 0.0)
    status = rec["status"]
    ps = rec["phase_set"]
    ps_txt = ["."] * n
    for i in np.flatnonzero((status == abi.ST_PASS) & (ps >= 0)).tolist():
        ps_txt[i] = str(int(ps[i]))
    rows = list(zip([chrom] * n, rec["tpos"].tolist(), [abi.CODE2BASE[c] for c in ref_c.tolist()],
                    [abi.CODE2BASE[c] for c in alt_c.tolist()], [_NAME_BY_CODE[c] for c in status.tolist()],
                    rec["gq"].tolist(), alt_bq.tolist(), depth.tolist(), n_ref.tolist(), n_alt.tolist(), vaf.tolist(), ps_txt))
    for i in np.flatnonzero(status == abi.ST_HETALT_SITE).tolist():
        rows[i] = record_to_tuple(chrom, rec[i])
    return rows if distinct else _sorted_rows(list(set(rows)))
