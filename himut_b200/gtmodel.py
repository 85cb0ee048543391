"""Host-side constants of the germline genotype model and the hm_params builder.

The device never evaluates log10 / pow: the three per-BQ terms and the four log10 priors of
the reference's gtlib (src/himut/gtlib.py:12-20,47-69) are tabulated here in Python with the
same expressions, so the fp64 values the kernels add are the reference's own, bit for bit.
"""
import math

from . import abi


def germline_priors(germline_snv_prior):
    """gtlib.init (src/himut/gtlib.py:12-20), order homref, het, hetalt, homalt"""
    p = germline_snv_prior
    het = p
    hetalt = p * p * 2
    homref = 1 - ((1.5 * p) + (p * p))
    homalt = p / 2
    return [homref, het, hetalt, homalt]


def bq_tables():
    """(lut_hom, lut_het, lut_err) for bq 0..255; entry 0 is nan (the reference raises on BQ 0)

    gtlib.get_log10_one_minus_epsilon / get_log10_one_half_minus_epsilon /
    get_log10_epsilon(bq / 3)  (src/himut/gtlib.py:47-69)
    """
    hom, het, err = [math.nan], [math.nan], [math.nan]
    for bq in range(1, 256):
        eps = 10 ** (-bq / 10)
        hom.append(math.log10(1 - eps))
        het.append(math.log10(0.5 - eps / 2.0))
        err.append(math.log10(10 ** (-(bq / 3) / 10)))
    return hom, het, err


_TABLES = None


def make_params(*, min_qv, min_mapq, qlen_lower_limit, qlen_upper_limit, min_sequence_identity,
                min_gq, min_bq, min_trim, max_mismatch_count, mismatch_window, md_threshold,
                min_ref_count, min_alt_count, min_hap_count, germline_snv_prior, phase=False,
                non_human_sample=False, create_panel_of_normals=False):
    """hm_params from the worker arguments (caller.py:208-241 / normcounts.py:206-238)"""
    global _TABLES
    if _TABLES is None:
        _TABLES = bq_tables()
    p = abi.hm_params()
    p.min_qv = int(min_qv)
    p.min_mapq = int(min_mapq)
    p.qlen_lower_limit = int(qlen_lower_limit)
    p.qlen_upper_limit = int(qlen_upper_limit)
    p.min_gq = int(min_gq)
    p.min_bq = int(min_bq)
    p.max_mismatch_count = int(max_mismatch_count)
    p.mismatch_window = int(mismatch_window)
    p.min_ref_count = int(min_ref_count)
    p.min_alt_count = int(min_alt_count)
    p.min_hap_count = int(min_hap_count)
    p.phase = int(bool(phase))
    p.non_human_sample = int(bool(non_human_sample))
    p.create_panel_of_normals = int(bool(create_panel_of_normals))
    p.min_sequence_identity = float(min_sequence_identity)
    p.min_trim = float(min_trim)
    p.md_threshold = float(md_threshold)
    for i, pr in enumerate(germline_priors(germline_snv_prior)):
        p.log10_prior[i] = math.log10(pr)
    hom, het, err = _TABLES
    for i in range(256):
        p.lut_hom[i] = hom[i]
        p.lut_het[i] = het[i]
        p.lut_err[i] = err[i]
    return p


# himut call defaults (src/himut/parse_args.py:91-196) with thresholds a 30x run produces
DEFAULT_CALL_ARGS = dict(
    min_qv=30, min_mapq=60, qlen_lower_limit=10000, qlen_upper_limit=20000,
    min_sequence_identity=0.99, min_gq=20, min_bq=93, min_trim=0.01, max_mismatch_count=0,
    mismatch_window=20, md_threshold=52, min_ref_count=3, min_alt_count=1, min_hap_count=3,
    germline_snv_prior=1e-3)
