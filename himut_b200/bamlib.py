"""BAM pre-pass of `himut call`: drop-in for himut.bamlib.get_thresholds
(src/himut/bamlib.py:137-178).

Same sampling as the reference — random.seed(10), 100 windows of 100 kb per contig drawn with
random.sample(range(chrom_len), 100) — and the same numpy / math expressions on the same list of
read lengths in the same order, so the three thresholds (and therefore the VCF header and the
read / depth filters they feed) are identical.  What changes is how the list is obtained: the
native decoder (csrc/bamdec.c: hm_bam_window_qlens) inflates the window's BGZF blocks on a thread
pool and looks only at record headers, CIGARs and the tp tag; no Python object per record.
"""
import math
import random

import numpy as np

from . import bamdec


def get_md_threshold(coverage):
    return math.ceil(coverage + (4 * math.sqrt(coverage)))


def get_thresholds(bam_file, chrom_lst, chrom2len):
    if len(chrom_lst) == 0:
        print("target is missing")
        print("Please check .vcf file or .target file")
        import himut.util
        himut.util.exit()

    parts = []
    random.seed(10)
    sample_count = 100
    sample_range = 100000
    genome_sample_sum = sample_count * sample_range * len(chrom_lst)
    bam = bamdec.NativeBam(bam_file)
    for chrom in chrom_lst:
        chrom_len = chrom2len[chrom]
        for start in random.sample(range(chrom_len), sample_count):
            parts.append(bam.window_qlens(chrom, start, start + 100000))
    bam.close()
    qlen_lst = np.concatenate(parts).astype(np.int64).tolist() if parts else []
    genome_read_sum = sum(qlen_lst)
    qlen_std = np.std(qlen_lst)
    qlen_mean = math.ceil(np.mean(qlen_lst))
    qlen_lower_limit = 0 if math.ceil(qlen_mean - 2 * qlen_std) < 0 else math.ceil(qlen_mean - 2 * qlen_std)
    qlen_upper_limit = math.ceil(qlen_mean + 2 * qlen_std)
    coverage = genome_read_sum / float(genome_sample_sum)
    md_threshold = get_md_threshold(coverage)
    return qlen_lower_limit, qlen_upper_limit, md_threshold
