"""Test-only glue: run the UNMODIFIED reference (/root/reference/src/himut) in this container.

The reference needs pysam / natsort / tabix / cyvcf2 / pyfastx / plotnine, none of which are
installed; tests/shims/ provides import-compatible stand-ins for the I/O they do, and this
module serves packed ReadBatch reads through the pysam look-alike.  /root/reference exists
only in the build container, never on the GPU box: callers must check have_reference().
"""
import array
import os
import sys

import numpy as np

from himut_b200 import abi

REF_SRC = "/root/reference/src"
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def have_reference():
    return os.path.isdir(os.path.join(REF_SRC, "himut"))


def import_reference():
    """-> the reference's `himut` package, imported unmodified through the shims"""
    if not have_reference():
        raise RuntimeError("reference sources are not present")
    for p in (REF_SRC, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import himut.caller  # noqa: F401
        import himut.normcounts  # noqa: F401
    import himut
    return himut


class Segment:
    """AlignedSegment look-alike (the attribute surface of SURVEY.md §8b)"""
    __slots__ = ("is_secondary", "reference_name", "reference_start", "reference_end", "query_name",
                 "query_alignment_start", "query_alignment_end", "query_sequence", "mapping_quality",
                 "query_qualities", "_tags")

    def get_tag(self, t):
        return self._tags[t]

    def has_tag(self, t):
        return t in self._tags


def ops_to_cs(ops, qseq, qstart):
    """cs:Z short form from the packed op stream (deleted bases are not stored: 'n')"""
    out, q = [], qstart
    for w in ops:
        kind, val = int(w) & 3, int(w) >> 2
        if kind == abi.OP_MATCH:
            out.append(":%d" % val); q += val
        elif kind == abi.OP_SUB:
            out.append("*%s%s" % ("atgcn"[val & 7], "atgcn"[(val >> 3) & 7])); q += 1
        elif kind == abi.OP_INS:
            out.append("+" + qseq[q:q + val].lower()); q += val
        else:
            out.append("-" + "n" * val)
    return "".join(out), q


class BatchProvider:
    """serves one contig's ReadBatch as pysam records"""

    def __init__(self, chrom, contig_len, batch, sample="synth"):
        self.chrom, self.batch = chrom, batch
        self.header_text = "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:%s\tLN:%d\n@RG\tID:rg\tSM:%s\n" % (chrom, contig_len, sample)
        self._cache = {}

    def segment(self, r):
        if r in self._cache:
            return self._cache[r]
        b = self.batch
        s = Segment()
        ql = int(b.qlen[r])
        so, bo, oo = int(b.seq_off[r]), int(b.bq_off[r]), int(b.op_off[r])
        packed = b.seq[so:so + (ql + 3) // 4]
        codes = ((packed[:, None] >> np.array([0, 2, 4, 6], np.uint8)) & 3).reshape(-1)[:ql]
        qseq = "".join("ATGC"[c] for c in codes)
        s.is_secondary = bool(b.flags[r] & abi.READ_SECONDARY)
        s.reference_name = self.chrom
        s.reference_start = int(b.tstart[r])
        s.reference_end = int(b.tend[r])
        s.query_name = "read%d" % int(b.qname_id[r])
        s.query_alignment_start = int(b.qstart[r])
        cs, qend = ops_to_cs(b.ops[oo:oo + int(b.n_ops[r])], qseq, int(b.qstart[r]))
        s.query_alignment_end = qend
        s.query_sequence = qseq
        s.mapping_quality = int(b.mapq[r])
        s.query_qualities = array.array("B", b.bq[bo:bo + ql].tobytes())
        s._tags = {"cs": cs, "tp": "P"}
        if len(self._cache) > 4096:
            self._cache.clear()
        self._cache[r] = s
        return s

    def fetch_records(self, chrom=None, start=None, end=None):
        b = self.batch
        if chrom is not None and chrom != self.chrom:
            return
        for r in range(b.n_reads):
            if start is not None and not (b.tstart[r] < end and b.tend[r] > start):
                if b.tstart[r] >= end:
                    break
                continue
            yield self.segment(r)


class ContigsProvider:
    """several contigs' batches behind one pysam.AlignmentFile look-alike"""

    def __init__(self, contigs, sample="synth"):
        """contigs: [(name, length, ReadBatch)] in @SQ order"""
        self.header_text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (c, n) for c, n, _ in contigs) \
                           + "@RG\tID:rg\tSM:%s\n" % sample
        self._by_name = {c: BatchProvider(c, n, b, sample) for c, n, b in contigs}
        self._order = [c for c, _, _ in contigs]

    def fetch_records(self, chrom=None, start=None, end=None):
        for c in ([chrom] if chrom is not None else self._order):
            if c in self._by_name:
                yield from self._by_name[c].fetch_records(c, start, end)


class StableArgsortNumpy:
    """numpy as the reference's gtlib should see it: np.argsort of a 10-vector is an insertion sort, hence stable, in the
    reference's pinned numpy 1.24.4 (poetry.lock), while numpy >= 1.25 on AVX-512 hosts dispatches to an unstable SIMD
    sort.  Install with `himut.gtlib.np = StableArgsortNumpy()`."""

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def argsort(a, *args, **kw):
        return np.argsort(a, kind="stable")
