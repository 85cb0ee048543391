"""GPU worker of `himut phase`'s edge counting: drop-in for himut.phaselib.get_edges
(src/himut/phaselib.py:16-67).

Same arguments, same return value: (natsorted list of (i, j) edges, {(i, j): np.array of the four
cis / trans counts}).  The BAM is decoded natively window by window (a read that two windows share
is counted with the window it starts in), the pair tables are accumulated on the device in a band
(hm_phase_edges_*), and only non-empty tables become dictionary entries — the reference creates an
entry when it first increments it.  The binomial tests and the BFS over the graph stay the
reference's own code (phaselib.py:70-195).
"""
from collections import defaultdict

import numpy as np

from . import natsort_compat, worker

BASE2CODE = {"A": 0, "T": 1, "G": 2, "C": 3}


def get_edges(chrom, bam_file, min_bq, min_mapq, hpos_lst, hetsnp_lst, hetsnp2hidx):
    try:
        from natsort import natsorted
    except ImportError:  # pragma: no cover - natsort is a dependency of the reference
        natsorted = natsort_compat.natsorted
    edge2counts = defaultdict(lambda: np.zeros(4))
    n = len(hpos_lst)
    if n >= 2:
        ctx = worker.context()
        hpos = np.asarray(hpos_lst, np.int32)
        href = np.array([BASE2CODE.get(h[1], 4) for h in hetsnp_lst], np.uint8)
        hidx = np.array([hetsnp2hidx[h] for h in hetsnp_lst], np.int64)
        src = worker.RegionSource(bam_file)
        length = src.reader.lengths[src.reader.references.index(chrom)]
        band = 64
        while True:
            ctx.phase_edges_begin(hpos, href, band)
            need, lo = 0, 0
            while lo < length and not need:
                hi = min(length, lo + worker.GROUP_SPAN)
                batch, _ = src.batch(chrom, [(chrom, lo, hi)], seq=False)  # a cs match at a hetSNP carries its reference allele
                if batch.n_reads:
                    ctx.upload(batch)
                    need = ctx.phase_edges_add(min_bq, min_mapq, lo if lo else -2**31)
                lo = hi
            if not need:
                break
            band = max(need, band * 2)  # a read paired hetSNPs further apart than the band: start over, wider
        table = ctx.phase_edges_end()
        src.close()
        a_idx, d_idx = np.nonzero(table.any(axis=2))
        for a, d in zip(a_idx.tolist(), d_idx.tolist()):
            edge2counts[(int(hidx[a]), int(hidx[a + d + 1]))] += table[a, d].astype(np.float64)
    edge_lst = natsorted(list(edge2counts.keys()))
    return edge_lst, edge2counts
