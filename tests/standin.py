"""Test infrastructure: a stand-in for lib.Context whose device calls are answered by the CPU oracle, so the host logic
of the worker mirrors (and the whole reference CLI with the mirrors patched in) can be pinned without a GPU.  Never
imported by the product package."""
import numpy as np

from oracle import oracle


class OracleContext:
    """the part of lib.Context the workers use, answered by the oracle"""

    def __init__(self):
        self.params = self.batch = self.phase = None
        self.common = self.pon = None
        self.uploads = 0

    def set_params(self, p):
        self.params = p

    def set_site_sets(self, common=None, pon=None):
        self.common, self.pon = common, pon

    def set_phase_sets(self, table):
        self.phase = table

    def upload(self, batch):
        assert batch.seq.size, "the oracle reads the bases: the test asks the decoder for them"
        self.batch = batch
        self.uploads += 1

    def upload_compact(self, batch, cq):
        """qualities as the decoder's parse pass leaves them (bitmap + exceptions): expanded here as k_bq_expand does on
        the device, then the oracle sees the same one-byte-per-base stream"""
        from himut_b200 import abi
        bits = np.unpackbits(cq.mask, bitorder="little").astype(bool)
        valid = np.zeros(bits.size, bool)
        for o, n in zip(batch.bq_off.tolist(), batch.qlen.tolist()):
            valid[o:o + n] = True
        bq = np.where(bits & valid, np.uint8(cq.modal), np.uint8(0)).astype(np.uint8)
        zero = ~bits & valid
        assert int(zero.sum()) == cq.n_exc
        bq[zero] = cq.exc[:cq.n_exc]  # exceptions are stored read by read, base by base: the stream's own order
        # copies: like the device upload, nothing may point into the decoder's buffers once this returns (the worker
        # hands them back to the decode thread right after)
        kw = {name: np.array(getattr(batch, name)) for name, _ in abi.ReadBatch._FIELDS}
        kw["bq"] = bq
        self.upload(abi.ReadBatch(**kw))

    def pin_arrays(self, arrays):
        pass

    def unpin_arrays(self, arrays):
        pass

    def call_chunks(self, table):
        self._seen = np.zeros(int(self.batch.qname_id.max()) + 1 if self.batch.n_reads else 1, np.uint8)
        return oracle.call_chunks(self.params, self.batch, table, self.common, self.pon, self.phase, qseen=self._seen)

    def qname_seen(self):
        return self._seen

    def normcounts_chunks(self, refseq, table):
        self._seen = np.zeros(int(self.batch.qname_id.max()) + 1 if self.batch.n_reads else 1, np.uint8)
        ref = refseq.encode() if isinstance(refseq, str) else refseq
        return oracle.normcounts_chunks(self.params, self.batch, ref, table, self.common, self.pon, self.phase, qseen=self._seen)

    def ref_tricounts(self, refseq):
        return oracle.ref_tricounts(refseq.encode() if isinstance(refseq, str) else bytes(refseq))

    def phase_edges_begin(self, hpos, href, band):
        self._edges = (np.ascontiguousarray(hpos, np.int32), np.ascontiguousarray(href, np.uint8), int(band))
        self._table = np.zeros((len(hpos), int(band), 4), np.uint32)

    def phase_edges_add(self, min_bq, min_mapq, min_tstart=-2**31):
        hpos, href, band = self._edges
        counts, need = oracle.phase_edges(self.batch, hpos, href, band, min_bq, min_mapq, min_tstart)
        if need:
            return int(need)
        self._table += counts
        return 0

    def phase_edges_end(self):
        return self._table
