"""ctypes wrapper of the synthetic CCS data generator (himut_b200/csrc/synth.c).

There is no network and no dataset here, so tests and bench.py run on a random reference
with simulated ~15 kb, 30x HiFi reads carrying minimap2-style cs ops.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class hm_synth_spec(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("contig_len", C.c_int32), ("read_len_min", C.c_int32),
        ("read_len_max", C.c_int32), ("phase_block", C.c_int32), ("depth", C.c_double),
        ("read_len_mean", C.c_double), ("read_len_sd", C.c_double), ("het_rate", C.c_double),
        ("hom_rate", C.c_double), ("somatic_rate", C.c_double), ("sub_err_rate", C.c_double),
        ("indel_rate", C.c_double), ("lowq_frac", C.c_double), ("mapq_low_frac", C.c_double),
        ("softclip_frac", C.c_double),
    ]


class hm_synth_data(C.Structure):
    _fields_ = [
        ("batch", abi.hm_read_batch),
        ("ref", C.c_void_p), ("ref_len", C.c_uint64),
        ("n_germ", C.c_uint64), ("germ_pos", C.c_void_p), ("germ_ref", C.c_void_p),
        ("germ_alt", C.c_void_p), ("germ_gt", C.c_void_p),
        ("n_som", C.c_uint64), ("som_pos", C.c_void_p), ("som_ref", C.c_void_p), ("som_alt", C.c_void_p),
        ("n_err", C.c_uint64), ("err_pos", C.c_void_p), ("err_ref", C.c_void_p), ("err_alt", C.c_void_p),
        ("aligned_bases", C.c_uint64),
    ]


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libhimut_synth.so")
        if not os.path.exists(path):
            raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
        _LIB = C.CDLL(path)
        _LIB.hm_synth_default_spec.argtypes = [C.POINTER(hm_synth_spec)]
        _LIB.hm_synth_generate.argtypes = [C.POINTER(hm_synth_spec), C.POINTER(C.POINTER(hm_synth_data))]
        _LIB.hm_synth_free.argtypes = [C.POINTER(hm_synth_data)]
        _LIB.hm_synth_free.restype = None
    return _LIB


def _view(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype)
    buf = (C.c_char * (int(n) * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(n))


class SynthData:
    """owns one generated contig; arrays are copied out so the C storage can be freed"""

    def __init__(self, batch, ref, germ, som, err, aligned_bases, spec):
        self.batch, self.ref, self.germ, self.som, self.err = batch, ref, germ, som, err
        self.aligned_bases, self.spec = aligned_bases, spec


def generate(contig_len=1_000_000, seed=20260101, copy=True, **overrides):
    """returns SynthData: .batch (abi.ReadBatch), .ref (bytes), .germ / .som / .err site arrays"""
    lib = _lib()
    spec = hm_synth_spec()
    lib.hm_synth_default_spec(C.byref(spec))
    spec.contig_len = int(contig_len)
    spec.seed = int(seed)
    for k, v in overrides.items():
        if not hasattr(spec, k):
            raise TypeError("unknown synth option %r" % k)
        setattr(spec, k, v)
    out = C.POINTER(hm_synth_data)()
    rc = lib.hm_synth_generate(C.byref(spec), C.byref(out))
    if rc != 0:
        raise RuntimeError("hm_synth_generate failed (%d)" % rc)
    d = out.contents
    b = d.batch
    n = int(b.n_reads)
    cp = (lambda a: a.copy()) if copy else (lambda a: a)
    arrays = dict(
        tstart=cp(_view(b.tstart, n, np.int32)), tend=cp(_view(b.tend, n, np.int32)),
        qstart=cp(_view(b.qstart, n, np.int32)), qlen=cp(_view(b.qlen, n, np.int32)),
        mapq=cp(_view(b.mapq, n, np.uint8)), flags=cp(_view(b.flags, n, np.uint8)),
        qname_id=cp(_view(b.qname_id, n, np.uint32)), seq_off=cp(_view(b.seq_off, n, np.uint64)),
        bq_off=cp(_view(b.bq_off, n, np.uint64)), op_off=cp(_view(b.op_off, n, np.uint64)),
        n_ops=cp(_view(b.n_ops, n, np.uint32)), seq=cp(_view(b.seq, b.seq_bytes, np.uint8)),
        bq=cp(_view(b.bq, b.bq_bytes, np.uint8)), ops=cp(_view(b.ops, b.n_ops_total, np.uint32)))
    sites = lambda pos, ref, alt, k: dict(pos=_view(pos, k, np.int32).copy(), ref=_view(ref, k, np.uint8).copy(),
                                          alt=_view(alt, k, np.uint8).copy())
    germ = sites(d.germ_pos, d.germ_ref, d.germ_alt, d.n_germ)
    germ["gt"] = _view(d.germ_gt, d.n_germ, np.uint8).copy()
    som = sites(d.som_pos, d.som_ref, d.som_alt, d.n_som)
    err = sites(d.err_pos, d.err_ref, d.err_alt, d.n_err)
    ref = bytes(_view(d.ref, d.ref_len, np.uint8))
    aligned = int(d.aligned_bases)
    if copy:
        lib.hm_synth_free(out)
        keep = None
    else:
        keep = _Owner(lib, out)
    batch = abi.ReadBatch(keepalive=keep, **arrays)
    return SynthData(batch, ref, germ, som, err, aligned, spec)


class _Owner:
    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        try:
            self.lib.hm_synth_free(self.ptr)
        except Exception:
            pass


def site_keys(pos, ref, alt):
    """sorted uint64 keys pos << 4 | ref << 2 | alt for hm_set_site_sets"""
    k = (pos.astype(np.uint64) << np.uint64(4)) | (ref.astype(np.uint64) << np.uint64(2)) | alt.astype(np.uint64)
    return np.unique(k)


def phase_table(germ, phase_block):
    """phased hetSNP table (hpos, href, halt, hbit, set_off, set_start) from the synthetic germline:
    one phase set per `phase_block` span that holds >= 1 het SNV; hbit = h0 bit ("1|0" -> 1)."""
    het = germ["gt"] < 2
    pos, ref, alt = germ["pos"][het], germ["ref"][het], germ["alt"][het]
    hbit = (germ["gt"][het] == 0).astype(np.uint8)  # gt 0: alt on hap 0 -> "1|0" -> h0 bit 1
    block = (pos - 1) // phase_block
    starts = np.flatnonzero(np.r_[True, block[1:] != block[:-1]]) if pos.size else np.zeros(0, np.int64)
    set_off = np.r_[starts, pos.size].astype(np.uint64)
    return dict(hpos=pos.astype(np.int32), href=ref.astype(np.uint8), halt=alt.astype(np.uint8),
                hbit=hbit, set_off=set_off)
