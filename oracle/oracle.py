"""ctypes wrapper of the CPU oracle (oracle/himut_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (himut_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from himut_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class orc_phase(C.Structure):
    _fields_ = [("hpos", C.c_void_p), ("href", C.c_void_p), ("halt", C.c_void_p), ("hbit", C.c_void_p),
                ("n_hetsnp", C.c_uint64), ("set_off", C.c_void_p), ("n_sets", C.c_uint64)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libhimut_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def _phase_struct(phase):
    if phase is None:
        return None, None
    keep = {k: np.ascontiguousarray(phase[k]) for k in ("hpos", "href", "halt", "hbit", "set_off")}
    s = orc_phase(_p(keep["hpos"]), _p(keep["href"]), _p(keep["halt"]), _p(keep["hbit"]),
                  keep["hpos"].size, _p(keep["set_off"]), keep["set_off"].size - 1)
    return s, keep


def read_stats(batch):
    n = batch.n_reads
    out = dict(bq_total=np.zeros(n, np.int64), n_match=np.zeros(n, np.int32), n_sub=np.zeros(n, np.int32),
               ins_len=np.zeros(n, np.int32), del_len=np.zeros(n, np.int32), n_mismatch=np.zeros(n, np.int32))
    rc = lib().orc_read_stats(C.byref(batch.as_struct()), *[_p(out[k]) for k in
                              ("bq_total", "n_match", "n_sub", "ins_len", "del_len", "n_mismatch")])
    assert rc == 0
    return out


def call_chunks(params, batch, chunks, common=None, pon=None, phase=None, cap=None, qseen=None):
    """-> (records: structured array abi.SITE_DTYPE, log[15]) ; raises on error.
    qseen: optional uint8 array, one flag per qname_id, filled with the query names that passed the read gates"""
    common = np.zeros(0, np.uint64) if common is None else np.ascontiguousarray(common, np.uint64)
    pon = np.zeros(0, np.uint64) if pon is None else np.ascontiguousarray(pon, np.uint64)
    chunks = np.ascontiguousarray(chunks, dtype=abi.CHUNK_DTYPE)
    ph, keep = _phase_struct(phase)
    cap = cap or max(4096, int(batch.ops.size))
    out = np.zeros(cap, dtype=abi.SITE_DTYPE)
    n_out = C.c_size_t(0)
    log = np.zeros(abi.CALL_LOG_LEN, np.int64)
    rc = lib().orc_call_chunks_seen(
        C.byref(params), C.byref(batch.as_struct()), _p(chunks), C.c_size_t(len(chunks)),
        _p(common), C.c_size_t(common.size), _p(pon), C.c_size_t(pon.size),
        C.byref(ph) if ph is not None else None, _p(out), C.c_size_t(cap), C.byref(n_out), _p(log),
        None if qseen is None else _p(qseen), C.c_size_t(0 if qseen is None else qseen.size))
    if rc != 0:
        raise RuntimeError("orc_call_chunks failed: %d" % rc)
    return out[: n_out.value].copy(), log


def normcounts_chunks(params, batch, refseq, chunks, common=None, pon=None, phase=None, alt_order=None, qseen=None):
    """-> (ccs_tri[33], ref_tri[33], log[14], n_alt_tie)"""
    common = np.zeros(0, np.uint64) if common is None else np.ascontiguousarray(common, np.uint64)
    pon = np.zeros(0, np.uint64) if pon is None else np.ascontiguousarray(pon, np.uint64)
    chunks = np.ascontiguousarray(chunks, dtype=abi.CHUNK_DTYPE)
    ph, keep = _phase_struct(phase)
    ref = np.frombuffer(refseq, dtype=np.uint8)
    ao = None if alt_order is None else np.ascontiguousarray(alt_order, np.uint8)
    ccs = np.zeros(abi.TRI_BINS, np.int64)
    rt = np.zeros(abi.TRI_BINS, np.int64)
    log = np.zeros(abi.NORM_LOG_LEN, np.int64)
    ties = C.c_int64(0)
    rc = lib().orc_normcounts_chunks_seen(
        C.byref(params), C.byref(batch.as_struct()), _p(ref), C.c_size_t(ref.size), _p(chunks),
        C.c_size_t(len(chunks)), _p(common), C.c_size_t(common.size), _p(pon), C.c_size_t(pon.size),
        C.byref(ph) if ph is not None else None, _p(ao), _p(ccs), _p(rt), _p(log), C.byref(ties),
        None if qseen is None else _p(qseen), C.c_size_t(0 if qseen is None else qseen.size))
    if rc != 0:
        raise RuntimeError("orc_normcounts_chunks failed: %d" % rc)
    return ccs, rt, log, ties.value


def ref_tricounts(refseq):
    """-> int64[33] (reflib.get_chrom_tricount)"""
    ref = np.frombuffer(refseq, dtype=np.uint8)
    out = np.zeros(abi.TRI_BINS, np.int64)
    rc = lib().orc_ref_tricounts(_p(ref), C.c_size_t(ref.size), _p(out))
    assert rc == 0
    return out


def phase_edges(batch, hpos, href, band, min_bq, min_mapq, min_tstart=-2**31):
    """-> (uint32[n, band, 4], needed_band) (phaselib.get_edges in the band layout of hm_phase_edges_*)"""
    hpos = np.ascontiguousarray(hpos, np.int32)
    href = np.ascontiguousarray(href, np.uint8)
    out = np.zeros((hpos.size, band, 4), np.uint32)
    f = lib().orc_phase_edges
    f.restype = C.c_uint32
    need = f(C.byref(batch.as_struct()), _p(hpos), _p(href), C.c_size_t(hpos.size), C.c_int32(min_bq), C.c_int32(min_mapq),
             C.c_int32(max(min_tstart, -2**31)), C.c_uint32(band), _p(out))
    return out, int(need)

