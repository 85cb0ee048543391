"""Shared machinery of the two worker mirrors: one lazily created context per process, the
region -> packed batch step, and chunk grouping so a contig never has to fit in one batch."""
import os

import numpy as np

from . import abi, bamdec, lib

_CTX = None
_CTX_PID = None

# upper bound of reference span decoded into one batch (30x -> ~1.5 GB of packed reads)
GROUP_SPAN = int(os.environ.get("HIMUT_B200_GROUP_SPAN", 40_000_000))


def context():
    """the process's hm_ctx, created on first use *inside* the worker process (fork safe).
    Device: HIMUT_B200_DEVICE, else LOCAL_RANK, else worker index modulo visible GPUs."""
    global _CTX, _CTX_PID
    if _CTX is None or _CTX_PID != os.getpid():
        dev = os.environ.get("HIMUT_B200_DEVICE", os.environ.get("LOCAL_RANK"))
        if dev is None:
            import multiprocessing as mp
            ident = mp.current_process()._identity
            dev = (ident[0] - 1) if ident else 0
            n = int(os.environ.get("HIMUT_B200_NUM_DEVICES", "0")) or _device_count()
            dev = dev % max(n, 1)
        _CTX = lib.Context(int(dev))
        _CTX_PID = os.getpid()
    return _CTX


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 1


def group_chunks(chunkloci_lst, span=None):
    """consecutive chunks grouped so each group's decoded window stays under `span` bases"""
    span = span or GROUP_SPAN
    groups, cur, lo, hi = [], [], None, None
    for i, (_c, s, e) in enumerate(chunkloci_lst):
        nlo = s if lo is None else min(lo, s)
        nhi = e if hi is None else max(hi, e)
        if cur and nhi - nlo > span:
            groups.append(cur)
            cur, nlo, nhi = [], s, e
        cur.append(i)
        lo, hi = nlo, nhi
    if cur:
        groups.append(cur)
    return groups


class RegionSource:
    """decodes the records a group of chunks fetches into one packed batch"""

    def __init__(self, bam_file):
        # native decoder (csrc/bamdec.c): threaded BGZF inflate, record parse and cs -> ops in C;
        # query-name ids are interned per handle, so they stay global across groups
        # (num_ccs counts distinct names per contig)
        self.reader = bamdec.NativeBam(bam_file)

    def batch(self, chrom, chunkloci, phase_sets=None, seq=True):
        """seq=False: without the 2-bit base stream (`call`, phase edges: ReadBatch.without_seq says why)"""
        lo = min(s for _, s, _e in chunkloci)
        hi = max(e for _, _s, e in chunkloci)
        batch = self.reader.read_batch(chrom, lo, hi, copy=False, seq=seq)
        table = batch.chunk_table([(s, e) for _, s, e in chunkloci], phase_sets)
        return batch, table

    def close(self):
        self.reader.close()


class QnameTally:
    """distinct query names that passed the read gates, across groups (m.num_ccs)"""

    def __init__(self):
        self.seen = np.zeros(0, np.uint8)

    def add(self, flags):
        if flags.size > self.seen.size:
            self.seen = np.concatenate([self.seen, np.zeros(flags.size - self.seen.size, np.uint8)])
        self.seen[: flags.size] |= flags

    def count(self):
        return int(np.count_nonzero(self.seen))
