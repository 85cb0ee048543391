"""GPU worker of the reference trinucleotide count: drop-in for himut.reflib.get_chrom_tricount
(src/himut/reflib.py:11-33), the per-chromosome task of get_genome_tricounts' Pool.starmap."""
from . import worker
from .normcounts import TRI_LST


def get_chrom_tricount(chrom, seq, chrom2tri2count):
    refseq = seq.encode() if isinstance(seq, str) else bytes(seq)
    tri = worker.context().ref_tricounts(refseq)
    d = {t: int(tri[i]) for i, t in enumerate(TRI_LST)}
    if tri[32]:
        d["NNN"] = int(tri[32])  # keys outside mutlib.tri_lst are never read back (reflib.py:55-60)
    chrom2tri2count[chrom] = d
