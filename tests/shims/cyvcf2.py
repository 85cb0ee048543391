"""test shim: cyvcf2 is only touched by header helpers outside the hot path"""


class VCF:
    def __init__(self, path, *a, **k):
        import gzip
        with open(path, "rb") as f:
            magic = f.read(2)
        opener = gzip.open if magic == b"\x1f\x8b" else open
        with opener(path, "rt") as f:
            lines = f.read().split("\n")
        self.raw_header = "\n".join(l for l in lines if l.startswith("#"))
        cols = [l for l in lines if l.startswith("#CHROM")]
        self.samples = cols[0].split("\t")[9:] if cols else []
        self._rows = [l for l in lines if l and not l.startswith("#")]

    def __iter__(self):
        return iter(())
