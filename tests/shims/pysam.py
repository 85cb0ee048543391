"""test shim: a pysam.AlignmentFile look-alike over an in-memory source.

pysam/htslib are not installed here.  The reference only touches the
surface listed in SURVEY.md §8b; this shim serves it from a registered provider:
    pysam.register(path, provider)
where provider has .header_text, .references and .fetch_records(chrom, start, end) yielding
objects with the AlignedSegment attributes below (tests/refshim.py builds providers from a
packed ReadBatch or from a BAM file read with himut_b200.bamio).
"""
_REGISTRY = {}


def register(path, provider):
    _REGISTRY[path] = provider


def unregister(path):
    _REGISTRY.pop(path, None)


class _Header:
    def __init__(self, text):
        self._text = text

    def __str__(self):
        return self._text


class AlignmentFile:
    def __init__(self, path, mode="rb", **kw):
        if path not in _REGISTRY:
            raise FileNotFoundError("pysam shim: %r is not registered" % (path,))
        self._p = _REGISTRY[path]
        self.header = _Header(self._p.header_text)

    def fetch(self, chrom=None, start=None, end=None):
        return self._p.fetch_records(chrom, start, end)

    def count(self, chrom=None, start=None, end=None):
        return sum(1 for _ in self._p.fetch_records(chrom, start, end))

    def close(self):
        pass
