"""test shim: pyfastx.Fasta over a plain FASTA file"""


class _Seq:
    def __init__(self, name, seq):
        self.name, self.seq = name, seq

    def __str__(self):
        return self.seq

    def __len__(self):
        return len(self.seq)

    def __getitem__(self, i):
        return self.seq[i]


class Fasta:
    def __init__(self, path, *a, **k):
        self._d = {}
        name, buf = None, []
        with open(path) as f:
            for line in f:
                if line.startswith(">"):
                    if name is not None:
                        self._d[name] = _Seq(name, "".join(buf))
                    name, buf = line[1:].split()[0], []
                else:
                    buf.append(line.strip())
        if name is not None:
            self._d[name] = _Seq(name, "".join(buf))

    def __getitem__(self, k):
        return self._d[k]

    def __contains__(self, k):
        return k in self._d

    def keys(self):
        return self._d.keys()

    def __iter__(self):
        return iter(self._d.values())
