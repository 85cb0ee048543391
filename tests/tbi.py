"""Test helper: writes a BGZF-compressed, position-sorted VCF and the tabix index (`.tbi`) `tabix -p vcf` would
write for it — bins by reg2bin, merged chunks per bin, the 16 kb linear index and the pseudo-bin 37450 — so that the
indexed readers of himut_b200.vcfio can be tested without htslib.  Layout per the tabix specification
(samtools.github.io/hts-specs/tabix.pdf)."""
import gzip
import struct
import zlib


def reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def _bgzf_block(chunk):
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = c.compress(chunk) + c.flush()
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(comp) + 25) + comp
            + struct.pack("<II", zlib.crc32(chunk), len(chunk)))


def write_vcf_bgz_tbi(path, header, records, block_bytes=3000):
    """records: [(chrom, pos (1-based), ref, line without newline)] sorted by contig then position;
    contigs appear in the index in order of first appearance"""
    text = header.encode()
    starts = []
    for chrom, pos, ref, line in records:
        starts.append(len(text))
        text += line.encode() + b"\n"
    ends = starts[1:] + [len(text)]
    # blocks of fixed uncompressed size, like bgzip: lines may straddle blocks
    block_start, coffs, out = [], [], bytearray()
    for off in range(0, len(text), block_bytes):
        block_start.append(off)
        coffs.append(len(out))
        out += _bgzf_block(text[off:off + block_bytes])
    eof_coff = len(out)
    out += _bgzf_block(b"")  # the EOF marker block
    with open(path, "wb") as f:
        f.write(out)

    def voff(p):
        if p >= len(text):
            return eof_coff << 16
        b = p // block_bytes
        return (coffs[b] << 16) | (p - block_start[b])

    names, per = [], {}
    for (chrom, pos, ref, _line), s, e in zip(records, starts, ends):
        if chrom not in per:
            names.append(chrom)
            per[chrom] = dict(bins={}, lin=[], first=voff(s), last=voff(e), n=0)
        d = per[chrom]
        beg, end = pos - 1, pos - 1 + len(ref)
        b = reg2bin(beg, end)
        chunks = d["bins"].setdefault(b, [])
        if chunks and chunks[-1][1] == voff(s):
            chunks[-1][1] = voff(e)      # adjacent records of one bin share a chunk
        else:
            chunks.append([voff(s), voff(e)])
        for w in range(beg >> 14, ((end - 1) >> 14) + 1):
            while len(d["lin"]) <= w:
                d["lin"].append(0)
            if d["lin"][w] == 0:
                d["lin"][w] = voff(s)
        d["last"], d["n"] = voff(e), d["n"] + 1
    nm = b"".join(n.encode() + b"\0" for n in names)
    idx = bytearray(b"TBI\x01") + struct.pack("<8i", len(names), 2, 1, 2, 0, ord("#"), 0, len(nm)) + nm
    for n in names:
        d = per[n]
        idx += struct.pack("<i", len(d["bins"]) + 1)
        for b, chunks in d["bins"].items():
            idx += struct.pack("<Ii", b, len(chunks))
            for c0, c1 in chunks:
                idx += struct.pack("<QQ", c0, c1)
        idx += struct.pack("<Ii", 37450, 2) + struct.pack("<QQQQ", d["first"], d["last"], d["n"], 0)
        lin = d["lin"]
        for i in range(1, len(lin)):      # empty windows inherit the previous offset, as htslib fills them
            if lin[i] == 0:
                lin[i] = lin[i - 1]
        idx += struct.pack("<i", len(lin)) + b"".join(struct.pack("<Q", v) for v in lin)
    idx += struct.pack("<Q", 0)
    with open(path + ".tbi", "wb") as f:
        f.write(_bgzf_block(bytes(idx)) + _bgzf_block(b""))
