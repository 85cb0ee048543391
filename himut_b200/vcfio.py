"""Common-SNP / panel-of-normals / phased-hetSNP inputs -> the sorted key arrays and tables the
library takes (hm_set_site_sets, hm_set_phase_sets).

Mirrors the reference's loaders, quirks included (src/himut/vcflib.py:356-459):
  * plain `.vcf` common SNPs keep records whose CHROM *differs* from the worker's contig
    (`chrom != arr[0]`, vcflib.py:434) — reproduced so results stay identical;
  * `.bgz` inputs are queried per chunk ± qlen_upper_limit by the reference, which for sites
    inside the chunk equals contig-wide membership, so one contig-wide set is built;
  * only FILTER == PASS records; common SNPs need a single one-letter ALT, the PoN needs a
    bi-allelic SNP.
"""
import gzip
import os
import struct
import zlib

import numpy as np

from . import abi


def _open_text(path):
    with open(path, "rb") as f:
        magic = f.read(2)
    return gzip.open(path, "rt") if magic == b"\x1f\x8b" else open(path, "rt")


# ---- tabix ---------------------------------------------------------------------------------------
# The reference reads `.bgz` site files through pytabix region queries (vcflib.py:381,417,449,640), so a genome-wide
# common-SNP file costs it only the records of the region.  The workers here want one contig's records at a time:
# the `.tbi` index names the BGZF range that holds them, and only that range is inflated.  Without a usable index
# the file is scanned from the top (same result, slower on genome-wide files).
_TBI_PSEUDO_BIN = 37450


def tabix_contig_range(tbi_path, chrom):
    """-> (begin, end) BGZF virtual offsets of `chrom`'s records per the .tbi index, None when the index does not
    list the contig, raises ValueError on anything that is not a tabix index"""
    with open(tbi_path, "rb") as f:
        raw = gzip.decompress(f.read())
    if raw[:4] != b"TBI\x01":
        raise ValueError("not a tabix index")
    n_ref, _fmt, _cs, _cb, _ce, _meta, _skip, l_nm = struct.unpack_from("<8i", raw, 4)
    p = 36
    names = raw[p:p + l_nm].split(b"\0")[:n_ref]
    p += l_nm
    want = chrom.encode()
    for r in range(n_ref):
        n_bin = struct.unpack_from("<i", raw, p)[0]
        p += 4
        lo, hi, pseudo = None, None, None
        for _ in range(n_bin):
            bin_id, n_chunk = struct.unpack_from("<Ii", raw, p)
            p += 8
            chunks = struct.unpack_from("<%dQ" % (2 * n_chunk), raw, p)
            p += 16 * n_chunk
            if names[r] != want:
                continue
            if bin_id == _TBI_PSEUDO_BIN:   # (first, last) offsets of the contig + record counts
                if n_chunk >= 1:
                    pseudo = (chunks[0], chunks[1])
                continue
            for i in range(n_chunk):
                lo = chunks[2 * i] if lo is None else min(lo, chunks[2 * i])
                hi = chunks[2 * i + 1] if hi is None else max(hi, chunks[2 * i + 1])
        n_intv = struct.unpack_from("<i", raw, p)[0]
        p += 4 + 8 * n_intv
        if names[r] == want:
            if lo is not None:
                return lo, hi
            return pseudo
    return None


def _bgzf_lines(path, begin, end):
    """text lines of the BGZF file between two virtual offsets"""
    from .bamio import _read_block
    out = []
    with open(path, "rb") as f:
        coff, skip = begin >> 16, begin & 0xFFFF
        end_coff, end_uoff = end >> 16, end & 0xFFFF
        f.seek(coff)
        while coff <= end_coff:
            size, data = _read_block(f)
            if size == 0:
                break
            if coff == end_coff:
                data = data[:end_uoff]
            out.append(data[skip:] if skip else data)
            skip = 0
            coff += size
    return b"".join(out).decode().split("\n")


def _contig_lines(path, chrom):
    """the lines of a site file that can belong to `chrom`: through the tabix index when there is one"""
    tbi = path + ".tbi"
    if path.endswith(".bgz") and os.path.exists(tbi):
        try:
            rng = tabix_contig_range(tbi, chrom)
        except (ValueError, OSError, struct.error, EOFError, zlib.error):
            rng = False  # not a readable index (the tests' placeholder, a csi index, a truncated file): scan instead
        if rng is None:
            return []
        if rng:
            return _bgzf_lines(path, rng[0], rng[1])
    with _open_text(path) as f:
        return f.read().split("\n")


def _key(pos, ref, alt):
    r, a = abi.BASE2CODE.get(ref), abi.BASE2CODE.get(alt)
    if r is None or a is None:
        return None  # can never equal a candidate (ref, alt are A/T/G/C there)
    return (int(pos) << 4) | (r << 2) | a


def _finish(keys):
    return np.unique(np.asarray(sorted(keys), dtype=np.uint64)) if keys else np.zeros(0, np.uint64)


def load_common_snps(chrom, path):
    """vcflib.load_common_snp (.vcf) / load_bgz_common_snp (.bgz) as sorted keys"""
    if path is None:
        return np.zeros(0, np.uint64)
    plain = path.endswith(".vcf")
    keys = []
    if plain:
        with _open_text(path) as f:
            lines = f.read().split("\n")
    else:
        lines = _contig_lines(path, chrom)
    for line in lines:
        if line.startswith("#"):
            continue
        arr = line.strip().split()
        if len(arr) < 7:
            continue
        same = arr[0] == chrom
        if plain and same:      # vcflib.py:434 keeps `chrom != arr[0]`
            continue
        if not plain and not same:
            continue
        alts = arr[4].split(",")
        if arr[6] == "PASS" and len(alts) == 1 and len(arr[3]) == 1 and len(alts[0]) == 1:
            k = _key(arr[1], arr[3], alts[0])
            if k is not None:
                keys.append(k)
    return _finish(keys)


def load_pon(chrom, path):
    """vcflib.load_pon (.vcf) / load_bgz_pon (.bgz) as sorted keys"""
    if path is None:
        return np.zeros(0, np.uint64)
    keys = []
    for line in _contig_lines(path, chrom):
        if line.startswith("#"):
            continue
        arr = line.strip().split()
        if len(arr) < 7 or arr[0] != chrom:
            continue
        alts = arr[4].split(",")
        if arr[6] == "PASS" and len(alts) == 1 and len(arr[3]) == 1 and len(alts[0]) == 1:
            k = _key(arr[1], arr[3], alts[0])
            if k is not None:
                keys.append(k)
    return _finish(keys)


def phase_tables(chunkloci_lst, phase_set2hbit_lst, phase_set2hpos_lst, phase_set2hetsnp_lst):
    """the worker's three per-phase-set dicts (vcflib.load_phased_hetsnps, vcflib.py:617-663) ->
    (phase table for hm_set_phase_sets, phase-set index per chunk).  The reference looks the
    lists up with str(chunk_start) (caller.py:292)."""
    hpos, href, halt, hbit, set_off, chunk_sets = [], [], [], [], [0], []
    index = {}
    for (_c, start, _e) in chunkloci_lst:
        key = str(start)
        if key not in index:
            pos = phase_set2hpos_lst[key]
            snp = phase_set2hetsnp_lst[key]
            bits = phase_set2hbit_lst[key]
            index[key] = len(set_off) - 1
            for p, s, b in zip(pos, snp, bits):
                hpos.append(int(p))
                # the reference compares the read's one-letter base with the REF / ALT *strings*
                # (haplib.get_ccs_hbit, haplib.py:46-58): an indel allele (or a lower-case one) never equals it: 255.
                # A batch without its base stream (`call`) takes the base of a cs match run from the table: the
                # contig's base at hpos is REF's first letter, so a REF that can never be equal is stored as
                # 16 + code(REF[0]) — "never equal, but this is what a matching read shows" (the ALT test then sees
                # the real base, e.g. REF=AT ALT=A gives bit 1 for every read that matches at the anchor)
                ref_code = abi.BASE2CODE.get(s[1], 255) if len(s[1]) == 1 else 255
                if ref_code == 255 and s[1][:1].upper() in abi.BASE2CODE:
                    ref_code = 16 + abi.BASE2CODE[s[1][:1].upper()]
                href.append(ref_code)
                halt.append(abi.BASE2CODE.get(s[2], 255) if len(s[2]) == 1 else 255)
                hbit.append(int(b) if b in ("0", "1") else 2)
            set_off.append(len(hpos))
        chunk_sets.append(index[key])
    table = dict(hpos=np.asarray(hpos, np.int32), href=np.asarray(href, np.uint8), halt=np.asarray(halt, np.uint8),
                 hbit=np.asarray(hbit, np.uint8), set_off=np.asarray(set_off, np.uint64))
    return table, chunk_sets


# ---- writers -----------------------------------------------------------------------------------
# Text the reference's vcflib.dump_sbs / dump_phased_sbs (src/himut/vcflib.py:820-1021) produce from
# chrom2tsbs_lst, reproduced so that "BAM on disk -> VCF on disk" can be run and timed without the reference
# (bench.py); in a drop-in installation the reference's own writers stay in place.  Quirks kept: a HetAltSite row
# carries pre-formatted strings and goes to the single-molecule file when int(ref_count) == 1 (every other row:
# int(alt_count) == 1), and the phased main file names no PS key for a HetAltSite row although the sample has one.
_KEYS = "GT:GQ:BQ:DP:AD:VAF"
# one template per (row kind, with PS or not); %-formatting and str.format share the float conversion
# (PyOS_double_to_string), "%s" of an int is "{}" of it
_HEAD = "%s\t%s\t.\t%s\t%s\t.\t%s\t.\t"
_OTHER = "\t./.:%s:%0.1f:%0.0f:%0.0f,%0.0f:%.2f"
_HETALT = "\t./.:%s:%s:%0.0f:%0.0f,%s:%s"


def dump_sbs(vcf_file, vcf_header, chrom_lst, chrom2tsbs_lst, phased=False):
    """writes vcf_file and its .single_molecule_mutations.vcf twin; phased=True is dump_phased_sbs"""
    if not vcf_file.endswith(".vcf"):
        raise ValueError("VCF file must have .vcf suffix")
    ps_key, ps_val = (":PS", ":%s\n") if phased else ("", "\n")
    t_other = _HEAD + _KEYS + ps_key + _OTHER + ps_val
    t_hetalt_main = _HEAD + _KEYS + _HETALT + ps_val
    t_hetalt_single = _HEAD + _KEYS + ps_key + _HETALT + ps_val
    main, single = ["{}\n".format(vcf_header)], ["{}\n".format(vcf_header)]
    for chrom in chrom_lst:
        for row in chrom2tsbs_lst[chrom]:
            vals = (row[0], row[1], row[2], row[3], row[4], row[5], row[6], row[7], row[8], row[9], row[10])
            if phased:
                vals += (row[11],)
            if row[4] == "HetAltSite":
                main.append(t_hetalt_main % vals)
                if int(row[8]) == 1:
                    single.append(t_hetalt_single % vals)
            else:
                line = t_other % vals
                main.append(line)
                if int(row[9]) == 1:
                    single.append(line)
    with open(vcf_file, "w") as f:
        f.write("".join(main))
    with open(vcf_file.replace(".vcf", ".single_molecule_mutations.vcf"), "w") as f:
        f.write("".join(single))


def dump_phased_sbs(vcf_file, vcf_header, chrom_lst, chrom2tsbs_lst):
    dump_sbs(vcf_file, vcf_header, chrom_lst, chrom2tsbs_lst, phased=True)
