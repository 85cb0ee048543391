"""Host-side logic that needs no GPU: BAM round trip, fetch semantics, set loaders, natsort,
chunk tables, chunk grouping."""
import os

import numpy as np

import cases
from himut_b200 import abi, bamio, natsort_compat, pack, synth, vcfio, worker


def test_bam_round_trip_and_fetch(tmp_path):
    d = synth.generate(120_000, seed=9)
    path = str(tmp_path / "t.bam")
    bamio.write_batch_bam(path, "chr1", 120_000, d.batch)
    assert os.path.exists(path + ".bai")
    rd = bamio.BamReader(path)
    assert rd.references == ["chr1"] and rd.lengths == [120_000]
    assert "SM:synth" in rd.header_text
    back = bamio.read_batch(rd, "chr1", 0, 120_000)
    assert cases.batch_digest(back) == cases.batch_digest(d.batch)
    b = d.batch
    for s, e in [(1, 60_000), (60_000, 119_998), (33_333, 33_334), (119_990, 120_000), (0, 1)]:
        got = [r.query_name for r in rd.fetch("chr1", s, e)]
        exp = ["read%d" % i for i in range(b.n_reads) if b.tstart[i] < e and b.tend[i] > s]
        assert got == exp, (s, e)
    rd.close()


def test_chunk_table_is_the_fetch_range():
    d = synth.generate(120_000, seed=10)
    b = d.batch
    t = b.chunk_table([(1, 50_000), (50_000, 119_998), (70_000, 70_001)])
    for row in t:
        idx = [i for i in range(b.n_reads) if b.tstart[i] < row["end"] and b.tend[i] > row["start"]]
        assert idx and row["read_lo"] <= idx[0] and idx[-1] < row["read_hi"]
        assert row["read_hi"] == np.searchsorted(b.tstart, row["end"], "left")


def test_pack_rejects_what_the_reference_crashes_on():
    bb = pack.BatchBuilder()
    good = dict(tstart=5, tend=9, qstart=0, qend=4, qseq="ACGT", bq=b"\x20" * 4, mapq=60, is_secondary=False, qname="q")
    bb.add(cs=":4", **good)
    for bad_cs, kw in [(":5", {}), (":2*an:1", {}), (":4", {"qseq": "ACNT"}), (":4", {"bq": b"\x20" * 3})]:
        try:
            pack.BatchBuilder().add(cs=bad_cs, **dict(good, **kw))
            assert False, (bad_cs, kw)
        except pack.BatchFormatError:
            pass
    # N in a soft clip is fine; N as the *reference* base of a substitution is fine
    pack.BatchBuilder().add(cs=":2*ng:1", tstart=5, tend=9, qstart=1, qend=5, qseq="NACGTN", bq=b"\x20" * 6, mapq=60,
                            is_secondary=False, qname="q")


def test_set_loaders_keep_the_reference_quirks(tmp_path):
    p = str(tmp_path / "c.vcf")
    with open(p, "w") as f:
        f.write("##x\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS\n")
        f.write("chr1\t100\t.\tA\tG\t.\tPASS\t.\tGT\t0/1\n")
        f.write("chr2\t200\t.\tC\tT\t.\tPASS\t.\tGT\t0/1\n")
        f.write("chr2\t300\t.\tC\tT,G\t.\tPASS\t.\tGT\t0/1\n")
        f.write("chr2\t400\t.\tC\tT\t.\tLowQ\t.\tGT\t0/1\n")
        f.write("chr2\t500\t.\tCA\tC\t.\tPASS\t.\tGT\t0/1\n")
    # plain .vcf common SNPs: records whose CHROM differs are the ones kept (vcflib.py:434)
    assert list(vcfio.load_common_snps("chr1", p)) == [(200 << 4) | (3 << 2) | 1]
    assert list(vcfio.load_pon("chr1", p)) == [(100 << 4) | (0 << 2) | 2]
    q = str(tmp_path / "c.vcf.bgz")
    os.rename(p, q)
    assert list(vcfio.load_common_snps("chr1", q)) == [(100 << 4) | (0 << 2) | 2]


def test_natsort_compat():
    ns = natsort_compat.natsorted
    assert ns(["chr10", "chr2", "chr1", "chrX"]) == ["chr1", "chr2", "chr10", "chrX"]
    rows = [("chr1", 20, "A", "T", "PASS", 5), ("chr1", 3, "C", "G", "LowBQ", 7), ("chr1", 20, "A", "G", "PASS", 1)]
    assert ns(rows) == [rows[1], rows[2], rows[0]]
    mixed = [("c", 5, "A", "C,G", "HetAltSite", 9, "30.1,25.0"), ("c", 5, "A", "C", "PASS", 9, 30.0)]
    assert ns(mixed) == [mixed[1], mixed[0]]


def test_group_chunks():
    loci = [("c", 1, 200000), ("c", 200000, 400000), ("c", 400000, 600000), ("c", 600000, 799998)]
    assert worker.group_chunks(loci, span=450_000) == [[0, 1], [2, 3]]
    assert worker.group_chunks(loci, span=10) == [[0], [1], [2], [3]]
    assert worker.group_chunks(loci, span=10**9) == [[0, 1, 2, 3]]


def _rows_one_by_one(chrom, recs):
    from himut_b200 import records
    rows = [records.record_to_tuple(chrom, r) for r in recs if int(r["status"]) in records._EMITTED]
    return natsort_compat.natsorted(list(set(rows)))


def test_rows_from_whole_arrays_equal_rows_one_by_one():
    """records_to_tsbs_lst does the column arithmetic on arrays and sorts by (pos, ref, alt): same rows, same order
    as record_to_tuple per record + natsorted (caller.py:622-624)"""
    from himut_b200 import records
    from oracle import oracle
    for name in ("call_sets", "call_phase", "call_adversarial_a", "call_adversarial_b", "call_pon_params"):
        c = cases.build_case(name)
        rec, _ = oracle.call_chunks(c["params"], c["batch"], c["chunk_table"], c["common"], c["pon"], c["phase"])
        got = records.records_to_tsbs_lst(cases.CHROM, rec)
        assert got == _rows_one_by_one(cases.CHROM, rec), name
    assert records.records_to_tsbs_lst(cases.CHROM, np.zeros(0, abi.SITE_DTYPE)) == []


def test_row_order_falls_back_to_natsort_on_equal_sites():
    from himut_b200 import records
    a = ("chr1", 10, "A", "C", "LowBQ", 5, 20.0, 30.0, 28.0, 2.0, 0.07, ".")
    b = ("chr1", 10, "A", "C", "LowGQ", 5, 20.0, 30.0, 28.0, 2.0, 0.07, ".")
    c = ("chr1", 9, "T", "G", "PASS", 50, 93.0, 30.0, 29.0, 1.0, 0.03, ".")
    d = ("chr1", 10, "A", "C,G", "HetAltSite", 50, "93.0,93.0", 30.0, 0.0, "15,15", "0.50,0.50", ".")
    for rows in ([b, a, c], [d, b, c], [c, d, a, b]):
        assert records._sorted_rows(list(rows)) == natsort_compat.natsorted(list(rows))


def test_site_files_are_read_through_the_tabix_index(tmp_path, monkeypatch):
    """`.bgz` + `.tbi`: only the contig's BGZF range is inflated, and the keys equal those of a scan from the top"""
    import random
    import tbi
    from himut_b200 import vcfio
    rnd = random.Random(11)
    header = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
    records = []
    for chrom, n in (("chr1", 3_000_000), ("chr2", 40_000), ("chrX", 900_000)):
        for pos in sorted(rnd.sample(range(1, n), 1500)):
            ref = rnd.choice(["A", "T", "G", "C", "AT"])
            alt = rnd.choice(["A", "T", "G", "C", "A,T", "GC"])
            flt = rnd.choice(["PASS", "PASS", "PASS", "q10"])
            records.append((chrom, pos, ref, "%s\t%d\t.\t%s\t%s\t.\t%s\t." % (chrom, pos, ref, alt, flt)))
    path = str(tmp_path / "sites.vcf.bgz")
    tbi.write_vcf_bgz_tbi(path, header, records)
    assert vcfio.tabix_contig_range(path + ".tbi", "chr9") is None
    lo1, hi1 = vcfio.tabix_contig_range(path + ".tbi", "chr1")
    lo2, hi2 = vcfio.tabix_contig_range(path + ".tbi", "chr2")
    assert lo1 < hi1 <= lo2 < hi2
    indexed = {c: (vcfio.load_common_snps(c, path), vcfio.load_pon(c, path)) for c in ("chr1", "chr2", "chrX", "chr9")}
    # the index really was the way in: a scan from the top is not allowed here
    monkeypatch.setattr(vcfio, "_open_text", lambda p: (_ for _ in ()).throw(AssertionError("scanned")))
    again = vcfio.load_common_snps("chr2", path)
    monkeypatch.undo()
    assert np.array_equal(again, indexed["chr2"][0])
    os.rename(path + ".tbi", path + ".tbi.away")      # no index: scan
    for c, (common, pon) in indexed.items():
        assert np.array_equal(common, vcfio.load_common_snps(c, path)), c
        assert np.array_equal(pon, vcfio.load_pon(c, path)), c
    assert indexed["chr1"][0].size > 300 and indexed["chr9"][0].size == 0
    open(path + ".tbi", "wb").write(b"not an index")    # unreadable index: scan, same keys
    assert np.array_equal(indexed["chrX"][0], vcfio.load_common_snps("chrX", path))
