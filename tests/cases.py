"""Deterministic parity cases shared by tests/golden/make_golden.py (reference side, run in the
build container) and the parity tests (oracle / CUDA side, run anywhere)."""
import hashlib
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from himut_b200 import abi, gtmodel, pack, synth  # noqa: E402

CHROM = "chr1"
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def chunkloci(start, end):
    """reference util.chunkloci (src/himut/util.py:119-132) restated for (start, end)"""
    if end - start > 200000:
        out = [(1, 200000)]
        starts = list(range(200000, end, 200000))
        for i, s in enumerate(starts[:-1]):
            out.append((s, starts[i + 1]))
        if (starts[-1], end) not in out:
            out.append((starts[-1], end - 2))
        return out
    return [(start, end)]


def batch_digest(batch):
    h = hashlib.sha256()
    for name, _ in abi.ReadBatch._FIELDS:
        h.update(np.ascontiguousarray(getattr(batch, name)).tobytes())
    return h.hexdigest()


def adversarial_batch(seed, contig_len=3000, n_reads=120, max_len=900):
    """short, dirty reads: dense subs/indels, N reference bases, soft clips, BQ spread,
    secondary records, duplicate query names, reads ending at every kind of op boundary"""
    rnd = random.Random(seed)
    ref = [rnd.choice("ACGT") for _ in range(contig_len)]
    for _ in range(contig_len // 200):
        ref[rnd.randrange(contig_len)] = "N"
    # a handful of "germline" positions so genotypes other than homref appear
    germ = {}
    for _ in range(contig_len // 60):
        p = rnd.randrange(contig_len)
        if ref[p] != "N":
            alts = [b for b in "ACGT" if b != ref[p]]
            rnd.shuffle(alts)
            # kind 0/1: het on that haplotype, 2: hom-alt, 3: het-alt (a different alt per haplotype)
            germ[p] = (alts[0], rnd.choice([0, 1, 2, 2, 3]), alts[1])
    starts = sorted(rnd.randrange(0, contig_len - 60) for _ in range(n_reads))
    bb = pack.BatchBuilder()
    for i, ts in enumerate(starts):
        length = rnd.randrange(40, max_len)
        te = min(ts + length, contig_len)
        hap = rnd.randrange(2)
        lead = rnd.choice([0, 0, 0, rnd.randrange(1, 30)])
        trail = rnd.choice([0, 0, 0, rnd.randrange(1, 30)])
        bq_mode = rnd.choice(["hi", "hi", "mix", "low"])

        def bq():
            if bq_mode == "hi":
                return 93 if rnd.random() < 0.9 else rnd.randrange(1, 94)
            if bq_mode == "mix":
                return rnd.choice([93, 93, 60, 40, 20, 10, 3, 1])
            return rnd.randrange(1, 45)

        qseq, bqs, cs = [], [], []
        for _ in range(lead):
            qseq.append(rnd.choice("ACGTN")); bqs.append(bq())
        t, run, last_indel = ts, 0, True  # no indel before the first matched base
        div = rnd.choice([0.0, 0.002, 0.01, 0.05])
        while t < te:
            r = ref[t]
            g = germ.get(t)
            u = rnd.random()
            if r == "N":
                if run:
                    cs.append(":%d" % run); run = 0
                a = rnd.choice("ACGT")
                cs.append("*n%s" % a.lower()); qseq.append(a); bqs.append(bq()); t += 1; last_indel = False
            elif g and (g[1] >= 2 or g[1] == hap) and rnd.random() < 0.95:
                if run:
                    cs.append(":%d" % run); run = 0
                ga = g[2] if (g[1] == 3 and hap == 1) else g[0]
                cs.append("*%s%s" % (r.lower(), ga.lower())); qseq.append(ga); bqs.append(bq()); t += 1; last_indel = False
            elif u < div:
                if run:
                    cs.append(":%d" % run); run = 0
                a = rnd.choice([b for b in "ACGT" if b != r])
                cs.append("*%s%s" % (r.lower(), a.lower())); qseq.append(a); bqs.append(bq()); t += 1; last_indel = False
            elif u < 2 * div and not last_indel and t + 6 < te:
                if run:
                    cs.append(":%d" % run); run = 0
                n = rnd.randrange(1, 5)
                if rnd.random() < 0.5:
                    ins = [rnd.choice("ACGT") for _ in range(n)]
                    cs.append("+" + "".join(ins).lower()); qseq.extend(ins); bqs.extend(bq() for _ in range(n))
                else:
                    cs.append("-" + "".join(x.lower() for x in ref[t:t + n])); t += n
                last_indel = True
            else:
                qseq.append(r); bqs.append(bq()); run += 1; t += 1; last_indel = False
        if run:
            cs.append(":%d" % run)
        qend = len(qseq)
        for _ in range(trail):
            qseq.append(rnd.choice("ACGTN")); bqs.append(bq())
        qname = "r%d" % (i if rnd.random() > 0.03 else max(i - 1, 0))
        bb.add(tstart=ts, tend=None, qstart=lead, qend=qend, qseq="".join(qseq), bq=bytes(bqs),
               mapq=rnd.choice([60, 60, 60, 60, 30, 0]), is_secondary=rnd.random() < 0.03,
               qname=qname, cs="".join(cs))
    return bb.finish(), "".join(ref)


# ------------------------------------------------------------------------------------------
# case table: name -> dict(kind, make() -> (batch, refseq str), chunks, args, sets, phase)
# ------------------------------------------------------------------------------------------
def _synth_case(contig_len, seed, **over):
    d = synth.generate(contig_len, seed=seed, **over)
    return d


def call_args(**over):
    a = dict(gtmodel.DEFAULT_CALL_ARGS)
    a.update(over)
    return a


def site_sets_from_synth(d, seed):
    """synthetic common-SNP and PoN sets: most germline sites + a few somatic
    sites are 'common'; a few error + somatic sites are in the panel of normals"""
    rng = np.random.default_rng(seed)
    g, s, e = d.germ, d.som, d.err
    pick = lambda n, f: rng.random(n) < f
    gm, sm, em, sm2 = pick(g["pos"].size, 0.8), pick(s["pos"].size, 0.3), pick(e["pos"].size, 0.3), pick(s["pos"].size, 0.3)
    common = np.concatenate([synth.site_keys(g["pos"][gm], g["ref"][gm], g["alt"][gm]),
                             synth.site_keys(s["pos"][sm], s["ref"][sm], s["alt"][sm])])
    pon = np.concatenate([synth.site_keys(e["pos"][em], e["ref"][em], e["alt"][em]),
                          synth.site_keys(s["pos"][sm2], s["ref"][sm2], s["alt"][sm2])])
    return np.unique(common), np.unique(pon)


def phase_case(d, phase_block):
    """phase table + chunks = one (first_hpos, last_hpos) span per phase set (vcflib.py:655-662)"""
    ph = synth.phase_table(d.germ, phase_block)
    chunks, sets = [], []
    for s in range(ph["set_off"].size - 1):
        a, b = int(ph["set_off"][s]), int(ph["set_off"][s + 1])
        chunks.append((int(ph["hpos"][a]), int(ph["hpos"][b - 1])))
        sets.append(s)
    return ph, chunks, sets


def write_sites_vcf(path, chrom, keys):
    """minimal single-sample VCF of PASS SNVs from site keys"""
    with open(path, "w") as f:
        f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS\n")
        for k in keys:
            k = int(k)
            f.write("%s\t%d\t.\t%s\t%s\t.\tPASS\t.\tGT\t0/1\n" % (chrom, k >> 4, "ATGC"[(k >> 2) & 3], "ATGC"[k & 3]))


def phase_dicts(c):
    """the worker's phase_set2hbit_lst / hpos_lst / hetsnp_lst dicts (vcflib.load_phased_hetsnps).
    c["phase_snps"] (optional): the SNP table the dicts start from when c["phase"] holds the converted table of a case
    with indel records; c["phase_alleles"] (optional): {index in the table: (REF string, ALT string, hbit string)} —
    phased records that are not SNPs (load_phased_hetsnps does not filter them out)"""
    ph = c.get("phase_snps") if c.get("phase_snps") is not None else c["phase"]
    over = c.get("phase_alleles") or {}
    hbit, hpos, hetsnp = {}, {}, {}
    if ph is None:
        return hbit, hpos, hetsnp
    for s in range(ph["set_off"].size - 1):
        a, b = int(ph["set_off"][s]), int(ph["set_off"][s + 1])
        key = str(int(ph["hpos"][a]))
        hpos[key] = [int(x) for x in ph["hpos"][a:b]]
        hbit[key] = [str(int(x)) for x in ph["hbit"][a:b]]
        hetsnp[key] = [(int(p), "ATGC"[r], "ATGC"[al]) for p, r, al in zip(ph["hpos"][a:b], ph["href"][a:b], ph["halt"][a:b])]
        for i in range(a, b):
            if i in over:
                ref_s, alt_s, bit_s = over[i]
                hetsnp[key][i - a] = (hetsnp[key][i - a][0], ref_s, alt_s)
                hbit[key][i - a] = bit_s
    return hbit, hpos, hetsnp


CASES = {}


def case(name):
    def deco(fn):
        CASES[name] = fn
        return fn
    return deco


@case("call_basic")
def _call_basic():
    d = _synth_case(260_000, 11)
    return dict(kind="call", batch=d.batch, ref=d.ref.decode(), contig_len=260_000,
                chunks=chunkloci(0, 260_000), args=call_args())


@case("call_config0_1mb")
def _call_config0():
    # BASELINE.json configs[0]: 1 Mb contig, 30x, the reference's own 200 kb chunking
    d = _synth_case(1_000_000, 20260101)
    return dict(kind="call", batch=d.batch, ref=d.ref.decode(), contig_len=1_000_000,
                chunks=chunkloci(0, 1_000_000), args=call_args())


@case("norm_config0_1mb")
def _norm_config0():
    d = _synth_case(1_000_000, 20260101)
    return dict(kind="norm", batch=d.batch, ref=d.ref.decode(), contig_len=1_000_000,
                chunks=chunkloci(0, 1_000_000), args=call_args())


@case("call_sets")
def _call_sets():
    d = _synth_case(230_000, 12, somatic_rate=2e-5)
    common, pon = site_sets_from_synth(d, 12)
    return dict(kind="call", batch=d.batch, ref=d.ref.decode(), contig_len=230_000,
                chunks=chunkloci(0, 230_000), args=call_args(), common=common, pon=pon)


@case("call_lowdepth")
def _call_lowdepth():
    d = _synth_case(150_000, 13, depth=9.0, somatic_rate=2e-5, sub_err_rate=4e-4)
    return dict(kind="call", batch=d.batch, ref=d.ref.decode(), contig_len=150_000,
                chunks=chunkloci(0, 150_000), args=call_args(md_threshold=14, min_gq=10, min_ref_count=7))


@case("call_pon_params")
def _call_pon_params():
    # --create_panel_of_normals preset (util.load_pon_params, src/himut/util.py:44-63)
    d = _synth_case(120_000, 14, sub_err_rate=5e-4)
    common, pon = site_sets_from_synth(d, 14)
    return dict(kind="call", batch=d.batch, ref=d.ref.decode(), contig_len=120_000,
                chunks=chunkloci(0, 120_000),
                args=call_args(min_bq=20, min_gq=10, min_qv=20, min_mapq=30, min_trim=0, min_hap_count=0,
                               min_sequence_identity=0.8, create_panel_of_normals=True),
                common=common, pon=pon)


@case("call_phase")
def _call_phase():
    d = _synth_case(210_000, 15, somatic_rate=2e-5, phase_block=50_000)
    ph, chunks, sets = phase_case(d, 50_000)
    return dict(kind="call", batch=d.batch, ref=d.ref.decode(), contig_len=210_000,
                chunks=chunks, phase_sets=sets, phase=ph, args=call_args(phase=True))


@case("call_phase_indel")
def _call_phase_indel():
    # phased records that are not SNPs (vcflib.load_phased_hetsnps keeps them): haplib.get_ccs_hbit compares the read's
    # one-letter base with the REF / ALT *strings*, so at a deletion record (REF = two bases, ALT = the first) a read
    # that shows the reference base gets bit 1, at an insertion record (ALT = two bases) bit 0, and a read with the other
    # allele "-" (unphased).  The record's bit is set so that the haplotype with the reference base stays consistent.
    from himut_b200 import vcfio
    d = _synth_case(210_000, 18, somatic_rate=2e-5, phase_block=50_000)
    ph, chunks, sets = phase_case(d, 50_000)
    ref = d.ref.decode()
    over = {}
    for i in range(4, ph["hpos"].size, 9):
        p, r, bit = int(ph["hpos"][i]), "ATGC"[int(ph["href"][i])], int(ph["hbit"][i])
        if (i // 9) % 2 == 0:
            over[i] = (r + ref[p].upper(), r, str(1 - bit))       # deletion: hpos is 1-based, ref[p] is the next base
        else:
            over[i] = (r, r + "G", str(bit))                      # insertion
    c = dict(kind="call", batch=d.batch, ref=ref, contig_len=210_000, chunks=chunks, phase_sets=sets, phase=None,
             phase_snps=ph, phase_alleles=over, args=call_args(phase=True))
    hbit, hpos, hetsnp = phase_dicts(c)
    table, chunk_sets = vcfio.phase_tables([(CHROM, s_, e_) for s_, e_ in chunks], hbit, hpos, hetsnp)
    assert chunk_sets == sets
    c["phase"] = table  # what the worker mirror hands to the device for these dicts
    return c


@case("call_phase_shallow")
def _call_phase_shallow():
    # thin coverage: haplotype counts fall under min_hap_count -> Unphased rows
    d = _synth_case(150_000, 16, depth=14.0, somatic_rate=6e-5, phase_block=30_000)
    ph, chunks, sets = phase_case(d, 30_000)
    return dict(kind="call", batch=d.batch, ref=d.ref.decode(), contig_len=150_000,
                chunks=chunks, phase_sets=sets, phase=ph, args=call_args(phase=True, md_threshold=30))


@case("call_config3_phase_sets")
def _call_config3():
    # BASELINE.json configs[3] in small: --phase with a phased germline table AND common-SNP AND panel-of-normals sets
    d = _synth_case(240_000, 17, somatic_rate=4e-5, sub_err_rate=3e-4, phase_block=60_000)
    ph, chunks, sets = phase_case(d, 60_000)
    common, pon = site_sets_from_synth(d, 17)
    return dict(kind="call", batch=d.batch, ref=d.ref.decode(), contig_len=240_000,
                chunks=chunks, phase_sets=sets, phase=ph, args=call_args(phase=True), common=common, pon=pon)


@case("norm_config3_phase_sets")
def _norm_config3():
    d = _synth_case(60_000, 24, sub_err_rate=1e-3, somatic_rate=1e-4, phase_block=20_000)
    ph, chunks, sets = phase_case(d, 20_000)
    common, pon = site_sets_from_synth(d, 24)
    return dict(kind="norm", batch=d.batch, ref=d.ref.decode(), contig_len=60_000,
                chunks=chunks, phase_sets=sets, phase=ph, args=call_args(phase=True), common=common, pon=pon)


def _adv(seed, **over):
    batch, ref = adversarial_batch(seed)
    args = call_args(min_qv=20, min_mapq=20, qlen_lower_limit=30, qlen_upper_limit=900,
                     min_sequence_identity=0.9, min_gq=5, min_bq=30, min_trim=0.05,
                     max_mismatch_count=2, mismatch_window=8, md_threshold=45, min_ref_count=2)
    args.update(over)
    return batch, ref, args


@case("call_adversarial_a")
def _call_adv_a():
    batch, ref, args = _adv(101)
    return dict(kind="call", batch=batch, ref=ref, contig_len=len(ref), args=args,
                chunks=[(0, 1000), (1000, 2000), (2000, 3000)])


@case("call_adversarial_b")
def _call_adv_b():
    # overlapping / repeated regions (a --region_list with overlaps) and a strict window
    batch, ref, args = _adv(102, max_mismatch_count=0, mismatch_window=20, min_trim=0.01, min_ref_count=8)
    return dict(kind="call", batch=batch, ref=ref, contig_len=len(ref), args=args,
                chunks=[(0, 1500), (1200, 2400), (1200, 2400), (2399, 3000)])


@case("norm_basic")
def _norm_basic():
    d = _synth_case(60_000, 21)
    return dict(kind="norm", batch=d.batch, ref=d.ref.decode(), contig_len=60_000,
                chunks=chunkloci(0, 60_000), args=call_args())


@case("norm_sets")
def _norm_sets():
    d = _synth_case(50_000, 22, sub_err_rate=1e-3, somatic_rate=1e-4)
    common, pon = site_sets_from_synth(d, 22)
    return dict(kind="norm", batch=d.batch, ref=d.ref.decode(), contig_len=50_000,
                chunks=[(1, 20000), (20000, 49998)], args=call_args(min_bq=60), common=common, pon=pon)


@case("norm_phase")
def _norm_phase():
    d = _synth_case(60_000, 23, phase_block=20_000)
    ph, chunks, sets = phase_case(d, 20_000)
    return dict(kind="norm", batch=d.batch, ref=d.ref.decode(), contig_len=60_000,
                chunks=chunks, phase_sets=sets, phase=ph, args=call_args(phase=True))


@case("norm_adversarial")
def _norm_adv():
    batch, ref, args = _adv(103)
    ref = ref[:500] + ref[500:520].lower() + ref[520:]
    return dict(kind="norm", batch=batch, ref=ref, contig_len=len(ref), args=args,
                chunks=[(0, 1000), (1000, 2000), (2000, 3000)])


def build_case(name):
    c = CASES[name]()
    c["name"] = name
    c.setdefault("common", np.zeros(0, np.uint64))
    c.setdefault("pon", np.zeros(0, np.uint64))
    c.setdefault("phase", None)
    c.setdefault("phase_sets", None)
    # the reference only loads the sets when neither flag is given (caller.py:248-289)
    c["common_vcf"], c["pon_vcf"] = c["common"], c["pon"]
    if c["args"].get("create_panel_of_normals") or c["args"].get("non_human_sample"):
        c["common"], c["pon"] = np.zeros(0, np.uint64), np.zeros(0, np.uint64)
    c["chunk_table"] = c["batch"].chunk_table(c["chunks"], c["phase_sets"])
    c["params"] = gtmodel.make_params(**c["args"])
    return c


# ------------------------------------------------------------------------------------------
# randomised sweep (tests/test_gpu_random.py: CUDA == oracle; tests/golden/make_golden_random.py +
# tests/test_oracle_random.py: oracle == reference on the very same cases)
# ------------------------------------------------------------------------------------------
RANDOM_CALL_SEEDS = range(1000, 1030)
RANDOM_NORM_SEEDS = range(2000, 2030)


# a second family: dirty reads several 2048-position tiles long on a 30 kb contig (seeds >= LONG_SEED_BASE)
LONG_SEED_BASE = 4000
RANDOM_LONG_CALL_SEEDS = range(4000, 4010)
RANDOM_LONG_NORM_SEEDS = range(5000, 5008)
# long-read --phase seeds where two records sharing a query name disagree at a phase-checked site: the reference
# classifies the re-fetched records by query name (caller.py:556-567), DESIGN.md §7 says what the kernels do today
RANDOM_DUPNAME_CALL_SEEDS = (6081, 6090, 6123)


def random_setup(seed):
    rnd = random.Random(seed)
    long_reads = seed >= LONG_SEED_BASE
    if long_reads:
        n = rnd.choice([20_000, 30_000])
        batch, ref = adversarial_batch(seed, contig_len=n, n_reads=rnd.choice([60, 150, 220]), max_len=rnd.choice([2500, 5000, 9000]))
    else:
        n = rnd.choice([1500, 3000, 6000])
        batch, ref = adversarial_batch(seed, contig_len=n, n_reads=rnd.choice([40, 120, 300]), max_len=rnd.choice([300, 900, 2000]))
    args = call_args(
        min_qv=rnd.choice([0, 20, 30]), min_mapq=rnd.choice([0, 20, 60]),
        qlen_lower_limit=rnd.choice([0, 30, 2000] if long_reads else [0, 30, 200]),
        qlen_upper_limit=rnd.choice([4000, 12000, 100000] if long_reads else [500, 900, 5000]),
        min_sequence_identity=rnd.choice([0.0, 0.9, 0.99]),
        min_gq=rnd.choice([0, 5, 20]), min_bq=rnd.choice([1, 30, 93]), min_trim=rnd.choice([0.0, 0.01, 0.1]),
        max_mismatch_count=rnd.choice([0, 0, 1, 3]), mismatch_window=rnd.choice([0, 5, 20, 40]),
        md_threshold=rnd.choice([10, 45, 1000]), min_ref_count=rnd.choice([0, 2, 5]), min_alt_count=rnd.choice([1, 2]),
        min_hap_count=rnd.choice([0, 1, 3]), germline_snv_prior=rnd.choice([1e-3, 1e-2]))
    # chunk list: sometimes the reference's own tiling, sometimes overlapping / unordered windows
    if rnd.random() < 0.5:
        cuts = sorted(rnd.sample(range(1, n), rnd.choice([1, 2, 4])))
        edges = [0] + cuts + [n]
        chunks = [(edges[i], edges[i + 1]) for i in range(len(edges) - 1)]
    else:
        chunks = []
        for _ in range(rnd.choice([1, 3, 5])):
            a = rnd.randrange(0, n - 10)
            chunks.append((a, min(n, a + rnd.randrange(5, n))))
    # site sets drawn from positions that exist
    keys = [((rnd.randrange(1, n) << 4) | (rnd.randrange(4) << 2) | rnd.randrange(4)) for _ in range(n // 4)]
    common = np.unique(np.array(keys[: len(keys) // 2], np.uint64))
    pon = np.unique(np.array(keys[len(keys) // 2:], np.uint64))
    return rnd, batch, ref, args, chunks, common, pon


def random_phase(rnd, ref, chunks):
    """one phase set per chunk: random hetSNPs inside the chunk's window (some with non-matching alleles)"""
    hpos, href, halt, hbit, set_off, new_chunks = [], [], [], [], [0], []
    for (s, e) in chunks:
        cand = sorted(rnd.sample(range(max(s, 1), max(e, s + 2)), min(rnd.choice([2, 6, 20]), max(e - s - 1, 1))))
        cand = [p for p in cand if ref[p - 1] in "ATGC"]
        if len(cand) < 1:
            continue
        for p in cand:
            r = "ATGC".index(ref[p - 1])
            hpos.append(p); href.append(r); halt.append(rnd.choice([x for x in range(4) if x != r])); hbit.append(rnd.randrange(2))
        set_off.append(len(hpos))
        new_chunks.append((cand[0], cand[-1]))  # the reference's phase chunks: (first hpos, last hpos)
    ph = dict(hpos=np.array(hpos, np.int32), href=np.array(href, np.uint8), halt=np.array(halt, np.uint8),
              hbit=np.array(hbit, np.uint8), set_off=np.array(set_off, np.uint64))
    return ph, new_chunks


def random_case(kind, seed):
    """the case dict (as build_case returns it) of one seed of the sweep, or None when the seed draws no phase set"""
    rnd, batch, ref, args, chunks, common, pon = random_setup(seed)
    if kind == "norm":
        ref = "".join(c.lower() if rnd.random() < 0.02 else c for c in ref)
    phase, sets = None, None
    if seed % 3 == 0:
        phase, chunks = random_phase(rnd, ref.upper(), chunks)
        if not chunks:
            return None
        args["phase"] = True
        sets = list(range(len(chunks)))
    c = dict(name="rand_%s_%d" % (kind, seed), kind=kind, batch=batch, ref=ref, contig_len=len(ref), chunks=chunks,
             args=args, common=common, pon=pon, common_vcf=common, pon_vcf=pon, phase=phase, phase_sets=sets)
    c["chunk_table"] = batch.chunk_table(chunks, sets)
    c["params"] = gtmodel.make_params(**args)
    return c


# ------------------------------------------------------------------------------------------
# the whole CLI (tests/test_cli_dropin.py): three contigs whose names only sort right naturally
# ------------------------------------------------------------------------------------------
CLI_CONTIGS = [("chr1", 205_000, 31), ("chr2", 70_000, 32), ("chr10", 30_000, 33)]
CLI_DEPTH = 22.0  # the reference needs 2 - 8 us per aligned base: sized so that the CLI tests stay a minute or two in all


def cli_dataset(phase_block=None):
    """-> [(name, length, synth data)]: 22x CCS reads over a two-chunk contig and two short ones"""
    over = dict(somatic_rate=2e-5, depth=CLI_DEPTH)
    if phase_block:
        over["phase_block"] = phase_block
    return [(name, n, synth.generate(n, seed=seed, **over)) for name, n, seed in CLI_CONTIGS]
