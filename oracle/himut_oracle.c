/*
 * himut_oracle.c — CPU restatement of the reference's per-region calling hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker the CUDA path is compared against; it
 * may be imported / linked / executed only from tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  Nothing under himut_b200/ links it,
 * and the product path has no CPU fallback.
 *
 * What it restates (reference = sjin09/himut v1.0.4, pure Python; all arithmetic of the
 * path lives under src/himut/):
 *   cs op walk                       cslib.cs2tuple            src/himut/cslib.py:13-44
 *   substitution / mismatch lists    cslib.cs2subindel         src/himut/cslib.py:47-64
 *   read -> ref-position base map    cslib.cs2tpos2qbase       src/himut/cslib.py:153-170
 *   read statistics                  bamlib.BAM.get_qv, get_blast_sequence_identity
 *                                                              src/himut/bamlib.py:34-63
 *   candidate gates                  bamlib.get_tsbs_candidates, get_trimmed_range,
 *                                    is_trimmed, get_mismatch_range, is_mismatch_conflict
 *                                                              src/himut/bamlib.py:69-86,222-282
 *   pileup                           caller.update_allelecounts src/himut/caller.py:44-72
 *   genotype model                   gtlib.*                   src/himut/gtlib.py:9-174
 *   read haplotype                   haplib.get_ccs_hap/hbit   src/himut/haplib.py:46-83
 *   `call` worker                    caller.get_somatic_substitutions
 *                                                              src/himut/caller.py:243-641
 *   `normcounts` worker              normcounts.get_callable_tricounts, update_tri2count,
 *                                    get_tri_context           src/himut/normcounts.py:49-110,240-419
 *   reference tri counts             reflib.get_chrom_tricount src/himut/reflib.py:11-33
 * It keeps the reference's dataflow — per chunk, per read in fetch order, per-position
 * per-allele BQ lists in append order, left-to-right fp64 sums — so that the order-sensitive
 * results (PL, GQ) are reproduced bit for bit.  The three per-BQ log10 tables and the four
 * log10 priors come in through hm_params, evaluated by Python with the reference's formulas.
 *
 * Parity pinning: the reference ships no tests or golden vectors for this path
 * (SURVEY.md §4).  This restatement is pinned against outputs of the reference itself,
 * run in the build container through I/O shims: tests/golden/make_golden.py generates the
 * fixtures in tests/golden/, tests/test_oracle_golden.py checks this file against them;
 * tests/golden/make_golden_random.py does the same for a randomised sweep of dirty batches x
 * random parameters (81 committed seeds, tests/test_oracle_random.py; 3 400 more compared in
 * fuzzing sessions without a mismatch).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/himut_b200.h"

/* phased hetSNP table, same layout hm_set_phase_sets takes */
typedef struct orc_phase {
  const int32_t* hpos;
  const uint8_t* href;
  const uint8_t* halt;
  const uint8_t* hbit;
  uint64_t n_hetsnp;
  const uint64_t* set_off;
  uint64_t n_sets;
} orc_phase;

/* ---------------------------------------------------------------- op walk ------------ */
typedef struct {
  int kind, ref, alt;   /* ref/alt: base codes of a substitution */
  int ref_len, alt_len; /* tuple fields 4 and 5 of cs2tuple */
} op_t;

static inline op_t decode_op(uint32_t w) {
  op_t o;
  uint32_t v = w >> 2;
  o.kind = (int)(w & 3u);
  o.ref = o.alt = 0;
  switch (o.kind) {
    case HM_OP_MATCH: o.ref_len = (int)v; o.alt_len = (int)v; break;
    case HM_OP_SUB: o.ref = (int)(v & 7u); o.alt = (int)((v >> 3) & 7u); o.ref_len = 1; o.alt_len = 1; break;
    case HM_OP_INS: o.ref_len = 0; o.alt_len = (int)v; break;
    default: o.ref_len = (int)v; o.alt_len = 0; break;
  }
  return o;
}
static inline int read_base(const hm_read_batch* b, uint64_t r, int qpos) {
  const uint8_t* s = b->seq + b->seq_off[r];
  return (s[qpos >> 2] >> (2 * (qpos & 3))) & 3;
}
static inline int read_bq(const hm_read_batch* b, uint64_t r, int qpos) {
  return b->bq[b->bq_off[r] + (uint64_t)qpos];
}

/* ---------------------------------------------------------------- read statistics ---- */
/* bamlib.get_blast_sequence_identity (bamlib.py:47-63), BAM.get_qv (bamlib.py:34-36),
 * len(mismatch_lst) of cslib.cs2subindel (cslib.py:47-64) */
typedef struct {
  int64_t bq_total;
  int32_t n_match, n_sub, ins_len, del_len, n_mismatch;
} read_stats_t;

static read_stats_t read_stats(const hm_read_batch* b, uint64_t r) {
  read_stats_t s;
  memset(&s, 0, sizeof(s));
  const uint32_t* ops = b->ops + b->op_off[r];
  for (uint32_t k = 0; k < b->n_ops[r]; k++) {
    op_t o = decode_op(ops[k]);
    if (o.kind == HM_OP_MATCH) s.n_match += o.ref_len;
    else if (o.kind == HM_OP_SUB) { s.n_sub += o.alt_len; if (o.ref != (int)HM_BASE_N) s.n_mismatch++; }
    else if (o.kind == HM_OP_INS) { s.ins_len += o.alt_len; s.n_mismatch++; }
    else { s.del_len += o.ref_len; s.n_mismatch++; }
  }
  const uint8_t* q = b->bq + b->bq_off[r];
  for (int32_t i = 0; i < b->qlen[r]; i++) s.bq_total += q[i];
  return s;
}

int orc_read_stats(const hm_read_batch* b, int64_t* bq_total, int32_t* n_match, int32_t* n_sub,
                   int32_t* ins_len, int32_t* del_len, int32_t* n_mismatch) {
  for (uint64_t r = 0; r < b->n_reads; r++) {
    read_stats_t s = read_stats(b, r);
    if (bq_total) bq_total[r] = s.bq_total;
    if (n_match) n_match[r] = s.n_match;
    if (n_sub) n_sub[r] = s.n_sub;
    if (ins_len) ins_len[r] = s.ins_len;
    if (del_len) del_len[r] = s.del_len;
    if (n_mismatch) n_mismatch[r] = s.n_mismatch;
  }
  return HM_OK;
}

/* the four read gates of caller.py:310-317 in order; 1 = read passes */
static int read_passes(const hm_params* p, const hm_read_batch* b, uint64_t r) {
  read_stats_t s = read_stats(b, r);
  double qv = (double)s.bq_total / (double)b->qlen[r]; /* np.mean of exact ints */
  if (qv < (double)p->min_qv) return 0;
  if ((int)b->mapq[r] < p->min_mapq) return 0;
  int aln = s.n_match + s.n_sub + s.ins_len + s.del_len;
  double ident = (double)s.n_match / (double)aln;
  if (ident < p->min_sequence_identity) return 0;
  if (!(p->qlen_lower_limit < b->qlen[r] && b->qlen[r] < p->qlen_upper_limit)) return 0;
  return 1;
}

/* bamlib.get_trimmed_range / is_trimmed (bamlib.py:222-242) */
static inline void trimmed_range(int qlen, double min_trim, double* ts, double* te) {
  *ts = floor(min_trim * (double)qlen);
  *te = ceil((1.0 - min_trim) * (double)qlen);
}
static inline int is_trimmed(int qpos, double ts, double te) {
  return ((double)qpos < ts) || ((double)qpos > te);
}
/* bamlib.get_mismatch_range (bamlib.py:245-258) */
static inline void mismatch_range(int tpos, int qpos, int qlen, int w, int* lo, int* hi) {
  int qs = qpos - w, qe = qpos + w, u, d;
  if (qs < 0) { u = w + qs; d = w + abs(qs); }
  else if (qe > qlen) { u = w + abs(qe - qlen); d = qlen - qpos; }
  else { u = w; d = w; }
  *lo = tpos - u; *hi = tpos + d;
}
static inline int bisect_left_i32(const int32_t* a, int n, int x) {
  int lo = 0, hi = n;
  while (lo < hi) { int m = (lo + hi) >> 1; if (a[m] < x) lo = m + 1; else hi = m; }
  return lo;
}
static inline int bisect_right_i32(const int32_t* a, int n, int x) {
  int lo = 0, hi = n;
  while (lo < hi) { int m = (lo + hi) >> 1; if (x < a[m]) hi = m; else lo = m + 1; }
  return lo;
}

/* ---------------------------------------------------------------- genotype model ----- */
/* gtlib.gt_lst (gtlib.py:9) as base codes A0 T1 G2 C3 */
static const int GT_B1[10] = {0, 1, 3, 2, 1, 3, 2, 3, 2, 2};
static const int GT_B2[10] = {0, 0, 0, 0, 1, 1, 1, 3, 3, 2};
enum { ST_HOMREF = 0, ST_HET = 1, ST_HETALT = 2, ST_HOMALT = 3 };

/* gtlib.get_germ_gt_state (gtlib.py:23-38) */
static inline int gt_state(int b1, int b2, int ref) {
  if (b1 == ref && b2 == ref) return ST_HOMREF;
  if ((b1 == ref) != (b2 == ref)) return ST_HET;
  if (b1 != b2) return ST_HETALT;
  return ST_HOMALT;
}

/* per-allele BQ lists of one position, append order */
typedef struct {
  const uint8_t* bq[4];
  int n[4];
} bqlists_t;

/* gtlib.get_log10_gt_pD (gtlib.py:72-96) / the loop body of get_germ_gq (gtlib.py:146-165):
 * left-to-right sums per allele in the order A,T,G,C, then prior, then * -10.
 * skip < 0: nothing skipped (call form); else that allele is left out (normcounts form).
 * returns 1 if a BQ of 0 was met (log10(0) in the reference). */
static int gt_pl(const hm_params* p, const bqlists_t* L, int ref, int skip, double pl[10]) {
  int bad = 0;
  for (int g = 0; g < 10; g++) {
    int b1 = GT_B1[g], b2 = GT_B2[g];
    double acc = 0.0;
    for (int base = 0; base < 4; base++) {
      if (base == skip) continue;
      const double* lut;
      if (b1 == b2 && base == b1) lut = p->lut_hom;
      else if (b1 != b2 && (base == b1 || base == b2)) lut = p->lut_het;
      else lut = p->lut_err;
      double term = 0.0;
      for (int i = 0; i < L->n[base]; i++) {
        int q = L->bq[base][i];
        if (q == 0) bad = 1;
        term += lut[q];
      }
      acc += term;
    }
    acc += p->log10_prior[gt_state(b1, b2, ref)];
    pl[g] = -10.0 * acc;
  }
  return bad;
}

/* gtlib.get_argmin_gt (gtlib.py:113-119): stable ascending order, GQ = int(second - best)
 * capped at 99; *tie is set when the two smallest PL are equal. */
static int argmin_gt(const double pl[10], int* gq, int* tie) {
  int best = 0;
  for (int g = 1; g < 10; g++) if (pl[g] < pl[best]) best = g;
  double second = INFINITY;
  for (int g = 0; g < 10; g++) if (g != best && pl[g] < second) second = pl[g];
  double d = second - pl[best];
  *gq = d < 99.0 ? (int)d : 99;
  *tie = (second == pl[best]);
  return best;
}

/* caller.is_germ_gt (caller.py:111-147); g0,g1 = genotype after the ref-first flip */
static int is_germ_gt(int ref, int alt, int g0, int g1, int state, const int32_t c[6]) {
  if (state == ST_HET) return g0 == ref && g1 == alt;
  if (state == ST_HETALT) {
    int base_sum = c[0] + c[1] + c[2] + c[3];
    return base_sum == c[g0] + c[g1] && (alt == g0 || alt == g1);
  }
  if (state == ST_HOMALT) return c[ref] == 0 && g0 == alt && g1 == alt;
  return alt == g0;
}

/* ---------------------------------------------------------------- pileup of a chunk -- */
/* caller.update_allelecounts (caller.py:44-72) over every fetched primary read, kept as
 * counts[pos][6] plus per (pos, allele) lists of (BQ, read index) in append order. */
typedef struct {
  int32_t lo, hi;      /* covered reference window [lo, hi) */
  int32_t* counts;     /* 6 per position */
  uint32_t* off;       /* 4 per position + 1: CSR offsets of the BQ lists */
  uint8_t* bq;
  uint32_t* rd;        /* read index per list entry */
} pile_t;

static void pile_free(pile_t* P) { free(P->counts); free(P->off); free(P->bq); free(P->rd); memset(P, 0, sizeof(*P)); }

static int pile_build(pile_t* P, const hm_read_batch* b, const uint64_t* F, size_t nF) {
  memset(P, 0, sizeof(*P));
  if (nF == 0) return 0;
  int32_t lo = INT32_MAX, hi = INT32_MIN;
  for (size_t i = 0; i < nF; i++) {
    if (b->tstart[F[i]] < lo) lo = b->tstart[F[i]];
    if (b->tend[F[i]] > hi) hi = b->tend[F[i]];
  }
  hi += 1; /* an insertion after the last base lands at tend */
  size_t W = (size_t)(hi - lo);
  P->lo = lo; P->hi = hi;
  P->counts = (int32_t*)calloc(W * 6, sizeof(int32_t));
  P->off = (uint32_t*)calloc(W * 4 + 1, sizeof(uint32_t));
  if (!P->counts || !P->off) return -1;
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 1) {
      uint32_t acc = 0;
      for (size_t k = 0; k < W * 4; k++) { uint32_t n = P->off[k]; P->off[k] = acc; acc += n; }
      P->off[W * 4] = acc;
      P->bq = (uint8_t*)malloc(acc + 1);
      P->rd = (uint32_t*)malloc(((size_t)acc + 1) * 4);
      if (!P->bq || !P->rd) return -1;
    }
    for (size_t i = 0; i < nF; i++) {
      uint64_t r = F[i];
      const uint32_t* ops = b->ops + b->op_off[r];
      int tpos = b->tstart[r], qpos = b->qstart[r];
      for (uint32_t k = 0; k < b->n_ops[r]; k++) {
        op_t o = decode_op(ops[k]);
        if (o.kind == HM_OP_MATCH || o.kind == HM_OP_SUB) {
          for (int j = 0; j < o.ref_len; j++) {
            int a = o.kind == HM_OP_SUB ? o.alt : read_base(b, r, qpos + j);
            size_t cell = (size_t)(tpos + j - lo) * 4 + (size_t)a;
            if (pass == 0) { P->off[cell]++; P->counts[(size_t)(tpos + j - lo) * 6 + a]++; }
            else { uint32_t at = P->off[cell]++; P->bq[at] = (uint8_t)read_bq(b, r, qpos + j); P->rd[at] = (uint32_t)r; }
          }
        } else if (pass == 0) {
          if (o.kind == HM_OP_INS) P->counts[(size_t)(tpos - lo) * 6 + 4]++;
          else for (int j = 0; j < o.ref_len; j++) P->counts[(size_t)(tpos + j - lo) * 6 + 5]++;
        }
        tpos += o.ref_len; qpos += o.alt_len;
      }
    }
    if (pass == 1) { /* cursors ran to the next cell's start: shift back */
      for (size_t k = W * 4; k > 0; k--) P->off[k] = P->off[k - 1];
      P->off[0] = 0;
    }
  }
  return 0;
}
static inline void pile_lists(const pile_t* P, int32_t rpos, bqlists_t* L, int32_t c[6]) {
  if (!P->counts || rpos < P->lo || rpos >= P->hi) {
    memset(L, 0, sizeof(*L)); memset(c, 0, 6 * sizeof(int32_t));
    return;
  }
  size_t w = (size_t)(rpos - P->lo);
  for (int a = 0; a < 4; a++) { L->bq[a] = P->bq + P->off[w * 4 + a]; L->n[a] = (int)(P->off[w * 4 + a + 1] - P->off[w * 4 + a]); }
  memcpy(c, P->counts + w * 6, 6 * sizeof(int32_t));
}

/* ---------------------------------------------------------------- haplotype of a read  */
/* cslib.cs2tpos2qbase lookup of one 1-based position (cslib.py:153-170):
 * returns base code 0..3, 5 for a deleted base ("-"), -1 when the read does not cover it */
static int read_allele_at(const hm_read_batch* b, uint64_t r, int32_t pos1) {
  const uint32_t* ops = b->ops + b->op_off[r];
  int tpos = b->tstart[r], qpos = b->qstart[r];
  int rpos = pos1 - 1;
  for (uint32_t k = 0; k < b->n_ops[r]; k++) {
    op_t o = decode_op(ops[k]);
    if (o.ref_len > 0 && rpos >= tpos && rpos < tpos + o.ref_len) {
      if (o.kind == HM_OP_MATCH) return read_base(b, r, qpos + (rpos - tpos));
      if (o.kind == HM_OP_SUB) return o.alt;
      return 5;
    }
    tpos += o.ref_len; qpos += o.alt_len;
  }
  return -1;
}
/* haplib.get_ccs_hap / get_ccs_hbit (haplib.py:46-83): 0 / 1 / 2 (".") */
static int ccs_hap(const hm_read_batch* b, uint64_t r, const orc_phase* ph, int set) {
  if (!ph || set < 0 || (uint64_t)set >= ph->n_sets) return 2;
  uint64_t s0 = ph->set_off[set], s1 = ph->set_off[set + 1];
  const int32_t* hp = ph->hpos + s0;
  int n = (int)(s1 - s0);
  int idx = bisect_right_i32(hp, n, b->tstart[r]);
  int jdx = bisect_right_i32(hp, n, b->tend[r]);
  if (jdx - idx < 2) return 2;
  int is_h0 = 1, is_h1 = 1;
  for (int k = idx; k < jdx; k++) {
    int a = read_allele_at(b, r, hp[k]);
    int bit; /* 0 ref, 1 alt, 2 "-" */
    if (a >= 0 && a < 4 && a == (int)ph->href[s0 + k]) bit = 0;
    else if (a >= 0 && a < 4 && a == (int)ph->halt[s0 + k]) bit = 1;
    else bit = 2;
    int h0 = ph->hbit[s0 + k];
    if (bit != h0) is_h0 = 0;
    if (bit != 1 - h0) is_h1 = 0;
  }
  return is_h0 ? 0 : (is_h1 ? 1 : 2);
}

/* ---------------------------------------------------------------- small helpers ------ */
static inline int key_in(const uint64_t* a, size_t n, uint64_t key) {
  size_t lo = 0, hi = n;
  while (lo < hi) { size_t m = (lo + hi) >> 1; if (a[m] < key) lo = m + 1; else hi = m; }
  return lo < n && a[lo] == key;
}
static inline uint64_t site_key(int32_t tpos, int ref, int alt) {
  return ((uint64_t)(uint32_t)tpos << 4) | ((uint64_t)ref << 2) | (uint64_t)alt;
}
typedef struct { uint8_t* w; size_t n; } bitset_t;
static int bitset_init(bitset_t* s, size_t n) { s->n = n; s->w = (uint8_t*)calloc(n / 8 + 1, 1); return s->w ? 0 : -1; }
static inline int bitset_get(const bitset_t* s, size_t i) { return i < s->n && ((s->w[i >> 3] >> (i & 7)) & 1); }
static inline void bitset_set(bitset_t* s, size_t i) { if (i < s->n) s->w[i >> 3] |= (uint8_t)(1u << (i & 7)); }

static int cmp_u64(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return (x > y) - (x < y);
}

/* pysam fetch(chrom, start, end): records overlapping the 0-based half-open window, file order */
static size_t fetch_reads(const hm_read_batch* b, const hm_chunk* c, uint64_t* F) {
  size_t n = 0;
  for (uint64_t r = c->read_lo; r < c->read_hi && r < b->n_reads; r++)
    if (b->tstart[r] < c->end && b->tend[r] > c->start) F[n++] = r;
  return n;
}

/* ---------------------------------------------------------------- `himut call` ------- */
/* caller.get_somatic_substitutions (caller.py:243-641).  Emits one record per evaluated
 * candidate, germline restatements included (status HM_ST_GERM_*), in chunk order. */
/* qseen (may be NULL): one flag per qname_id, set where a record with that id passed the read gates — the set
 * m.num_ccs counts (caller.py:318-320), which a worker that feeds a contig in several batches has to carry */
int orc_call_chunks_seen(const hm_params* p, const hm_read_batch* b, const hm_chunk* chunks,
                         size_t n_chunks, const uint64_t* common, size_t n_common,
                         const uint64_t* pon, size_t n_pon, const orc_phase* ph,
                         hm_site_record* out, size_t cap, size_t* n_out,
                         int64_t log[HM_CALL_LOG_LEN], uint8_t* qseen, size_t qseen_cap) {
  int rc = HM_OK, bq_zero = 0;
  size_t nrec = 0;
  memset(log, 0, sizeof(int64_t) * HM_CALL_LOG_LEN);
  int32_t max_tend = 0;
  uint32_t max_qid = 0;
  size_t max_fetch = 1;
  for (uint64_t r = 0; r < b->n_reads; r++) {
    if (b->tend[r] > max_tend) max_tend = b->tend[r];
    if (b->qname_id[r] > max_qid) max_qid = b->qname_id[r];
  }
  for (size_t c = 0; c < n_chunks; c++)
    if (chunks[c].read_hi > chunks[c].read_lo && (size_t)(chunks[c].read_hi - chunks[c].read_lo) > max_fetch)
      max_fetch = chunks[c].read_hi - chunks[c].read_lo;
  bitset_t som_seen, ccs_seen;
  if (bitset_init(&som_seen, (size_t)max_tend + 4) || bitset_init(&ccs_seen, (size_t)max_qid + 1)) return HM_ERR_ARG;
  uint64_t* F = (uint64_t*)malloc(max_fetch * sizeof(uint64_t));
  uint8_t* hap = (uint8_t*)malloc(b->n_reads + 1);
  int32_t* mm = NULL; size_t mm_cap = 0;
  uint64_t* cand = NULL; size_t cand_cap = 0;
  /* reads overlapping a single position: prefix maximum of tend, monotone */
  int32_t* pmax = (int32_t*)malloc((b->n_reads + 1) * sizeof(int32_t));
  if (!F || !hap || !pmax) { rc = HM_ERR_ARG; goto done; }
  { int32_t m = INT32_MIN; for (uint64_t r = 0; r < b->n_reads; r++) { if (b->tend[r] > m) m = b->tend[r]; pmax[r] = m; } }

  for (size_t ci = 0; ci < n_chunks; ci++) {
    const hm_chunk* c = &chunks[ci];
    size_t nF = fetch_reads(b, c, F), nP = 0;
    /* BAM(i): skip secondary (bamlib.py:17) */
    for (size_t i = 0; i < nF; i++) if (!(b->flags[F[i]] & HM_READ_SECONDARY)) F[nP++] = F[i];
    nF = nP;
    pile_t P;
    if (pile_build(&P, b, F, nF)) { pile_free(&P); rc = HM_ERR_ARG; goto done; }
    size_t ncand = 0;
    for (size_t i = 0; i < nF; i++) {
      uint64_t r = F[i];
      if (p->phase) {
        hap[r] = (uint8_t)ccs_hap(b, r, ph, c->phase_set);
        if (hap[r] > 1) continue;
      }
      if (!read_passes(p, b, r)) continue;
      if (!bitset_get(&ccs_seen, b->qname_id[r])) { log[0]++; bitset_set(&ccs_seen, b->qname_id[r]); }
      /* cs2subindel: mismatch list (1-based), then the candidate gates per substitution */
      uint32_t nops = b->n_ops[r];
      if (nops > mm_cap) { mm_cap = nops * 2 + 64; mm = (int32_t*)realloc(mm, mm_cap * sizeof(int32_t)); if (!mm) { rc = HM_ERR_ARG; pile_free(&P); goto done; } }
      const uint32_t* ops = b->ops + b->op_off[r];
      int nmm = 0, tpos = b->tstart[r], qpos = b->qstart[r];
      for (uint32_t k = 0; k < nops; k++) {
        op_t o = decode_op(ops[k]);
        if ((o.kind == HM_OP_SUB && o.ref != (int)HM_BASE_N) || o.kind == HM_OP_INS || o.kind == HM_OP_DEL) mm[nmm++] = tpos + 1;
        tpos += o.ref_len; qpos += o.alt_len;
      }
      double ts, te;
      trimmed_range(b->qlen[r], p->min_trim, &ts, &te);
      tpos = b->tstart[r]; qpos = b->qstart[r];
      for (uint32_t k = 0; k < nops; k++) {
        op_t o = decode_op(ops[k]);
        if (o.kind == HM_OP_SUB && o.ref != (int)HM_BASE_N) {
          int t1 = tpos + 1, lo, hi;
          if (!bitset_get(&som_seen, (size_t)t1) && !is_trimmed(qpos, ts, te)) {
            mismatch_range(t1, qpos, b->qlen[r], p->mismatch_window, &lo, &hi);
            int cnt = bisect_right_i32(mm, nmm, hi) - bisect_left_i32(mm, nmm, lo) - 1;
            if (!(cnt > p->max_mismatch_count)) {
              if (ncand == cand_cap) { cand_cap = cand_cap * 2 + 1024; cand = (uint64_t*)realloc(cand, cand_cap * sizeof(uint64_t)); if (!cand) { rc = HM_ERR_ARG; pile_free(&P); goto done; } }
              cand[ncand++] = site_key(t1, o.ref, o.alt);
            }
          }
        }
        tpos += o.ref_len; qpos += o.alt_len;
      }
    }
    /* set(somatic_tsbs_candidate_lst) */
    if (ncand) qsort(cand, ncand, sizeof(uint64_t), cmp_u64);
    size_t nu = 0;
    for (size_t i = 0; i < ncand; i++) if (i == 0 || cand[i] != cand[i - 1]) cand[nu++] = cand[i];

    for (size_t i = 0; i < nu; i++) {
      int32_t tpos = (int32_t)(cand[i] >> 4);
      int ref = (int)((cand[i] >> 2) & 3), alt = (int)(cand[i] & 3);
      if (!(c->start <= tpos && tpos <= c->end)) continue; /* is_chunk */
      log[1]++;
      int32_t rpos = tpos - 1;
      bqlists_t L; int32_t cnt[6];
      pile_lists(&P, rpos, &L, cnt);
      int ins = cnt[4], del = cnt[5];
      int depth = cnt[0] + cnt[1] + cnt[2] + cnt[3] + cnt[4] + cnt[5] - ins; /* get_read_depth */
      int ref_count = cnt[ref], alt_count = cnt[alt];
      double pl[10];
      bq_zero |= gt_pl(p, &L, ref, -1, pl);
      int gq, tie;
      int best = argmin_gt(pl, &gq, &tie);
      int g0 = GT_B1[best], g1 = GT_B2[best];
      int state = gt_state(g0, g1, ref);
      if (g0 != ref && ((g0 == ref) + (g1 == ref)) == 1) { int t = g0; g0 = g1; g1 = t; }
      hm_site_record R;
      memset(&R, 0, sizeof(R));
      R.tpos = tpos; R.ref = (uint8_t)ref; R.alt = (uint8_t)alt; R.chunk = (int32_t)ci;
      R.flags = tie ? HM_SITE_PL_TIE : 0;
      R.germ_gt[0] = (uint8_t)g0; R.germ_gt[1] = (uint8_t)g1; R.germ_state = (uint8_t)state;
      R.phase_set = -1;
      for (int a = 0; a < 6; a++) R.counts[a] = cnt[a];
      for (int a = 0; a < 4; a++) { int s = 0; for (int k = 0; k < L.n[a]; k++) s += L.bq[a][k]; R.bq_sum[a] = s; }
      /* get_germ_gq with the 2-char som_gt skips nothing (caller.py:348): same PL, same GQ */
      R.gq = gq;
      int status;
      if (is_germ_gt(ref, alt, g0, g1, state, cnt)) {
        if (state == ST_HET) { log[2]++; status = HM_ST_GERM_HET; }
        else if (state == ST_HETALT) { log[3]++; status = HM_ST_GERM_HETALT; }
        else if (state == ST_HOMALT) { log[4]++; status = HM_ST_GERM_HOMALT; }
        else status = HM_ST_GERM_HOMREF;
      } else {
        bitset_set(&som_seen, (size_t)tpos);
        if (state == ST_HET) { log[5]++; status = HM_ST_HET_SITE; }
        else if (state == ST_HETALT) { log[5]++; status = HM_ST_HETALT_SITE; }
        else if (state == ST_HOMALT) { log[5]++; status = HM_ST_HOMALT_SITE; }
        else if (del != 0 || ins != 0) { log[7]++; status = HM_ST_INDEL_SITE; }
        else {
          log[6]++;
          int hi_bq = 0; /* is_low_bq (caller.py:160-171) */
          for (int k = 0; k < L.n[alt]; k++) if ((int)L.bq[alt][k] >= p->min_bq) hi_bq++;
          uint64_t key = site_key(tpos, ref, alt);
          if (gq < p->min_gq) { log[8]++; status = HM_ST_LOW_GQ; }
          else if (hi_bq == 0) { log[9]++; status = HM_ST_LOW_BQ; }
          else if (!p->non_human_sample && !p->create_panel_of_normals && key_in(pon, n_pon, key)) { log[10]++; status = HM_ST_PON; }
          else if (!p->non_human_sample && key_in(common, n_common, key)) { log[11]++; status = HM_ST_COMSNP; }
          else if (!(ref_count >= p->min_ref_count && alt_count >= p->min_alt_count)) { log[13]++; status = HM_ST_LOW_DEPTH; }
          else if ((double)depth > p->md_threshold) { log[12]++; status = HM_ST_HIGH_DEPTH; }
          else {
            log[14]++;
            status = HM_ST_PASS;
            if (p->phase) { /* caller.py:552-603 */
              size_t w = (size_t)(rpos - P.lo);
              const uint32_t* wt = P.rd + P.off[w * 4 + ref]; int nwt = L.n[ref];
              const uint32_t* al = P.rd + P.off[w * 4 + alt]; int nal = L.n[alt];
              int h[3] = {0, 0, 0}, som_mask = 0;
              /* alignments.fetch(chrom, tpos, tpos + 1) over the whole file */
              uint64_t lo_r = 0, hi_r = b->n_reads;
              { uint64_t a = 0, z = b->n_reads; while (a < z) { uint64_t m = (a + z) >> 1; if (pmax[m] > tpos) z = m; else a = m + 1; } lo_r = a; }
              for (uint64_t r = lo_r; r < hi_r && b->tstart[r] < tpos + 1; r++) {
                if (!(b->tend[r] > tpos)) continue;
                if (b->flags[r] & HM_READ_SECONDARY) continue;
                uint32_t q = b->qname_id[r];
                int in_wt = 0, in_alt = 0;
                for (int k = 0; k < nwt && !in_wt; k++) in_wt = b->qname_id[wt[k]] == q;
                if (!in_wt) for (int k = 0; k < nal && !in_alt; k++) in_alt = b->qname_id[al[k]] == q;
                if (in_wt) h[ccs_hap(b, r, ph, c->phase_set)]++;
                else if (in_alt) { int hh = ccs_hap(b, r, ph, c->phase_set); if (hh < 2) som_mask |= 1 << hh; }
              }
              R.hap_count[0] = h[0]; R.hap_count[1] = h[1]; R.som_hap_mask = som_mask;
              if (h[0] >= p->min_hap_count && h[1] >= p->min_hap_count && (som_mask == 1 || som_mask == 2))
                R.phase_set = c->start;
              else status = HM_ST_UNPHASED;
            }
          }
        }
      }
      R.status = (uint8_t)status;
      if (nrec < cap) out[nrec] = R;
      nrec++;
    }
    pile_free(&P);
  }
done:
  if (qseen)
    for (size_t q = 0; q <= (size_t)max_qid && q < qseen_cap; q++) qseen[q] = (uint8_t)bitset_get(&ccs_seen, q);
  free(F); free(hap); free(mm); free(cand); free(pmax);
  free(som_seen.w); free(ccs_seen.w);
  if (n_out) *n_out = nrec;
  if (rc != HM_OK) return rc;
  if (bq_zero) return HM_ERR_BQ_ZERO;
  if (nrec > cap) return HM_ERR_CAPACITY;
  return HM_OK;
}

int orc_call_chunks(const hm_params* p, const hm_read_batch* b, const hm_chunk* chunks,
                    size_t n_chunks, const uint64_t* common, size_t n_common,
                    const uint64_t* pon, size_t n_pon, const orc_phase* ph,
                    hm_site_record* out, size_t cap, size_t* n_out,
                    int64_t log[HM_CALL_LOG_LEN]) {
  return orc_call_chunks_seen(p, b, chunks, n_chunks, common, n_common, pon, n_pon, ph, out, cap, n_out, log, NULL, 0);
}

/* ---------------------------------------------------------------- `himut normcounts` - */
/* normcounts.get_tri_context (normcounts.py:49-62) -> bin in mutlib.tri_lst order, 32 = other */
static int tri_bin(const uint8_t* seq, size_t n, int64_t pos) {
  if (pos < 1 || (size_t)pos + 1 >= n) return 32; /* "NNN" */
  int t[3];
  for (int k = 0; k < 3; k++) {
    switch (seq[pos - 1 + k]) { case 'A': t[k] = 0; break; case 'C': t[k] = 1; break; case 'G': t[k] = 2; break; case 'T': t[k] = 3; break; default: t[k] = -1; }
  }
  if (t[1] == 0 || t[1] == 2) { /* purine centre: reverse complement */
    int u0 = t[2] < 0 ? -1 : 3 - t[2], u1 = 3 - t[1], u2 = t[0] < 0 ? -1 : 3 - t[0];
    t[0] = u0; t[1] = u1; t[2] = u2;
  }
  if (t[0] < 0 || t[1] < 0 || t[2] < 0) return 32;
  return t[0] * 8 + (t[1] == 3 ? 4 : 0) + t[2]; /* centre is C (1) or T (3) */
}
static inline int ascii2code(uint8_t ch) {
  switch (ch) { case 'A': return 0; case 'T': return 1; case 'G': return 2; case 'C': return 3; default: return -1; }
}

/* alt_order[ref][0..2]: iteration order of list(base_set.difference(ref)) in the reference
 * (normcounts.py:370) — hash-seed dependent there; pass NULL for the canonical A,T,G,C order. */
int orc_normcounts_chunks_seen(const hm_params* p, const hm_read_batch* b, const uint8_t* refseq,
                               size_t ref_len, const hm_chunk* chunks, size_t n_chunks,
                               const uint64_t* common, size_t n_common, const uint64_t* pon,
                               size_t n_pon, const orc_phase* ph, const uint8_t* alt_order,
                               int64_t ccs_tri[HM_TRI_BINS], int64_t ref_tri[HM_TRI_BINS],
                               int64_t log[HM_NORM_LOG_LEN], int64_t* n_alt_tie, uint8_t* qseen, size_t qseen_cap) {
  int rc = HM_OK, bq_zero = 0;
  int64_t ties = 0;
  memset(ccs_tri, 0, sizeof(int64_t) * HM_TRI_BINS);
  memset(ref_tri, 0, sizeof(int64_t) * HM_TRI_BINS);
  memset(log, 0, sizeof(int64_t) * HM_NORM_LOG_LEN);
  uint32_t max_qid = 0;
  size_t max_fetch = 1;
  for (uint64_t r = 0; r < b->n_reads; r++) if (b->qname_id[r] > max_qid) max_qid = b->qname_id[r];
  for (size_t c = 0; c < n_chunks; c++)
    if (chunks[c].read_hi > chunks[c].read_lo && (size_t)(chunks[c].read_hi - chunks[c].read_lo) > max_fetch)
      max_fetch = chunks[c].read_hi - chunks[c].read_lo;
  bitset_t seen;
  if (bitset_init(&seen, (size_t)max_qid + 1)) return HM_ERR_ARG;
  uint64_t* F = (uint64_t*)malloc(max_fetch * sizeof(uint64_t));
  uint8_t* hap = (uint8_t*)malloc(b->n_reads + 1);
  int32_t* mm = NULL; size_t mm_cap = 0;
  if (!F || !hap) { rc = HM_ERR_ARG; goto done; }

  for (size_t ci = 0; ci < n_chunks; ci++) {
    const hm_chunk* c = &chunks[ci];
    size_t nF = fetch_reads(b, c, F), nP = 0;
    for (size_t i = 0; i < nF; i++) if (!(b->flags[F[i]] & HM_READ_SECONDARY)) F[nP++] = F[i];
    nF = nP;
    pile_t P;
    if (pile_build(&P, b, F, nF)) { pile_free(&P); rc = HM_ERR_ARG; goto done; }
    size_t W = nF ? (size_t)(P.hi - P.lo) : 0;
    int32_t* callable = (int32_t*)calloc(W + 1, sizeof(int32_t)); /* rpos2count */
    int32_t* hapcnt = p->phase ? (int32_t*)calloc(W * 2 + 2, sizeof(int32_t)) : NULL; /* rpos2hap2count["0"/"1"] */
    if (!callable || (p->phase && !hapcnt)) { free(callable); free(hapcnt); pile_free(&P); rc = HM_ERR_ARG; goto done; }
    for (size_t i = 0; i < nF; i++) {
      uint64_t r = F[i];
      const uint32_t* ops = b->ops + b->op_off[r];
      uint32_t nops = b->n_ops[r];
      if (p->phase) {
        hap[r] = (uint8_t)ccs_hap(b, r, ph, c->phase_set);
        if (hap[r] < 2) { /* update_phased_allelecounts: hap counts on match / sub positions */
          int tpos = b->tstart[r];
          for (uint32_t k = 0; k < nops; k++) {
            op_t o = decode_op(ops[k]);
            if (o.kind == HM_OP_MATCH || o.kind == HM_OP_SUB)
              for (int j = 0; j < o.ref_len; j++) hapcnt[(size_t)(tpos + j - P.lo) * 2 + hap[r]]++;
            tpos += o.ref_len;
          }
        }
        if (hap[r] > 1) continue;
      }
      if (!read_passes(p, b, r)) continue;
      if (!bitset_get(&seen, b->qname_id[r])) { log[0]++; bitset_set(&seen, b->qname_id[r]); }
      /* update_tri2count (normcounts.py:65-110) */
      if (nops > mm_cap) { mm_cap = nops * 2 + 64; mm = (int32_t*)realloc(mm, mm_cap * sizeof(int32_t)); if (!mm) { free(callable); free(hapcnt); pile_free(&P); rc = HM_ERR_ARG; goto done; } }
      int nmm = 0, rpos = b->tstart[r], qpos = b->qstart[r];
      for (uint32_t k = 0; k < nops; k++) {
        op_t o = decode_op(ops[k]);
        if ((o.kind == HM_OP_SUB && o.ref != (int)HM_BASE_N) || o.kind == HM_OP_INS || o.kind == HM_OP_DEL) mm[nmm++] = rpos + 1;
        rpos += o.ref_len; qpos += o.alt_len;
      }
      double ts, te;
      trimmed_range(b->qlen[r], p->min_trim, &ts, &te);
      rpos = b->tstart[r]; qpos = b->qstart[r];
      for (uint32_t k = 0; k < nops; k++) {
        op_t o = decode_op(ops[k]);
        if (o.kind == HM_OP_MATCH) {
          int lo, hi;
          mismatch_range(rpos, qpos, b->qlen[r], p->mismatch_window, &lo, &hi);
          for (int j = 0; j < o.ref_len; j++) {
            int cnt = bisect_right_i32(mm, nmm, hi + j) - bisect_left_i32(mm, nmm, lo + j);
            if (read_bq(b, r, qpos + j) < p->min_bq) continue;
            if (cnt > p->max_mismatch_count) continue;
            if (is_trimmed(qpos + j, ts, te)) continue;
            callable[rpos + j - P.lo]++;
          }
        } else if (o.kind == HM_OP_SUB) {
          callable[rpos - P.lo]++;
        }
        rpos += o.ref_len; qpos += o.alt_len;
      }
    }
    /* position loop (normcounts.py:317-400) */
    for (int64_t rpos = c->start; rpos < c->end; rpos++) {
      if (rpos < 0 || (size_t)rpos >= ref_len) continue;
      int ridx = ascii2code(refseq[rpos]);
      if (ridx < 0) continue;
      if (!nF || rpos < P.lo || rpos >= P.hi) continue;
      int64_t tri_sum = callable[rpos - P.lo];
      if (tri_sum == 0) continue;
      log[1] += tri_sum;
      if (p->phase) {
        int h0 = hapcnt[(size_t)(rpos - P.lo) * 2], h1 = hapcnt[(size_t)(rpos - P.lo) * 2 + 1];
        if (!(h0 >= p->min_hap_count && h1 >= p->min_hap_count)) { log[2] += tri_sum; continue; }
      }
      bqlists_t L; int32_t cnt[6];
      pile_lists(&P, (int32_t)rpos, &L, cnt);
      int ins = cnt[4], del = cnt[5];
      int depth = cnt[0] + cnt[1] + cnt[2] + cnt[3] + cnt[5];
      double pl[10];
      bq_zero |= gt_pl(p, &L, ridx, -1, pl);
      int gq, tie;
      int best = argmin_gt(pl, &gq, &tie);
      int state = gt_state(GT_B1[best], GT_B2[best], ridx);
      if (state == ST_HET) { log[3] += tri_sum; continue; }
      if (state == ST_HETALT) { log[4] += tri_sum; continue; }
      if (state == ST_HOMALT) { log[5] += tri_sum; continue; }
      log[6] += tri_sum;
      if (del != 0 || ins != 0) { log[7] += tri_sum; continue; }
      if ((double)depth > p->md_threshold) { log[8] += tri_sum; continue; }
      int ref_count = cnt[ridx];
      int tri = tri_bin(refseq, ref_len, rpos);
      if (depth == ref_count) {
        if (gq < p->min_gq) { log[10] += tri_sum; continue; }
        if (ref_count < p->min_ref_count) { log[9] += tri_sum; continue; }
      } else {
        int order[3], no = 0, filtered = 0;
        if (alt_order) for (int k = 0; k < 3; k++) order[no++] = alt_order[ridx * 3 + k];
        else for (int a = 0; a < 4; a++) if (a != ridx) order[no++] = a;
        int n_hit_kinds = 0;
        for (int k = 0; k < 3 && !filtered; k++) {
          int alt = order[k];
          if (cnt[alt] == 0) continue;
          uint64_t key = site_key((int32_t)rpos + 1, ridx, alt);
          if (!p->non_human_sample && key_in(pon, n_pon, key)) { log[11] += tri_sum; filtered = 1; }
          else if (!p->non_human_sample && key_in(common, n_common, key)) { log[12] += tri_sum; filtered = 1; }
        }
        (void)n_hit_kinds;
        if (filtered) continue;
        int alt = order[0], nmax = 1;
        for (int k = 1; k < 3; k++) {
          if (cnt[order[k]] > cnt[alt]) { alt = order[k]; nmax = 1; }
          else if (cnt[order[k]] == cnt[alt]) nmax++;
        }
        if (nmax > 1) ties++;
        bq_zero |= gt_pl(p, &L, ridx, alt, pl); /* get_germ_gq(alt 1-char): that allele is skipped */
        int gq2, tie2;
        argmin_gt(pl, &gq2, &tie2);
        if (gq2 < p->min_gq) { log[10] += tri_sum; continue; }
        if (!(ref_count >= p->min_ref_count && cnt[alt] >= p->min_alt_count)) { log[9] += tri_sum; continue; }
      }
      ref_tri[tri] += 1;
      ccs_tri[tri] += tri_sum;
      log[13] += tri_sum;
    }
    free(callable); free(hapcnt);
    pile_free(&P);
  }
done:
  if (qseen) /* as in orc_call_chunks_seen */
    for (size_t q = 0; q <= (size_t)max_qid && q < qseen_cap; q++) qseen[q] = (uint8_t)bitset_get(&seen, q);
  free(F); free(hap); free(mm); free(seen.w);
  if (n_alt_tie) *n_alt_tie = ties;
  if (rc != HM_OK) return rc;
  return bq_zero ? HM_ERR_BQ_ZERO : HM_OK;
}

int orc_normcounts_chunks(const hm_params* p, const hm_read_batch* b, const uint8_t* refseq,
                          size_t ref_len, const hm_chunk* chunks, size_t n_chunks,
                          const uint64_t* common, size_t n_common, const uint64_t* pon,
                          size_t n_pon, const orc_phase* ph, const uint8_t* alt_order,
                          int64_t ccs_tri[HM_TRI_BINS], int64_t ref_tri[HM_TRI_BINS],
                          int64_t log[HM_NORM_LOG_LEN], int64_t* n_alt_tie) {
  return orc_normcounts_chunks_seen(p, b, refseq, ref_len, chunks, n_chunks, common, n_common, pon, n_pon, ph, alt_order,
                                    ccs_tri, ref_tri, log, n_alt_tie, NULL, 0);
}

/* ---------------------------------------------------------------- reference tri counts */
/* reflib.get_chrom_tricount (src/himut/reflib.py:11-33) */
int orc_ref_tricounts(const uint8_t* seq, size_t n, int64_t tri[HM_TRI_BINS]) {
  memset(tri, 0, sizeof(int64_t) * HM_TRI_BINS);
  for (size_t i = 0; i + 2 < n; i++) {
    if (seq[i] == 'N') continue;
    int t[3];
    for (int k = 0; k < 3; k++) {
      switch (seq[i + k]) { case 'A': t[k] = 0; break; case 'C': t[k] = 1; break; case 'G': t[k] = 2; break; case 'T': t[k] = 3; break; default: t[k] = -1; }
    }
    if (t[1] == 0 || t[1] == 2) {
      int u0 = t[2] < 0 ? -1 : 3 - t[2], u1 = 3 - t[1], u2 = t[0] < 0 ? -1 : 3 - t[0];
      t[0] = u0; t[1] = u1; t[2] = u2;
    }
    tri[(t[0] < 0 || t[1] < 0 || t[2] < 0) ? 32 : t[0] * 8 + (t[1] == 3 ? 4 : 0) + t[2]]++;
  }
  return HM_OK;
}

/* ---------------------------------------------------------------- `himut phase` edges */
/* phaselib.get_edges (src/himut/phaselib.py:16-67) over one packed batch: same band layout as
 * hm_phase_edges_* (include/himut_b200.h).  Returns the widest pair distance seen (0 = fits). */
static int read_bq_at(const hm_read_batch* b, uint64_t r, int32_t pos1) {
  const uint32_t* ops = b->ops + b->op_off[r];
  int tpos = b->tstart[r], qpos = b->qstart[r];
  int rpos = pos1 - 1;
  for (uint32_t k = 0; k < b->n_ops[r]; k++) {
    op_t o = decode_op(ops[k]);
    if (o.ref_len > 0 && rpos >= tpos && rpos < tpos + o.ref_len) {
      if (o.kind == HM_OP_MATCH) return b->bq[b->bq_off[r] + (uint64_t)(qpos + (rpos - tpos))];
      if (o.kind == HM_OP_SUB) return b->bq[b->bq_off[r] + (uint64_t)qpos];
      return 0; /* deleted base: ("-", 0) */
    }
    tpos += o.ref_len; qpos += o.alt_len;
  }
  return 0;
}
uint32_t orc_phase_edges(const hm_read_batch* b, const int32_t* hpos, const uint8_t* href, size_t n_snp, int32_t min_bq,
                         int32_t min_mapq, int32_t min_tstart, uint32_t band, uint32_t* counts) {
  uint32_t need = 0;
  for (uint64_t r = 0; r < b->n_reads; r++) {
    if (b->flags[r] & HM_READ_SECONDARY) continue;
    if ((int32_t)b->mapq[r] < min_mapq) continue;
    if (b->tstart[r] < min_tstart) continue;
    int idx = bisect_right_i32(hpos, (int)n_snp, b->tstart[r]);
    int jdx = bisect_right_i32(hpos, (int)n_snp, b->tend[r]);
    if (jdx - idx < 2) continue;
    if ((uint32_t)(jdx - idx - 1) > band) { if ((uint32_t)(jdx - idx - 1) > need) need = (uint32_t)(jdx - idx - 1); continue; }
    for (int x = idx; x < jdx; x++) {
      if (read_bq_at(b, r, hpos[x]) < min_bq) continue;
      int ax = read_allele_at(b, r, hpos[x]);
      int sx = (ax >= 0 && ax < 4 && ax == (int)href[x]) ? 0 : 1;
      for (int y = x + 1; y < jdx; y++) {
        if (read_bq_at(b, r, hpos[y]) < min_bq) continue;
        int ay = read_allele_at(b, r, hpos[y]);
        int sy = (ay >= 0 && ay < 4 && ay == (int)href[y]) ? 0 : 1;
        int kind = sx == 0 ? (sy == 0 ? 0 : 2) : (sy == 1 ? 1 : 3);
        counts[((uint64_t)x * band + (uint64_t)(y - x - 1)) * 4 + kind]++;
      }
    }
  }
  return need;
}

