/*
 * synth.c — deterministic synthetic CCS data for tests and bench.py (no network, no
 * datasets): a random reference contig, a diploid germline (het + hom-alt SNVs, phased in
 * blocks), and coordinate-sorted ~15 kb HiFi-like reads carrying minimap2-style cs ops,
 * written straight into the packed hm_read_batch layout of include/himut_b200.h.
 *
 * Shape follows SURVEY.md §8(d) / Appendix C: read lengths have spread (the reference's
 * qlen gate is strict mean±2σ, bamlib.py:168-175), BQ >= 1 everywhere, bases are ACGT only.
 * Host-side only; nothing here is on the calling path.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/himut_b200.h"

typedef struct hm_synth_spec {
  uint64_t seed;
  int32_t contig_len;
  int32_t read_len_min;
  int32_t read_len_max;
  int32_t phase_block; /* span of one phase set */
  double depth;
  double read_len_mean;
  double read_len_sd;
  double het_rate;      /* per reference bp */
  double hom_rate;      /* per reference bp */
  double somatic_rate;  /* per read base: single-read substitutions at BQ 93 */
  double sub_err_rate;  /* per read base: sequencing substitution errors */
  double indel_rate;    /* per read base: 1-3 bp insertions / deletions */
  double lowq_frac;     /* P(BQ != 93) for ordinary bases */
  double mapq_low_frac; /* P(MAPQ < 60) */
  double softclip_frac; /* fraction of reads with 50-500 bp soft clips */
} hm_synth_spec;

typedef struct hm_synth_data {
  hm_read_batch batch;
  uint8_t* ref;         /* ASCII, contig_len bytes */
  uint64_t ref_len;
  uint64_t n_germ;      /* germline SNVs, ascending position */
  int32_t* germ_pos;    /* 1-based */
  uint8_t* germ_ref;    /* base codes */
  uint8_t* germ_alt;
  uint8_t* germ_gt;     /* 0: "1|0" (alt on hap 0), 1: "0|1" (alt on hap 1), 2: "1/1" */
  uint64_t n_som;       /* somatic single-read substitutions (truth) */
  int32_t* som_pos;     /* 1-based */
  uint8_t* som_ref;
  uint8_t* som_alt;
  uint64_t n_err;       /* sequencing substitution errors */
  int32_t* err_pos;
  uint8_t* err_ref;
  uint8_t* err_alt;
  uint64_t aligned_bases; /* sum of query_alignment_end - query_alignment_start */
  /* owned storage behind batch.* */
  int32_t *tstart, *tend, *qstart, *qlen;
  uint8_t *mapq, *flags;
  uint32_t* qname_id;
  uint64_t *seq_off, *bq_off, *op_off;
  uint32_t* n_ops;
  uint8_t *seq, *bq;
  uint32_t* ops;
} hm_synth_data;

/* ---- rng: splitmix64 seeding + xoshiro256** ---------------------------------------- */
typedef struct { uint64_t s[4]; } rng_t;
static uint64_t splitmix(uint64_t* x) {
  uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void rng_seed(rng_t* r, uint64_t seed, uint64_t stream) {
  uint64_t x = seed ^ (stream * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull);
  for (int i = 0; i < 4; i++) r->s[i] = splitmix(&x);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t* r) {
  uint64_t* s = r->s;
  uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
  return result;
}
static inline double rng_unif(rng_t* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t rng_below(rng_t* r, uint32_t n) { return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32); }
static double rng_normal(rng_t* r) {
  double u1 = rng_unif(r), u2 = rng_unif(r);
  if (u1 < 1e-300) u1 = 1e-300;
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
/* gap to the next Bernoulli(p) success, >= 1 */
static inline int64_t rng_geom(rng_t* r, double p) {
  if (p <= 0.0) return (int64_t)1 << 40;
  double u = rng_unif(r);
  if (u < 1e-300) u = 1e-300;
  return 1 + (int64_t)floor(log(u) / log1p(-p));
}

static const char CODE2ASCII[4] = {'A', 'T', 'G', 'C'};

static int cmp_i32(const void* a, const void* b) {
  int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
  return (x > y) - (x < y);
}

/* growable byte / word arrays */
typedef struct { uint8_t* p; uint64_t n, cap; } bytes_t;
static int bytes_reserve(bytes_t* b, uint64_t need) {
  if (need <= b->cap) return 0;
  uint64_t c = b->cap ? b->cap : 1024;
  while (c < need) c += c / 2 + 1024;
  uint8_t* q = (uint8_t*)realloc(b->p, c);
  if (!q) return -1;
  b->p = q; b->cap = c;
  return 0;
}

typedef struct { int32_t* pos; uint8_t *ref, *alt; uint64_t n, cap; } sites_t;
static int sites_push(sites_t* s, int32_t pos, uint8_t ref, uint8_t alt) {
  if (s->n == s->cap) {
    uint64_t c = s->cap ? s->cap * 2 : 4096;
    s->pos = (int32_t*)realloc(s->pos, c * sizeof(int32_t));
    s->ref = (uint8_t*)realloc(s->ref, c);
    s->alt = (uint8_t*)realloc(s->alt, c);
    if (!s->pos || !s->ref || !s->alt) return -1;
    s->cap = c;
  }
  s->pos[s->n] = pos; s->ref[s->n] = ref; s->alt[s->n] = alt; s->n++;
  return 0;
}

static inline uint8_t draw_bq(rng_t* r, double lowq_frac) {
  if (rng_unif(r) >= lowq_frac) return 93;
  return (uint8_t)(20 + rng_below(r, 73)); /* 20..92 */
}
static inline uint8_t other_base(rng_t* r, uint8_t c) { return (uint8_t)((c + 1 + rng_below(r, 3)) & 3); }

void hm_synth_default_spec(hm_synth_spec* s) {
  memset(s, 0, sizeof(*s));
  s->seed = 20260101ull;
  s->contig_len = 1000000;
  s->read_len_min = 5000;
  s->read_len_max = 25000;
  s->phase_block = 200000;
  s->depth = 30.0;
  s->read_len_mean = 15000.0;
  s->read_len_sd = 2000.0;
  s->het_rate = 1e-3;
  s->hom_rate = 5e-4;
  s->somatic_rate = 2e-6;
  s->sub_err_rate = 1e-4;
  s->indel_rate = 1e-4;
  s->lowq_frac = 0.1;
  s->mapq_low_frac = 0.03;
  s->softclip_frac = 0.01;
}

void hm_synth_free(hm_synth_data* d) {
  if (!d) return;
  free(d->ref); free(d->germ_pos); free(d->germ_ref); free(d->germ_alt); free(d->germ_gt);
  free(d->som_pos); free(d->som_ref); free(d->som_alt);
  free(d->err_pos); free(d->err_ref); free(d->err_alt);
  free(d->tstart); free(d->tend); free(d->qstart); free(d->qlen); free(d->mapq); free(d->flags);
  free(d->qname_id); free(d->seq_off); free(d->bq_off); free(d->op_off); free(d->n_ops);
  free(d->seq); free(d->bq); free(d->ops);
  free(d);
}

int hm_synth_generate(const hm_synth_spec* sp, hm_synth_data** out) {
  if (!sp || !out || sp->contig_len < 1000 || sp->read_len_min < 100 ||
      sp->read_len_max < sp->read_len_min)
    return HM_ERR_ARG;
  hm_synth_data* d = (hm_synth_data*)calloc(1, sizeof(hm_synth_data));
  if (!d) return HM_ERR_ARG;
  const int32_t L = sp->contig_len;
  rng_t rg;

  /* reference */
  uint8_t* refc = (uint8_t*)malloc((size_t)L); /* codes */
  d->ref = (uint8_t*)malloc((size_t)L);
  d->ref_len = (uint64_t)L;
  if (!refc || !d->ref) goto fail;
  rng_seed(&rg, sp->seed, 1);
  for (int32_t i = 0; i < L; i += 32) {
    uint64_t w = rng_next(&rg);
    for (int k = 0; k < 32 && i + k < L; k++) { refc[i + k] = (uint8_t)(w & 3); w >>= 2; }
  }
  for (int32_t i = 0; i < L; i++) d->ref[i] = (uint8_t)CODE2ASCII[refc[i]];

  /* germline SNVs */
  {
    rng_seed(&rg, sp->seed, 2);
    double p = sp->het_rate + sp->hom_rate;
    sites_t g = {0};
    bytes_t gt = {0};
    int64_t pos = -1;
    int32_t block = -1, flip = 0;
    for (;;) {
      pos += rng_geom(&rg, p);
      if (pos >= L) break;
      uint8_t rc = refc[pos], ac = other_base(&rg, rc);
      if (sites_push(&g, (int32_t)pos + 1, rc, ac)) goto fail;
      uint8_t t;
      if (rng_unif(&rg) * p < sp->hom_rate) t = 2;
      else {
        int32_t b = (int32_t)(pos / (sp->phase_block > 0 ? sp->phase_block : L));
        if (b != block) { block = b; flip = (int)(rng_next(&rg) & 1); }
        t = (uint8_t)((rng_next(&rg) & 1) ^ (uint64_t)flip) & 1;
      }
      if (bytes_reserve(&gt, gt.n + 1)) goto fail;
      gt.p[gt.n++] = t;
    }
    d->n_germ = g.n; d->germ_pos = g.pos; d->germ_ref = g.ref; d->germ_alt = g.alt; d->germ_gt = gt.p;
  }

  /* read placement */
  uint64_t n_reads = (uint64_t)(sp->depth * (double)L / sp->read_len_mean + 0.5);
  if (n_reads < 1) n_reads = 1;
  d->tstart = (int32_t*)malloc(n_reads * 4); d->tend = (int32_t*)malloc(n_reads * 4);
  d->qstart = (int32_t*)malloc(n_reads * 4); d->qlen = (int32_t*)malloc(n_reads * 4);
  d->mapq = (uint8_t*)malloc(n_reads); d->flags = (uint8_t*)calloc(n_reads, 1);
  d->qname_id = (uint32_t*)malloc(n_reads * 4);
  d->seq_off = (uint64_t*)malloc(n_reads * 8); d->bq_off = (uint64_t*)malloc(n_reads * 8);
  d->op_off = (uint64_t*)malloc(n_reads * 8); d->n_ops = (uint32_t*)malloc(n_reads * 4);
  if (!d->tstart || !d->tend || !d->qstart || !d->qlen || !d->mapq || !d->flags ||
      !d->qname_id || !d->seq_off || !d->bq_off || !d->op_off || !d->n_ops)
    goto fail;
  rng_seed(&rg, sp->seed, 3);
  {
    /* starts uniform over [-mean/2, L - min_len) clipped at 0 so the contig head is covered */
    int64_t lo = -(int64_t)(sp->read_len_mean / 2), hi = (int64_t)L - sp->read_len_min;
    if (hi <= 0) hi = 1;
    for (uint64_t i = 0; i < n_reads; i++) {
      int64_t s = lo + (int64_t)(rng_unif(&rg) * (double)(hi - lo));
      d->tstart[i] = (int32_t)(s < 0 ? 0 : s);
    }
    qsort(d->tstart, n_reads, sizeof(int32_t), cmp_i32);
  }

  bytes_t seq = {0}, bq = {0}, ops = {0};
  {
    uint64_t est = (uint64_t)((double)n_reads * (sp->read_len_mean * 1.03 + 64.0));
    if (bytes_reserve(&bq, est) || bytes_reserve(&seq, est / 4 + 64) ||
        bytes_reserve(&ops, n_reads * 64 * 4))
      goto fail;
  }
  sites_t som = {0}, err = {0};
  int32_t tmp_cap = sp->read_len_max + 4096;
  uint8_t* qc = (uint8_t*)malloc((size_t)tmp_cap);  /* query base codes */
  uint8_t* qq = (uint8_t*)malloc((size_t)tmp_cap);  /* query qualities  */
  uint32_t* rops = (uint32_t*)malloc((size_t)tmp_cap * 4);
  if (!qc || !qq || !rops) goto fail;
  const double p_rand = sp->somatic_rate + sp->sub_err_rate + sp->indel_rate;
  uint64_t g_lo = 0;

  for (uint64_t i = 0; i < n_reads; i++) {
    rng_t r;
    rng_seed(&r, sp->seed, 1000 + i);
    int32_t ts = d->tstart[i];
    double ln = sp->read_len_mean + sp->read_len_sd * rng_normal(&r);
    int32_t len = (int32_t)(ln + 0.5);
    if (len < sp->read_len_min) len = sp->read_len_min;
    if (len > sp->read_len_max) len = sp->read_len_max;
    int32_t te = ts + len;
    if (te > L) te = L;
    int hap = (int)(rng_next(&r) & 1);
    int32_t nq = 0, nops = 0, lead = 0, trail = 0;
    if (rng_unif(&r) < sp->softclip_frac) {
      if (rng_next(&r) & 1) lead = 50 + (int32_t)rng_below(&r, 451);
      if (!lead || (rng_next(&r) & 1)) trail = 50 + (int32_t)rng_below(&r, 451);
    }
    for (int32_t k = 0; k < lead; k++) { qc[nq] = (uint8_t)rng_below(&r, 4); qq[nq++] = draw_bq(&r, sp->lowq_frac); }

    while (g_lo < d->n_germ && d->germ_pos[g_lo] - 1 < ts) g_lo++;
    uint64_t g = g_lo;
    int32_t t = ts;          /* next reference position to emit */
    int32_t run = 0;         /* pending match run */
    int32_t min_ev = ts + 1; /* events must leave >= 1 matched base on each side */
    int64_t nr = (int64_t)ts + rng_geom(&r, p_rand);
    while (t < te) {
      /* next germline SNV this haplotype carries */
      int32_t pg = te;
      while (g < d->n_germ) {
        int32_t gp = d->germ_pos[g] - 1;
        if (gp >= te) break;
        if (gp >= t && (d->germ_gt[g] == 2 || d->germ_gt[g] == (uint8_t)hap)) { pg = gp; break; }
        g++;
      }
      int32_t pr = nr < te ? (int32_t)nr : te;
      int32_t pe = pg < pr ? pg : pr;
      if (pe >= te - 1) pe = te; /* keep the last base matched */
      /* matched bases up to the event */
      for (; t < pe; t++) { qc[nq] = refc[t]; qq[nq++] = draw_bq(&r, sp->lowq_frac); run++; }
      if (t >= te) break;
      if (pe == pg) { /* germline SNV (wins a collision with a random event) */
        if (run) { rops[nops++] = HM_MAKE_OP(HM_OP_MATCH, run); run = 0; }
        rops[nops++] = HM_MAKE_SUB(d->germ_ref[g], d->germ_alt[g]);
        qc[nq] = d->germ_alt[g]; qq[nq++] = draw_bq(&r, sp->lowq_frac);
        t++; g++;
        if (pr == pg) nr = (int64_t)t + rng_geom(&r, p_rand);
        if (min_ev < t) min_ev = t;
        continue;
      }
      /* random event at t == pr */
      nr = (int64_t)t + rng_geom(&r, p_rand);
      double u = rng_unif(&r) * p_rand;
      if (u < sp->somatic_rate + sp->sub_err_rate) {
        int is_som = u < sp->somatic_rate;
        uint8_t rc = refc[t], ac = other_base(&r, rc);
        if (run) { rops[nops++] = HM_MAKE_OP(HM_OP_MATCH, run); run = 0; }
        rops[nops++] = HM_MAKE_SUB(rc, ac);
        qc[nq] = ac;
        if (is_som) { qq[nq++] = 93; if (sites_push(&som, t + 1, rc, ac)) goto fail; }
        else {
          qq[nq++] = (rng_unif(&r) < 0.1) ? draw_bq(&r, sp->lowq_frac) : (uint8_t)(1 + rng_below(&r, 40));
          if (sites_push(&err, t + 1, rc, ac)) goto fail;
        }
        t++;
      } else {
        int32_t n = 1 + (int32_t)rng_below(&r, 3);
        int is_ins = (int)(rng_next(&r) & 1);
        if (t < min_ev || run == 0 || t + n + 1 >= te) { /* no room: emit as a match */
          qc[nq] = refc[t]; qq[nq++] = draw_bq(&r, sp->lowq_frac); run++; t++;
          continue;
        }
        rops[nops++] = HM_MAKE_OP(HM_OP_MATCH, run); run = 0;
        if (is_ins) {
          rops[nops++] = HM_MAKE_OP(HM_OP_INS, n);
          for (int32_t k = 0; k < n; k++) { qc[nq] = (uint8_t)rng_below(&r, 4); qq[nq++] = draw_bq(&r, sp->lowq_frac); }
        } else {
          rops[nops++] = HM_MAKE_OP(HM_OP_DEL, n);
          t += n;
        }
        /* force one matched base after an indel */
        qc[nq] = refc[t]; qq[nq++] = draw_bq(&r, sp->lowq_frac); run++; t++;
        min_ev = t;
        if (nr < t) nr = (int64_t)t + rng_geom(&r, p_rand);
      }
    }
    if (run) rops[nops++] = HM_MAKE_OP(HM_OP_MATCH, run);
    int32_t qend = nq;
    for (int32_t k = 0; k < trail; k++) { qc[nq] = (uint8_t)rng_below(&r, 4); qq[nq++] = draw_bq(&r, sp->lowq_frac); }

    d->tend[i] = te; d->qstart[i] = lead; d->qlen[i] = nq;
    d->mapq[i] = (rng_unif(&r) < sp->mapq_low_frac) ? (uint8_t)rng_below(&r, 60) : 60;
    d->qname_id[i] = (uint32_t)i;
    d->aligned_bases += (uint64_t)(qend - lead);
    /* append, 16-byte aligned */
    uint64_t so = (seq.n + 15) & ~15ull, bo = (bq.n + 15) & ~15ull;
    uint64_t sb = ((uint64_t)nq + 3) / 4;
    if (bytes_reserve(&seq, so + sb + 16) || bytes_reserve(&bq, bo + (uint64_t)nq + 16) ||
        bytes_reserve(&ops, ops.n + (uint64_t)nops * 4))
      goto fail;
    memset(seq.p + seq.n, 0, so - seq.n); memset(bq.p + bq.n, 0, bo - bq.n);
    memset(seq.p + so, 0, sb);
    for (int32_t k = 0; k < nq; k++) seq.p[so + (k >> 2)] |= (uint8_t)(qc[k] << (2 * (k & 3)));
    memcpy(bq.p + bo, qq, (size_t)nq);
    d->seq_off[i] = so; d->bq_off[i] = bo; seq.n = so + sb; bq.n = bo + (uint64_t)nq;
    d->op_off[i] = ops.n / 4; d->n_ops[i] = (uint32_t)nops;
    memcpy(ops.p + ops.n, rops, (size_t)nops * 4); ops.n += (uint64_t)nops * 4;
  }
  free(qc); free(qq); free(rops); free(refc); refc = NULL;
  {
    uint64_t sp16 = (seq.n + 15) & ~15ull, bp16 = (bq.n + 15) & ~15ull;
    if (bytes_reserve(&seq, sp16 + 16) || bytes_reserve(&bq, bp16 + 16)) goto fail;
    memset(seq.p + seq.n, 0, sp16 - seq.n); memset(bq.p + bq.n, 0, bp16 - bq.n);
    seq.n = sp16; bq.n = bp16;
  }
  d->seq = seq.p; d->bq = bq.p; d->ops = (uint32_t*)ops.p;
  d->n_som = som.n; d->som_pos = som.pos; d->som_ref = som.ref; d->som_alt = som.alt;
  d->n_err = err.n; d->err_pos = err.pos; d->err_ref = err.ref; d->err_alt = err.alt;

  d->batch.n_reads = n_reads;
  d->batch.tstart = d->tstart; d->batch.tend = d->tend; d->batch.qstart = d->qstart;
  d->batch.qlen = d->qlen; d->batch.mapq = d->mapq; d->batch.flags = d->flags;
  d->batch.qname_id = d->qname_id; d->batch.seq_off = d->seq_off; d->batch.bq_off = d->bq_off;
  d->batch.op_off = d->op_off; d->batch.n_ops = d->n_ops;
  d->batch.seq = d->seq; d->batch.seq_bytes = seq.n;
  d->batch.bq = d->bq; d->batch.bq_bytes = bq.n;
  d->batch.ops = d->ops; d->batch.n_ops_total = ops.n / 4;
  *out = d;
  return HM_OK;
fail:
  free(refc);
  hm_synth_free(d);
  return HM_ERR_ARG;
}
