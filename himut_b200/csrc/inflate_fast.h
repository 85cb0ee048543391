/*
 * inflate_fast.h — raw DEFLATE (RFC 1951) decoder for BGZF blocks, host side.
 *
 * the BAM decode is bounded by the host inflate (zlib spends 75 % of a single-threaded
 * region decode there).  BGZF blocks are small (<= 64 KB inflated), independent and of known inflated size, so
 * this decoder drops what a streaming inflate pays for: no window management; a 64-bit bit buffer refilled eight
 * bytes at a time, two-level lookup tables (10 root bits for literal/length codes, 8 for distances) whose entries
 * carry base value and extra-bit count.  On x86-64 with BMI2 + SSSE3 the symbols are decoded by branch-free steps
 * (hi_step) and two blocks advance in one loop (hi_run2): see the comment above hi_step.  Every block is checked
 * against its CRC32 by the caller and falls back to zlib on any disagreement.
 *
 * hm_inflate_raw(in, in_len, out, out_len) -> 0 when exactly out_len bytes were produced and the final block
 * ended inside the input; anything else is an error (the caller then uses zlib).  hm_inflate_raw2 does the same
 * for two independent streams at once.
 */
#ifndef HM_INFLATE_FAST_H
#define HM_INFLATE_FAST_H
#include <stdint.h>
#include <string.h>

#define HI_LL_ROOT 10
#define HI_D_ROOT 8
#define HI_PRE_ROOT 7
#define HI_LL_SIZE 1408 /* >= 1334: 288 symbols, 10 root bits, 15-bit codes */
#define HI_D_SIZE 416   /* >= 402: 32 symbols, 8 root bits */

/* table entry: bits 0-3 code bits consumed at this level, 4-7 extra bits (or subtable index bits), 8-9 type,
 * 16-31 value (literal / base length / base distance / subtable offset); 0 = no such code */
enum { HI_LIT = 0, HI_BASE = 1, HI_EOB = 2, HI_SUB = 3 };
#define HI_ENTRY(value, type, extra, nbits) (((uint32_t)(value) << 16) | ((uint32_t)(type) << 8) | ((uint32_t)(extra) << 4) | (uint32_t)(nbits))

static const uint16_t HI_LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t HI_LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t HI_DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t HI_DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

static inline uint32_t hi_bitrev(uint32_t code, int len) {
  uint32_t r = 0;
  for (int i = 0; i < len; i++) { r = (r << 1) | (code & 1u); code >>= 1; }
  return r;
}

/* kind: 0 precode (value = symbol), 1 literal/length, 2 distance.  Returns 0 ok, -1 over-subscribed / bad symbol use. */
static int hi_build(uint32_t* tab, int tab_size, int root, const uint8_t* lens, int nsyms, int kind) {
  uint16_t count[16];
  uint16_t first[16];
  uint16_t sorted[288];
  uint8_t sub_extra[1 << HI_LL_ROOT];
  memset(count, 0, sizeof(count));
  for (int s = 0; s < nsyms; s++) count[lens[s]]++;
  count[0] = 0;
  /* over-subscription check */
  {
    int left = 1;
    for (int l = 1; l <= 15; l++) { left = (left << 1) - count[l]; if (left < 0) return -1; }
  }
  int off = 0;
  for (int l = 1; l <= 15; l++) { first[l] = (uint16_t)off; off += count[l]; }
  const int total = off;
  {
    uint16_t next[16];
    memcpy(next, first, sizeof(next));
    for (int s = 0; s < nsyms; s++) if (lens[s]) sorted[next[lens[s]]++] = (uint16_t)s;
  }
  const int root_size = 1 << root;
  memset(tab, 0, (size_t)root_size * sizeof(uint32_t));
  memset(sub_extra, 0, (size_t)root_size);
  /* pass 1: how deep is the subtable behind each root prefix */
  {
    uint32_t code = 0;
    int prev_len = 0;
    for (int i = 0; i < total; i++) {
      const int l = lens[sorted[i]];
      code <<= (l - prev_len); prev_len = l;
      if (l > root) {
        const uint32_t prefix = hi_bitrev(code >> (l - root), root);
        if (l - root > sub_extra[prefix]) sub_extra[prefix] = (uint8_t)(l - root);
      }
      code++;
    }
  }
  int next_free = root_size;
  {
    uint32_t code = 0;
    int prev_len = 0;
    for (int i = 0; i < total; i++) {
      const int sym = sorted[i];
      const int l = lens[sym];
      code <<= (l - prev_len); prev_len = l;
      uint32_t payload; /* entry without its nbits field */
      int usable = 1;
      if (kind == 0) payload = HI_ENTRY(sym, HI_LIT, 0, 0);
      else if (kind == 1) {
        if (sym < 256) payload = HI_ENTRY(sym, HI_LIT, 0, 0);
        else if (sym == 256) payload = HI_ENTRY(0, HI_EOB, 0, 0);
        else if (sym <= 285) payload = HI_ENTRY(HI_LEN_BASE[sym - 257], HI_BASE, HI_LEN_EXTRA[sym - 257], 0);
        else { payload = 0; usable = 0; } /* 286, 287: a code for them may exist, the data may not use it */
      } else {
        if (sym <= 29) payload = HI_ENTRY(HI_DIST_BASE[sym], HI_BASE, HI_DIST_EXTRA[sym], 0);
        else { payload = 0; usable = 0; }
      }
      const uint32_t rev = hi_bitrev(code, l);
      if (l <= root) {
        if (usable)
          for (uint32_t k = rev; k < (uint32_t)root_size; k += 1u << l) tab[k] = payload | (uint32_t)l;
      } else {
        const uint32_t prefix = rev & (uint32_t)(root_size - 1);
        const int sb = sub_extra[prefix];
        if ((tab[prefix] >> 8 & 3u) != HI_SUB || (tab[prefix] & 15u) == 0) {
          if (next_free + (1 << sb) > tab_size) return -1;
          memset(tab + next_free, 0, (size_t)(1 << sb) * sizeof(uint32_t));
          tab[prefix] = HI_ENTRY(next_free, HI_SUB, sb, root);
          next_free += 1 << sb;
        }
        const uint32_t base = tab[prefix] >> 16;
        if (usable)
          for (uint32_t k = rev >> root; k < (1u << sb); k += 1u << (l - root)) tab[base + k] = payload | (uint32_t)(l - root);
      }
      code++;
    }
  }
  return 0;
}

static inline uint64_t hi_load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

typedef struct {
  uint32_t ll[HI_LL_SIZE];
  uint32_t d[HI_D_SIZE];
  uint32_t pre[1 << HI_PRE_ROOT];
} hi_tables;

/* careful refill (byte-wise, never reads past in_end) and word-wise refill (needs 8 readable bytes at `in`) */
#define HI_REFILL_SLOW() do { while (bitcnt <= 56 && in < in_end) { bitbuf |= (uint64_t)(*in++) << bitcnt; bitcnt += 8; } } while (0)
#define HI_REFILL_FAST() do { bitbuf |= hi_load64(in) << bitcnt; in += (63 - bitcnt) >> 3; bitcnt |= 56; } while (0)
#define HI_TAKE(n) do { bitbuf >>= (n); bitcnt -= (n); } while (0)

/* ---- branch-free symbol runs (x86-64 with BMI2 + SSSE3) ---------------------------------------------------
 * BAM payloads decode to about as many short matches (7 bytes on average) as literals, in no predictable order: a
 * decoder that branches on "literal or match", on the distance class and on the copy length keeps recovering from
 * mispredicted branches.  hi_run decodes symbols in a straight line instead: the distance code is looked up
 * speculatively and masked away for a literal; the output is always one 16-byte store of the source bytes spread by
 * a byte shuffle whose mask repeats a pattern of `dist` bytes (identity for dist >= 16) — a literal's source is a
 * row of 16 copies of itself — so overlapping copies need no byte loop.  The run stops in front of anything rare
 * (codes longer than the root tables, end of block) and the general step of hm_inflate_raw takes that symbol.
 * Lean entries for the run, derived from the two-level tables after each build:
 *   bits 0-5 bits consumed (code + extra), 8-13 code length, 15 rare (leave the run), 16-30 base value, 31 match */
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define HI_RUN 1
#define HI_RARE 0x8000u
typedef struct {
  uint32_t ll[1 << HI_LL_ROOT];
  uint32_t d[1 << HI_D_ROOT];
} hi_lean;
static const uint8_t HI_SPREAD[17][16] __attribute__((aligned(16))) = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1},       {0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 1, 2, 0},
    {0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3},       {0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 0},
    {0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3},       {0, 1, 2, 3, 4, 5, 6, 0, 1, 2, 3, 4, 5, 6, 0, 1},
    {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3, 4, 5, 6, 7},       {0, 1, 2, 3, 4, 5, 6, 7, 8, 0, 1, 2, 3, 4, 5, 6},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 1, 2, 3, 4, 5},       {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 0, 1, 2, 3, 4},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 0, 1, 2, 3},     {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 0, 1, 2},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 0, 1},   {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}};
/* bytes a spread pattern advances per 16-byte store: the largest multiple of dist that fits (16 for dist >= 16) */
static const uint8_t HI_STRIDE[17] = {16, 16, 16, 15, 16, 15, 12, 14, 16, 9, 10, 11, 12, 13, 14, 15, 16};
static uint8_t HI_LITROW[256][16] __attribute__((aligned(16)));
static int hi_run_ok = -1;
static int hi_have_run(void) {
  if (hi_run_ok < 0) {
    for (int v = 0; v < 256; v++) memset(HI_LITROW[v], v, 16);
    hi_run_ok = __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("ssse3");
  }
  return hi_run_ok;
}
static void hi_make_lean(const uint32_t* ll, const uint32_t* d, hi_lean* L) {
  for (int i = 0; i < (1 << HI_LL_ROOT); i++) {
    const uint32_t e = ll[i], nb = e & 15u, ex = (e >> 4) & 15u, ty = (e >> 8) & 3u;
    L->ll[i] = (nb == 0 || ty >= HI_EOB) ? HI_RARE : ((nb + ex) | (nb << 8) | ((e >> 16) << 16) | (ty == HI_BASE ? 0x80000000u : 0u));
  }
  for (int i = 0; i < (1 << HI_D_ROOT); i++) {
    const uint32_t e = d[i], nb = e & 15u, ex = (e >> 4) & 15u, ty = (e >> 8) & 3u;
    L->d[i] = (nb == 0 || ty != HI_BASE) ? HI_RARE : ((nb + ex) | (nb << 8) | ((e >> 16) << 16));
  }
}
/* one symbol; 0 = done, 1 = something rare is next (nothing consumed), -1 = a match reaching in front of the output.
 * Needs 16 readable input bytes and 320 writable output bytes. */
__attribute__((target("bmi2,ssse3"), always_inline)) static inline int hi_step(const hi_lean* L, const uint8_t** in_p, uint8_t** out_p,
                                                                             const uint8_t* out_begin, uint64_t* bitbuf_p, int* bitcnt_p) {
  const uint8_t* in = *in_p;
  uint8_t* out = *out_p;
  uint64_t bitbuf = *bitbuf_p;
  int bitcnt = *bitcnt_p;
  HI_REFILL_FAST();
  *in_p = in; *bitbuf_p = bitbuf; *bitcnt_p = bitcnt;                     /* a refill never loses anything */
  uint64_t bb = bitbuf;
  const uint32_t e = L->ll[bb & ((1u << HI_LL_ROOT) - 1)];
  if (__builtin_expect(e & HI_RARE, 0)) return 1;
  const uint32_t mm = (uint32_t)((int32_t)e >> 31);                       /* all ones for a match */
  const uint32_t val = ((e >> 16) & 0x7fffu) + (uint32_t)(_bzhi_u64(bb, e & 0xffu) >> ((e >> 8) & 63u));
  bb >>= (e & 63u);
  const uint32_t ed = L->d[bb & ((1u << HI_D_ROOT) - 1)];                 /* speculative: only a match consumes it */
  if (__builtin_expect(ed & mm & HI_RARE, 0)) return 1;
  const uint32_t edm = ed & mm;
  const uint32_t dist = (edm >> 16) + (uint32_t)(_bzhi_u64(bb, edm & 0xffu) >> ((edm >> 8) & 63u));
  bb >>= (edm & 63u);
  if (__builtin_expect(dist > (size_t)(out - out_begin), 0)) return -1;
  *bitbuf_p = bb; *bitcnt_p = bitcnt - (int)((e & 63u) + (edm & 63u));
  const uint32_t len = 1u + ((val - 1u) & mm);
  /* selects written as masks: the compiler turns ?: on these into the very branches this step exists to avoid */
  const uint32_t near = (uint32_t)((int32_t)(dist - 16u) >> 31) & mm;     /* match with dist 1..15: spread a pattern */
  const uint32_t cls = (dist & near) | (16u & ~near);                     /* else (far match, literal): plain bytes */
  const uint64_t mm64 = (uint64_t)(int64_t)(int32_t)mm;
  const uint8_t* src = (const uint8_t*)(((uintptr_t)(out - dist) & mm64) | ((uintptr_t)HI_LITROW[val & 255u] & ~mm64));
  const __m128i v = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)src), _mm_load_si128((const __m128i*)HI_SPREAD[cls]));
  _mm_storeu_si128((__m128i*)out, v);
  if (__builtin_expect(len > 16u, 0)) {
    if (cls == 16u) {
      for (uint32_t k = 16; k < len; k += 16) _mm_storeu_si128((__m128i*)(out + k), _mm_loadu_si128((const __m128i*)(src + k)));
    } else {
      const uint32_t stride = HI_STRIDE[cls];
      for (uint32_t k = stride; k < len; k += stride) _mm_storeu_si128((__m128i*)(out + k), v);
    }
  }
  *out_p = out + len;
  return 0;
}
#else
#define HI_RUN 0
#endif

static inline int in_end_ok(const uint8_t* in, const uint8_t* in_end) { return in_end - in >= 16; }
static inline int out_end_ok(const uint8_t* out, const uint8_t* out_end) { return out_end - out >= 320; }

/* ---- one raw DEFLATE stream, decoded piecewise so that two of them can share a loop ---- */
enum { HI_P_HEADER = 0, HI_P_SYMBOLS = 1, HI_P_DONE = 2, HI_P_ERROR = 3 };
typedef struct {
  const uint8_t* in;
  const uint8_t* in_end;
  uint8_t* out;
  uint8_t* out_begin;
  uint8_t* out_end;
  uint64_t bitbuf;
  int bitcnt;
  int final; /* the block in progress is the last one */
  int phase;
  hi_tables T;
#if HI_RUN
  hi_lean lean;
#endif
} hi_stream;

static void hi_init(hi_stream* s, const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
  s->in = in; s->in_end = in + in_len; s->out = out; s->out_begin = out; s->out_end = out + out_len;
  s->bitbuf = 0; s->bitcnt = 0; s->final = 0; s->phase = HI_P_HEADER;
}
#define HI_FAST_OK(s) ((s)->in_end - (s)->in >= 16 && (s)->out_end - (s)->out >= 320)
#define HI_OPEN(s)                                                                                            \
  const uint8_t* in = (s)->in; const uint8_t* const in_end = (s)->in_end; uint8_t* out = (s)->out;           \
  uint8_t* const out_begin = (s)->out_begin; uint8_t* const out_end = (s)->out_end; uint64_t bitbuf = (s)->bitbuf; \
  int bitcnt = (s)->bitcnt; (void)out_begin; (void)out_end; (void)in_end
#define HI_CLOSE(s) do { (s)->in = in; (s)->out = out; (s)->bitbuf = bitbuf; (s)->bitcnt = bitcnt; } while (0)
#define HI_FAIL(s) do { (s)->phase = HI_P_ERROR; return; } while (0)
#define HI_END_BLOCK(s) ((s)->phase = (s)->final ? HI_P_DONE : HI_P_HEADER)

/* block header: a stored block is copied here, for a compressed one the tables are built */
static void hi_header(hi_stream* s, int run) {
  static const uint8_t PRE_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  HI_OPEN(s);
  HI_REFILL_SLOW();
  if (bitcnt < 3) HI_FAIL(s);
  s->final = (int)(bitbuf & 1);
  const int type = (int)((bitbuf >> 1) & 3);
  HI_TAKE(3);
  if (type == 0) { /* stored */
    const int drop = bitcnt & 7;
    HI_TAKE(drop);
    /* give whole bytes back to the input */
    in -= bitcnt >> 3; bitbuf = 0; bitcnt = 0;
    if (in_end - in < 4) HI_FAIL(s);
    const uint32_t len = in[0] | (in[1] << 8), nlen = in[2] | (in[3] << 8);
    in += 4;
    if ((len ^ nlen) != 0xffffu || (size_t)(in_end - in) < len || (size_t)(out_end - out) < len) HI_FAIL(s);
    memcpy(out, in, len);
    in += len; out += len;
    HI_CLOSE(s);
    HI_END_BLOCK(s);
    return;
  }
  if (type == 3) HI_FAIL(s);
  uint8_t lens[288 + 32];
  int nlit, ndist;
  if (type == 1) {
    nlit = 288; ndist = 32;
    for (int i = 0; i < 144; i++) lens[i] = 8;
    for (int i = 144; i < 256; i++) lens[i] = 9;
    for (int i = 256; i < 280; i++) lens[i] = 7;
    for (int i = 280; i < 288; i++) lens[i] = 8;
    for (int i = 0; i < 32; i++) lens[288 + i] = 5;
  } else {
    HI_REFILL_SLOW();
    if (bitcnt < 14) HI_FAIL(s);
    nlit = 257 + (int)(bitbuf & 31); ndist = 1 + (int)((bitbuf >> 5) & 31);
    const int npre = 4 + (int)((bitbuf >> 10) & 15);
    HI_TAKE(14);
    if (nlit > 286 || ndist > 30) HI_FAIL(s);
    uint8_t plens[19];
    memset(plens, 0, sizeof(plens));
    for (int i = 0; i < npre; i++) {
      HI_REFILL_SLOW();
      if (bitcnt < 3) HI_FAIL(s);
      plens[PRE_ORDER[i]] = (uint8_t)(bitbuf & 7);
      HI_TAKE(3);
    }
    if (hi_build(s->T.pre, 1 << HI_PRE_ROOT, HI_PRE_ROOT, plens, 19, 0)) HI_FAIL(s);
    int i = 0;
    while (i < nlit + ndist) {
      HI_REFILL_SLOW();
      const uint32_t e = s->T.pre[bitbuf & ((1u << HI_PRE_ROOT) - 1)];
      const int nb = (int)(e & 15u);
      if (nb == 0 || nb > bitcnt) HI_FAIL(s);
      HI_TAKE(nb);
      const int sym = (int)(e >> 16);
      if (sym < 16) lens[i++] = (uint8_t)sym;
      else {
        int rep, val = 0;
        if (sym == 16) { if (i == 0 || bitcnt < 2) HI_FAIL(s); val = lens[i - 1]; rep = 3 + (int)(bitbuf & 3); HI_TAKE(2); }
        else if (sym == 17) { if (bitcnt < 3) HI_FAIL(s); rep = 3 + (int)(bitbuf & 7); HI_TAKE(3); }
        else { if (bitcnt < 7) HI_FAIL(s); rep = 11 + (int)(bitbuf & 127); HI_TAKE(7); }
        if (i + rep > nlit + ndist) HI_FAIL(s);
        memset(lens + i, val, (size_t)rep);
        i += rep;
      }
    }
    if (lens[256] == 0) HI_FAIL(s); /* no end-of-block code */
    /* the distance lengths follow the literal/length ones directly: move them to a fixed place */
    memmove(lens + 288, lens + nlit, (size_t)ndist);
    memset(lens + nlit, 0, (size_t)(288 - nlit));
    memset(lens + 288 + ndist, 0, (size_t)(32 - ndist));
    nlit = 288; ndist = 32;
  }
  if (hi_build(s->T.ll, HI_LL_SIZE, HI_LL_ROOT, lens, nlit, 1)) HI_FAIL(s);
  if (hi_build(s->T.d, HI_D_SIZE, HI_D_ROOT, lens + 288, ndist, 2)) HI_FAIL(s);
#if HI_RUN
  if (run) hi_make_lean(s->T.ll, s->T.d, &s->lean);
#else
  (void)run;
#endif
  HI_CLOSE(s);
  s->phase = HI_P_SYMBOLS;
}

/* one symbol by the general route (root or second-level table); after a literal, up to two more root-table literals.
 * Needs HI_FAST_OK: one word refill covers a whole length / distance pair (15 + 5 + 15 + 13 bits) and a match copy may
 * overrun by up to 7 bytes. */
static void hi_general_step(hi_stream* s) {
  HI_OPEN(s);
  const uint32_t ll_mask = (1u << HI_LL_ROOT) - 1, d_mask = (1u << HI_D_ROOT) - 1;
  const uint32_t* const ll = s->T.ll;
  const uint32_t* const dt = s->T.d;
  HI_REFILL_FAST();
  uint32_t e = ll[bitbuf & ll_mask];
  if (((e >> 8) & 3u) == HI_SUB) { HI_TAKE(HI_LL_ROOT); e = ll[(e >> 16) + (bitbuf & ((1u << ((e >> 4) & 15u)) - 1u))]; }
  if ((e & 15u) == 0) HI_FAIL(s);
  HI_TAKE((int)(e & 15u));
  const uint32_t ty = (e >> 8) & 3u;
  if (ty == HI_LIT) {
    *out++ = (uint8_t)(e >> 16);
    /* two more root-table literals fit in what the refill left (>= 41 bits) */
    e = ll[bitbuf & ll_mask];
    if ((e & 0x30fu) > 0 && ((e >> 8) & 3u) == HI_LIT) {
      HI_TAKE((int)(e & 15u)); *out++ = (uint8_t)(e >> 16);
      e = ll[bitbuf & ll_mask];
      if ((e & 0x30fu) > 0 && ((e >> 8) & 3u) == HI_LIT) { HI_TAKE((int)(e & 15u)); *out++ = (uint8_t)(e >> 16); }
    }
    HI_CLOSE(s);
    return;
  }
  if (ty == HI_EOB) { HI_CLOSE(s); HI_END_BLOCK(s); return; }
  if (ty != HI_BASE) HI_FAIL(s);
  const int lx = (int)((e >> 4) & 15u);
  const uint32_t length = (e >> 16) + (uint32_t)(bitbuf & ((1u << lx) - 1u));
  HI_TAKE(lx);
  uint32_t ed = dt[bitbuf & d_mask];
  if (((ed >> 8) & 3u) == HI_SUB) { HI_TAKE(HI_D_ROOT); ed = dt[(ed >> 16) + (bitbuf & ((1u << ((ed >> 4) & 15u)) - 1u))]; }
  if ((ed & 15u) == 0 || ((ed >> 8) & 3u) != HI_BASE) HI_FAIL(s);
  HI_TAKE((int)(ed & 15u));
  const int dx = (int)((ed >> 4) & 15u);
  const uint32_t dist = (ed >> 16) + (uint32_t)(bitbuf & ((1u << dx) - 1u));
  HI_TAKE(dx);
  if (dist > (size_t)(out - out_begin)) HI_FAIL(s);
  const uint8_t* src = out - dist;
  uint8_t* dst = out;
  out += length;
  if (dist >= 8) {
    do { memcpy(dst, src, 8); dst += 8; src += 8; } while (dst < out);
  } else if (dist == 1) {
    memset(dst, *src, length);
  } else {
    do { *dst++ = *src++; } while (dst < out);
  }
  HI_CLOSE(s);
}

/* the rest of the block, byte-careful: for the tail of the input / output */
static void hi_tail(hi_stream* s) {
  HI_OPEN(s);
  const uint32_t ll_mask = (1u << HI_LL_ROOT) - 1, d_mask = (1u << HI_D_ROOT) - 1;
  const uint32_t* const ll = s->T.ll;
  const uint32_t* const dt = s->T.d;
  for (;;) {
    HI_REFILL_SLOW();
    uint32_t e = ll[bitbuf & ll_mask];
    if (((e >> 8) & 3u) == HI_SUB) {
      if (bitcnt < HI_LL_ROOT) HI_FAIL(s);
      HI_TAKE(HI_LL_ROOT);
      e = ll[(e >> 16) + (bitbuf & ((1u << ((e >> 4) & 15u)) - 1u))];
    }
    int nb = (int)(e & 15u);
    if (nb == 0 || nb > bitcnt) HI_FAIL(s);
    HI_TAKE(nb);
    const uint32_t ty = (e >> 8) & 3u;
    if (ty == HI_LIT) {
      if (out >= out_end) HI_FAIL(s);
      *out++ = (uint8_t)(e >> 16);
      continue;
    }
    if (ty == HI_EOB) break;
    if (ty != HI_BASE) HI_FAIL(s);
    const int lx = (int)((e >> 4) & 15u);
    if (lx > bitcnt) HI_FAIL(s);
    const uint32_t length = (e >> 16) + (uint32_t)(bitbuf & ((1u << lx) - 1u));
    HI_TAKE(lx);
    HI_REFILL_SLOW();
    uint32_t ed = dt[bitbuf & d_mask];
    if (((ed >> 8) & 3u) == HI_SUB) {
      if (bitcnt < HI_D_ROOT) HI_FAIL(s);
      HI_TAKE(HI_D_ROOT);
      ed = dt[(ed >> 16) + (bitbuf & ((1u << ((ed >> 4) & 15u)) - 1u))];
    }
    nb = (int)(ed & 15u);
    if (nb == 0 || nb > bitcnt || ((ed >> 8) & 3u) != HI_BASE) HI_FAIL(s);
    HI_TAKE(nb);
    const int dx = (int)((ed >> 4) & 15u);
    if (dx > bitcnt) HI_FAIL(s);
    const uint32_t dist = (ed >> 16) + (uint32_t)(bitbuf & ((1u << dx) - 1u));
    HI_TAKE(dx);
    if (dist > (size_t)(out - out_begin) || length > (size_t)(out_end - out)) HI_FAIL(s);
    const uint8_t* src = out - dist;
    for (uint32_t k = 0; k < length; k++) out[k] = src[k];
    out += length;
  }
  HI_CLOSE(s);
  HI_END_BLOCK(s);
}

#if HI_RUN
/* straight-line symbols of one stream; a rare symbol (long code, end of block) goes through hi_general_step and the
 * run goes on while the block lasts and the fast region holds */
__attribute__((target("bmi2,ssse3"))) static void hi_run(hi_stream* s) {
  const uint8_t* in = s->in;
  uint8_t* out = s->out;
  uint64_t bitbuf = s->bitbuf;
  int bitcnt = s->bitcnt;
  while (in_end_ok(in, s->in_end) && out_end_ok(out, s->out_end)) {
    const int rc = hi_step(&s->lean, &in, &out, s->out_begin, &bitbuf, &bitcnt);
    if (__builtin_expect(rc != 0, 0)) {
      s->in = in; s->out = out; s->bitbuf = bitbuf; s->bitcnt = bitcnt;
      if (rc < 0) { s->phase = HI_P_ERROR; return; }
      hi_general_step(s);
      if (s->phase != HI_P_SYMBOLS) return;
      in = s->in; out = s->out; bitbuf = s->bitbuf; bitcnt = s->bitcnt;
    }
  }
  s->in = in; s->out = out; s->bitbuf = bitbuf; s->bitcnt = bitcnt;
}
/* the same for two streams in one loop: their dependency chains (table lookup -> shift -> table lookup) are
 * independent, so the core overlaps them.  Returns when either stream leaves its block or its fast region. */
__attribute__((target("bmi2,ssse3"))) static void hi_run2(hi_stream* a, hi_stream* b) {
  const uint8_t* in_a = a->in; uint8_t* out_a = a->out; uint64_t bitbuf_a = a->bitbuf; int bitcnt_a = a->bitcnt;
  const uint8_t* in_b = b->in; uint8_t* out_b = b->out; uint64_t bitbuf_b = b->bitbuf; int bitcnt_b = b->bitcnt;
  const uint8_t* const in_end_a = a->in_end; const uint8_t* const out_end_a = a->out_end; const uint8_t* const out_begin_a = a->out_begin;
  const uint8_t* const in_end_b = b->in_end; const uint8_t* const out_end_b = b->out_end; const uint8_t* const out_begin_b = b->out_begin;
  const hi_lean* const la = &a->lean;
  const hi_lean* const lb = &b->lean;
  while (in_end_ok(in_a, in_end_a) && out_end_ok(out_a, out_end_a) && in_end_ok(in_b, in_end_b) && out_end_ok(out_b, out_end_b)) {
    const int rc_a = hi_step(la, &in_a, &out_a, out_begin_a, &bitbuf_a, &bitcnt_a);
    const int rc_b = hi_step(lb, &in_b, &out_b, out_begin_b, &bitbuf_b, &bitcnt_b);
    if (__builtin_expect((rc_a | rc_b) != 0, 0)) {
      a->in = in_a; a->out = out_a; a->bitbuf = bitbuf_a; a->bitcnt = bitcnt_a;
      b->in = in_b; b->out = out_b; b->bitbuf = bitbuf_b; b->bitcnt = bitcnt_b;
      if (rc_a < 0) a->phase = HI_P_ERROR;
      if (rc_b < 0) b->phase = HI_P_ERROR;
      if (rc_a > 0) hi_general_step(a);
      if (rc_b > 0) hi_general_step(b);
      if (a->phase != HI_P_SYMBOLS || b->phase != HI_P_SYMBOLS) return;
      in_a = a->in; out_a = a->out; bitbuf_a = a->bitbuf; bitcnt_a = a->bitcnt;
      in_b = b->in; out_b = b->out; bitbuf_b = b->bitbuf; bitcnt_b = b->bitcnt;
    }
  }
  a->in = in_a; a->out = out_a; a->bitbuf = bitbuf_a; a->bitcnt = bitcnt_a;
  b->in = in_b; b->out = out_b; b->bitbuf = bitbuf_b; b->bitcnt = bitcnt_b;
}
#endif

/* some progress on one stream: a block header, or a straight run plus the symbol that stopped it, or the tail */
static void hi_advance(hi_stream* s, int run) {
  if (s->phase > HI_P_SYMBOLS) return;
  if (s->phase == HI_P_HEADER) { hi_header(s, run); return; }
  if (!HI_FAST_OK(s)) { hi_tail(s); return; }
#if HI_RUN
  if (run) { hi_run(s); return; } /* comes back at the end of the block or of the fast region */
#endif
  hi_general_step(s);
}
static int hi_result(const hi_stream* s) { return s->phase == HI_P_DONE && s->out == s->out_end ? 0 : -1; }

static int hi_use_run(void) {
#if HI_RUN
  return hi_have_run();
#else
  return 0;
#endif
}

static int hm_inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
  hi_stream s;
  const int run = hi_use_run();
  hi_init(&s, in, in_len, out, out_len);
  while (s.phase <= HI_P_SYMBOLS) hi_advance(&s, run);
  return hi_result(&s);
}

/* two independent streams (two BGZF blocks) at once; rc[i] as hm_inflate_raw */
static void hm_inflate_raw2(const uint8_t* in0, size_t in_len0, uint8_t* out0, size_t out_len0, const uint8_t* in1, size_t in_len1,
                            uint8_t* out1, size_t out_len1, int rc[2]) {
#if HI_RUN
  if (hi_have_run()) {
    hi_stream a, b;
    hi_init(&a, in0, in_len0, out0, out_len0);
    hi_init(&b, in1, in_len1, out1, out_len1);
    while (a.phase <= HI_P_SYMBOLS && b.phase <= HI_P_SYMBOLS) {
      if (a.phase == HI_P_SYMBOLS && b.phase == HI_P_SYMBOLS && HI_FAST_OK(&a) && HI_FAST_OK(&b)) {
        hi_run2(&a, &b);
      } else {
        hi_advance(&a, 1);
        hi_advance(&b, 1);
      }
    }
    while (a.phase <= HI_P_SYMBOLS) hi_advance(&a, 1);
    while (b.phase <= HI_P_SYMBOLS) hi_advance(&b, 1);
    rc[0] = hi_result(&a); rc[1] = hi_result(&b);
    return;
  }
#endif
  rc[0] = hm_inflate_raw(in0, in_len0, out0, out_len0);
  rc[1] = hm_inflate_raw(in1, in_len1, out1, out_len1);
}

#undef HI_OPEN
#undef HI_CLOSE
#undef HI_FAIL
#undef HI_END_BLOCK
#undef HI_REFILL_SLOW
#undef HI_REFILL_FAST
#undef HI_TAKE
#endif
