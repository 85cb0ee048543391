/*
 * inflate_fast.h — raw DEFLATE (RFC 1951) decoder for BGZF blocks, host side.
 *
 * the BAM decode is bounded by the host inflate (zlib spends 75 % of a single-threaded
 * region decode there).  BGZF blocks are small (<= 64 KB inflated), independent and of known inflated size, so
 * this decoder drops what a streaming inflate pays for: no window management, no resumable state; a 64-bit bit
 * buffer refilled eight bytes at a time, two-level lookup tables (10 root bits for literal/length codes, 8 for
 * distances) whose entries carry base value and extra-bit count, word-wise match copies.  Every block is checked
 * against its CRC32 by the caller and falls back to zlib on any disagreement.
 *
 * hm_inflate_raw(in, in_len, out, out_len) -> 0 when exactly out_len bytes were produced and the final block
 * ended inside the input; anything else is an error (the caller then uses zlib).
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

static int hm_inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
  const uint8_t* const in_end = in + in_len;
  uint8_t* const out_begin = out;
  uint8_t* const out_end = out + out_len;
  uint64_t bitbuf = 0;
  int bitcnt = 0;
  hi_tables T;
  static const uint8_t PRE_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  for (;;) {
    HI_REFILL_SLOW();
    if (bitcnt < 3) return -1;
    const int final = (int)(bitbuf & 1), type = (int)((bitbuf >> 1) & 3);
    HI_TAKE(3);
    if (type == 0) { /* stored */
      const int drop = bitcnt & 7;
      HI_TAKE(drop);
      /* give whole bytes back to the input */
      in -= bitcnt >> 3; bitbuf = 0; bitcnt = 0;
      if (in_end - in < 4) return -1;
      const uint32_t len = in[0] | (in[1] << 8), nlen = in[2] | (in[3] << 8);
      in += 4;
      if ((len ^ nlen) != 0xffffu || (size_t)(in_end - in) < len || (size_t)(out_end - out) < len) return -1;
      memcpy(out, in, len);
      in += len; out += len;
    } else if (type == 1 || type == 2) {
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
        if (bitcnt < 14) return -1;
        nlit = 257 + (int)(bitbuf & 31); ndist = 1 + (int)((bitbuf >> 5) & 31);
        const int npre = 4 + (int)((bitbuf >> 10) & 15);
        HI_TAKE(14);
        if (nlit > 286 || ndist > 30) return -1;
        uint8_t plens[19];
        memset(plens, 0, sizeof(plens));
        for (int i = 0; i < npre; i++) {
          HI_REFILL_SLOW();
          if (bitcnt < 3) return -1;
          plens[PRE_ORDER[i]] = (uint8_t)(bitbuf & 7);
          HI_TAKE(3);
        }
        if (hi_build(T.pre, 1 << HI_PRE_ROOT, HI_PRE_ROOT, plens, 19, 0)) return -1;
        int i = 0;
        while (i < nlit + ndist) {
          HI_REFILL_SLOW();
          const uint32_t e = T.pre[bitbuf & ((1u << HI_PRE_ROOT) - 1)];
          const int nb = (int)(e & 15u);
          if (nb == 0 || nb > bitcnt) return -1;
          HI_TAKE(nb);
          const int sym = (int)(e >> 16);
          if (sym < 16) lens[i++] = (uint8_t)sym;
          else {
            int rep, val = 0;
            if (sym == 16) { if (i == 0 || bitcnt < 2) return -1; val = lens[i - 1]; rep = 3 + (int)(bitbuf & 3); HI_TAKE(2); }
            else if (sym == 17) { if (bitcnt < 3) return -1; rep = 3 + (int)(bitbuf & 7); HI_TAKE(3); }
            else { if (bitcnt < 7) return -1; rep = 11 + (int)(bitbuf & 127); HI_TAKE(7); }
            if (i + rep > nlit + ndist) return -1;
            memset(lens + i, val, (size_t)rep);
            i += rep;
          }
        }
        if (lens[256] == 0) return -1; /* no end-of-block code */
        /* the distance lengths follow the literal/length ones directly: move them to a fixed place */
        memmove(lens + 288, lens + nlit, (size_t)ndist);
        memset(lens + nlit, 0, (size_t)(288 - nlit));
        memset(lens + 288 + ndist, 0, (size_t)(32 - ndist));
        nlit = 288; ndist = 32;
      }
      if (hi_build(T.ll, HI_LL_SIZE, HI_LL_ROOT, lens, nlit, 1)) return -1;
      if (hi_build(T.d, HI_D_SIZE, HI_D_ROOT, lens + 288, ndist, 2)) return -1;
      /* ---- symbols ---- */
      const uint32_t ll_mask = (1u << HI_LL_ROOT) - 1, d_mask = (1u << HI_D_ROOT) - 1;
      int done = 0;
      /* fast loop: at least 16 input bytes and 280 output bytes ahead, so one word refill covers a whole
       * length / distance pair (15 + 5 + 15 + 13 bits) and match copies may overrun by up to 7 bytes */
      while (!done && in_end - in >= 16 && out_end - out >= 280) {
        HI_REFILL_FAST();
        uint32_t e = T.ll[bitbuf & ll_mask];
        if (((e >> 8) & 3u) == HI_SUB) { HI_TAKE(HI_LL_ROOT); e = T.ll[(e >> 16) + (bitbuf & ((1u << ((e >> 4) & 15u)) - 1u))]; }
        if ((e & 15u) == 0) return -1;
        HI_TAKE((int)(e & 15u));
        uint32_t ty = (e >> 8) & 3u;
        if (ty == HI_LIT) {
          *out++ = (uint8_t)(e >> 16);
          /* two more root-table literals fit in what the refill left (>= 41 bits) */
          e = T.ll[bitbuf & ll_mask];
          if ((e & 0x30fu) > 0 && ((e >> 8) & 3u) == HI_LIT) {
            HI_TAKE((int)(e & 15u)); *out++ = (uint8_t)(e >> 16);
            e = T.ll[bitbuf & ll_mask];
            if ((e & 0x30fu) > 0 && ((e >> 8) & 3u) == HI_LIT) { HI_TAKE((int)(e & 15u)); *out++ = (uint8_t)(e >> 16); }
          }
          continue;
        }
        if (ty == HI_EOB) { done = 1; break; }
        if (ty != HI_BASE) return -1;
        const int lx = (int)((e >> 4) & 15u);
        const uint32_t length = (e >> 16) + (uint32_t)(bitbuf & ((1u << lx) - 1u));
        HI_TAKE(lx);
        uint32_t ed = T.d[bitbuf & d_mask];
        if (((ed >> 8) & 3u) == HI_SUB) { HI_TAKE(HI_D_ROOT); ed = T.d[(ed >> 16) + (bitbuf & ((1u << ((ed >> 4) & 15u)) - 1u))]; }
        if ((ed & 15u) == 0 || ((ed >> 8) & 3u) != HI_BASE) return -1;
        HI_TAKE((int)(ed & 15u));
        const int dx = (int)((ed >> 4) & 15u);
        const uint32_t dist = (ed >> 16) + (uint32_t)(bitbuf & ((1u << dx) - 1u));
        HI_TAKE(dx);
        if (dist > (size_t)(out - out_begin)) return -1;
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
      }
      /* careful loop for the tail of the input / output */
      while (!done) {
        HI_REFILL_SLOW();
        uint32_t e = T.ll[bitbuf & ll_mask];
        if (((e >> 8) & 3u) == HI_SUB) {
          if (bitcnt < HI_LL_ROOT) return -1;
          HI_TAKE(HI_LL_ROOT);
          e = T.ll[(e >> 16) + (bitbuf & ((1u << ((e >> 4) & 15u)) - 1u))];
        }
        int nb = (int)(e & 15u);
        if (nb == 0 || nb > bitcnt) return -1;
        HI_TAKE(nb);
        const uint32_t ty = (e >> 8) & 3u;
        if (ty == HI_LIT) {
          if (out >= out_end) return -1;
          *out++ = (uint8_t)(e >> 16);
          continue;
        }
        if (ty == HI_EOB) break;
        if (ty != HI_BASE) return -1;
        const int lx = (int)((e >> 4) & 15u);
        if (lx > bitcnt) return -1;
        const uint32_t length = (e >> 16) + (uint32_t)(bitbuf & ((1u << lx) - 1u));
        HI_TAKE(lx);
        HI_REFILL_SLOW();
        uint32_t ed = T.d[bitbuf & d_mask];
        if (((ed >> 8) & 3u) == HI_SUB) {
          if (bitcnt < HI_D_ROOT) return -1;
          HI_TAKE(HI_D_ROOT);
          ed = T.d[(ed >> 16) + (bitbuf & ((1u << ((ed >> 4) & 15u)) - 1u))];
        }
        nb = (int)(ed & 15u);
        if (nb == 0 || nb > bitcnt || ((ed >> 8) & 3u) != HI_BASE) return -1;
        HI_TAKE(nb);
        const int dx = (int)((ed >> 4) & 15u);
        if (dx > bitcnt) return -1;
        const uint32_t dist = (ed >> 16) + (uint32_t)(bitbuf & ((1u << dx) - 1u));
        HI_TAKE(dx);
        if (dist > (size_t)(out - out_begin) || length > (size_t)(out_end - out)) return -1;
        const uint8_t* src = out - dist;
        for (uint32_t k = 0; k < length; k++) out[k] = src[k];
        out += length;
      }
    } else return -1;
    if (final) break;
  }
  return out == out_end ? 0 : -1;
}

#undef HI_REFILL_SLOW
#undef HI_REFILL_FAST
#undef HI_TAKE
#endif
