/*
 * crc32_clmul.h — CRC-32 (the gzip / BGZF polynomial, reflected 0xEDB88320) by carry-less multiplication.
 *
 * Every BGZF block the decoder inflates is checked against its CRC; zlib's table-driven crc32 runs at about
 * 2.8 GB/s and costs 8 % of a region decode.  Folding 64 bytes per step with PCLMULQDQ (the method of Gopal et al.,
 * "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ Instruction", Intel 2009; constants for this
 * polynomial in the bit-reflected domain: x^(4*128+32), x^(4*128-32), x^(128+32), x^(128-32), x^64 mod P and the
 * Barrett pair P', mu) is several times faster.  hm_crc32(buf, len) returns what zlib's crc32(0, buf, len) returns;
 * the caller decides at run time (hm_crc32_available) and keeps zlib for short inputs and other CPUs.
 * tests/test_inflate.py compares the two on random lengths and contents.
 */
#ifndef HM_CRC32_CLMUL_H
#define HM_CRC32_CLMUL_H
#include <stddef.h>
#include <stdint.h>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define HM_CRC32_CLMUL 1

static int hm_crc32_available(void) {
  static int known = -1;
  if (known < 0) known = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
  return known;
}

/* one 128-bit lane folded forward by the distance the constant pair `k` stands for, then combined with `next` */
__attribute__((target("pclmul,sse4.1"))) static inline __m128i hm_crc_fold(__m128i lane, __m128i k, __m128i next) {
  const __m128i lo = _mm_clmulepi64_si128(lane, k, 0x00);
  const __m128i hi = _mm_clmulepi64_si128(lane, k, 0x11);
  return _mm_xor_si128(_mm_xor_si128(lo, hi), next);
}

/* state in, state out (both in the register convention: the caller applies the ~ on either side); len >= 64 */
__attribute__((target("pclmul,sse4.1"))) static uint32_t hm_crc32_fold(uint32_t state, const uint8_t* p, size_t len, size_t* used) {
  const __m128i k_512 = _mm_set_epi64x(0x01c6e41596, 0x0154442bd4); /* fold by 4 lanes */
  const __m128i k_128 = _mm_set_epi64x(0x00ccaa009e, 0x01751997d0); /* fold by 1 lane  */
  const __m128i k_64 = _mm_set_epi64x(0, 0x0163cd6124);
  const __m128i barrett = _mm_set_epi64x(0x01f7011641, 0x01db710641);
  const __m128i low32 = _mm_setr_epi32(~0, 0, ~0, 0);
  const uint8_t* const start = p;
  __m128i a = _mm_xor_si128(_mm_loadu_si128((const __m128i*)p), _mm_cvtsi32_si128((int)state));
  __m128i b = _mm_loadu_si128((const __m128i*)(p + 16));
  __m128i c = _mm_loadu_si128((const __m128i*)(p + 32));
  __m128i d = _mm_loadu_si128((const __m128i*)(p + 48));
  p += 64; len -= 64;
  for (; len >= 64; p += 64, len -= 64) {
    a = hm_crc_fold(a, k_512, _mm_loadu_si128((const __m128i*)p));
    b = hm_crc_fold(b, k_512, _mm_loadu_si128((const __m128i*)(p + 16)));
    c = hm_crc_fold(c, k_512, _mm_loadu_si128((const __m128i*)(p + 32)));
    d = hm_crc_fold(d, k_512, _mm_loadu_si128((const __m128i*)(p + 48)));
  }
  /* four lanes -> one */
  a = hm_crc_fold(a, k_128, b);
  a = hm_crc_fold(a, k_128, c);
  a = hm_crc_fold(a, k_128, d);
  for (; len >= 16; p += 16, len -= 16) a = hm_crc_fold(a, k_128, _mm_loadu_si128((const __m128i*)p));
  /* 128 -> 64 bits */
  __m128i t = _mm_clmulepi64_si128(a, k_128, 0x10);
  a = _mm_xor_si128(_mm_srli_si128(a, 8), t);
  t = _mm_srli_si128(a, 4);
  a = _mm_xor_si128(_mm_clmulepi64_si128(_mm_and_si128(a, low32), k_64, 0x00), t);
  /* Barrett reduction 64 -> 32 bits */
  t = _mm_and_si128(_mm_clmulepi64_si128(_mm_and_si128(a, low32), barrett, 0x10), low32);
  a = _mm_xor_si128(a, _mm_clmulepi64_si128(t, barrett, 0x00));
  *used = (size_t)(p - start);
  return (uint32_t)_mm_extract_epi32(a, 1);
}
#else
#define HM_CRC32_CLMUL 0
static int hm_crc32_available(void) { return 0; }
#endif

#endif
