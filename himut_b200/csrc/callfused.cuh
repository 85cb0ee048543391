// callfused.cuh — the fused device path of `himut call` (sm_100a).
//
// The reference piles up every aligned base of every read (caller.update_allelecounts, caller.py:44-72) and then looks
// at the pileup only where a read carries a candidate substitution (caller.py:324-336).  The first version of this
// path (kernels.cuh: k_read_scan -> k_candidates -> library sort -> k_site_entries_by_read -> k_site_reduce) touched
// the quality stream twice: once to sum it for the QV gate, and once more, one 32-byte sector per looked-up byte, to
// gather the pileup columns.  Here the quality stream is read exactly once:
//
//   k_call_pairs   (ops only)  one warp per (chunk, read) pair: cs op prefix scan, mismatch list, identity / MAPQ /
//                  length gates, [--phase] read haplotype, then trim + mismatch-window test of every substitution
//                  (bamlib.get_tsbs_candidates, bamlib.py:69-86) -> candidate keys into the chunk's own segment.
//                  The QV gate needs every quality byte, so candidates are emitted *speculatively* with their read.
//   k_site_sort    one CTA per 2^17-position tile of a chunk: the chunk's keys -> sorted distinct sites without a
//                  comparison sort: position bitmap in shared memory, popcount prefix = rank, 16-bit (ref, alt) mask
//                  per distinct position (set(somatic_tsbs_candidate_lst), caller.py:324).
//   k_tile_scan    exclusive scan of the per-tile site counts (one CTA); k_site_range2: final key array, the
//                  file-order read range of every site, entry slots initialised.
//   k_call_scan    the one pass over the quality stream.  One warp per (chunk, read) pair; the first pair of a read
//                  streams the read's qualities through a 3-stage cp.async.bulk (TMA) ring in shared memory, sums
//                  them (QV gate, bamlib.get_qv) and, while a block is in shared memory, answers the pileup lookups
//                  of the sites its read covers from it.  No second touch of the stream, no sector gather.
//   k_site_valid   only if some read failed the QV gate: sites keep only candidates of reads that passed.
//   k_site_reduce  (kernels.cuh) per site, reads in file order: counts, ordered fp64 sums, PL / GQ, cascade, record.
//   k_compact_sites  records the host wants, stably compacted (no germline restatements when asked, no dropped sites).
//
// No host synchronisation between the kernels: capacities are bounds known on the host (ops per chunk) or a
// high-water estimate with an overflow flag (distinct sites), counts stay on the device.
#pragma once
#include "kernels.cuh"
#include "normcounts.cuh" // mbarrier / cp.async.bulk helpers

#define HC_MAX_OPS 192                 // ops of a read staged per warp; longer lists use the batch's global op_t / op_q
#define HC_TILE_BITS 17
#define HC_TILE (1u << HC_TILE_BITS)   // positions per sort tile
#define HC_TILE_WORDS (HC_TILE / 32u)
#define HC_CAPD 8192u                  // distinct positions of a tile kept in shared memory (more: global scratch)
#define HC_SORT_THREADS 512
#define HC_BLK 2048u                   // bytes per quality stage
#define HC_NS 3                        // stages per warp
#define HC_WARPS 8                     // warps per CTA of k_call_pairs / k_call_scan
#define HC_KEY_POS_BITS 28             // a chunk may span < 2^28 positions on this path (else the first version runs)

struct OpView { const uint32_t* w; const uint32_t* t; const uint32_t* q; uint32_t n; };

// allele of a read at reference offset `off` from its start: -1 none, 5 deleted, 8 the base of a match run, else the
// substituted base; *q = query position of the base, *ins = insertions whose reference position is `off`
// (same rules as read_allele_at, kernels.cuh)
__device__ __forceinline__ int op_allele(const OpView& v, uint32_t off, uint32_t* q, int* ins) {
  *q = 0; *ins = 0;
  if (v.n == 0) return -1;
  uint32_t lo = 0, hi = v.n;
  while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (v.t[m] <= off) lo = m + 1; else hi = m; }
  const int k = (int)lo - 1;
  if (k < 0) return -1;
  int cnt = 0;
  for (int j = k; j >= 0 && v.t[j] == off; j--) cnt += ((v.w[j] & 3u) == HM_OP_INS);
  *ins = cnt;
  const uint32_t wd = v.w[k], kind = wd & 3u, val = wd >> 2, t0 = v.t[k];
  const uint32_t rl = (uint32_t)op_ref_len(wd);
  if (rl == 0 || off >= t0 + rl) return -1;
  if (kind == HM_OP_DEL) return 5;
  *q = v.q[k] + (kind == HM_OP_MATCH ? off - t0 : 0u);
  return kind == HM_OP_SUB ? (int)((val >> 3) & 7u) : 8;
}

// warp-cooperative: the chunk c with off[c] <= x < off[c + 1] (off: n + 1 ascending entries)
__device__ __forceinline__ uint32_t warp_find_chunk(const uint64_t* off, uint32_t n, uint64_t x, int lane) {
  uint32_t lo = 0, len = n + 1;
  while (len > 32) {
    const uint32_t step = (len + 32) / 33, end = lo + len;
    const uint32_t idx = lo + step * (uint32_t)(lane + 1) - 1;
    const bool below = idx < end && __ldg(off + idx) <= x;
    const uint32_t c = __popc(__ballot_sync(HM_FULL, below));
    lo += c * step;
    len = min(step, end - lo);
  }
  const bool below = (uint32_t)lane < len && __ldg(off + lo + lane) <= x;
  return lo + __popc(__ballot_sync(HM_FULL, below)) - 1;
}

// haplib.get_ccs_hap (haplib.py:61-83) for a warp over staged ops: lanes = hetSNPs
template <bool kSeq>
__device__ __forceinline__ int warp_read_hap_ops(const DevBatch& b, uint32_t r, int32_t ts, int32_t te, const OpView& v,
                                                 const DevPhase& ph, int set, int lane) {
  if (set < 0 || (uint32_t)set >= ph.n_sets) return 2;
  const uint64_t s0 = ph.set_off[set];
  const uint32_t n = (uint32_t)(ph.set_off[set + 1] - s0);
  const int32_t* hp = ph.hpos + s0;
  const uint32_t idx = upper_bound_dev(hp, n, ts);
  const uint32_t jdx = upper_bound_dev(hp, n, te);
  if (jdx - idx < 2) return 2;
  bool h0 = true, h1 = true;
  for (uint32_t k = idx + lane; k < jdx; k += 32) {
    uint32_t q; int ins;
    int a = op_allele(v, (uint32_t)(__ldg(hp + k) - 1 - ts), &q, &ins);
    const int hr = (int)ph.href[s0 + k];
    if (a == 8) a = kSeq ? (int)((b.seq[__ldg(b.seq_off + r) + (q >> 2)] >> (2 * (q & 3u))) & 3u) : href_match_base(hr);
    int bit = 2;
    if (a >= 0 && a < 4) {
      if (a == hr) bit = 0;
      else if (a == (int)ph.halt[s0 + k]) bit = 1;
    }
    const int hb = ph.hbit[s0 + k];
    if (bit != hb) h0 = false;
    if (bit != 1 - hb) h1 = false;
  }
  h0 = __all_sync(HM_FULL, h0);
  h1 = __all_sync(HM_FULL, h1);
  return h0 ? 0 : (h1 ? 1 : 2);
}

// ============================================================================ k_call_pairs
// pair_hap (--phase only): 0 / 1 / 2 ("."), 3 = the chunk does not fetch the read.
// read_counted[r] = 1: some chunk fetched r, r passed the MAPQ / identity / length gates there and (--phase) got a
//   haplotype: if its QV gate passes too (k_call_scan) it counts in num_ccs and its candidates are real.
// first_pair[r]: the lowest pair index that fetches r; that pair's warp streams the read's qualities in k_call_scan.
template <bool kSeq>
__global__ void __launch_bounds__(32 * HC_WARPS) k_call_pairs(DevBatch b, DevParams p, DevPhase ph, const hm_chunk* chunks, uint32_t n_chunks,
                                                              const uint64_t* pair_off, uint64_t n_pairs, uint8_t* pair_hap,
                                                              uint8_t* read_counted, uint32_t* first_pair, const uint64_t* seg_off,
                                                              uint32_t* seg_cnt, uint32_t* seg_keys, uint32_t* seg_read) {
  __shared__ uint32_t s_w[HC_WARPS][HC_MAX_OPS], s_t[HC_WARPS][HC_MAX_OPS], s_q[HC_WARPS][HC_MAX_OPS];
  __shared__ int32_t s_mm[HC_WARPS][HC_MAX_OPS];
  const uint64_t pr = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (pr >= n_pairs) return;
  const uint32_t c = warp_find_chunk(pair_off, n_chunks, pr, lane);
  const hm_chunk ch = chunks[c];
  const uint32_t r = ch.read_lo + (uint32_t)(pr - __ldg(pair_off + c));
  if (p.phase && lane == 0) pair_hap[pr] = 3;
  if (__ldg(b.flags + r) & HM_READ_SECONDARY) return;
  const int32_t ts = __ldg(b.tstart + r), te = __ldg(b.tend + r);
  if (!(ts < ch.end && te > ch.start)) return;
  if (lane == 0) atomicMin(first_pair + r, (uint32_t)pr);
  const uint32_t n = __ldg(b.n_ops + r);
  const uint64_t o0 = __ldg(b.op_off + r);
  const int32_t qlen = __ldg(b.qlen + r);
  const bool staged = n <= HC_MAX_OPS;
  uint32_t* Tw = staged ? s_t[wid] : b.op_t + o0;
  uint32_t* Qw = staged ? s_q[wid] : b.op_q + o0;
  int32_t* Mw = staged ? s_mm[wid] : b.mm_pos + o0;
  const uint32_t* Wr = staged ? s_w[wid] : b.ops + o0;

  // sweep 1: prefix scan of (reference, query) lengths, mismatch list, identity tallies (cslib.cs2subindel,
  // bamlib.get_blast_sequence_identity); pairs of one read write the same values where the arrays are global
  uint32_t t_carry = 0, q_carry = (uint32_t)__ldg(b.qstart + r);
  int mm_base = 0, nm = 0, ns = 0, il = 0, dl = 0;
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t k = base + lane;
    const bool valid = k < n;
    const uint32_t w = valid ? __ldg(b.ops + o0 + k) : 0u;
    const uint32_t kind = w & 3u, v = w >> 2;
    const uint32_t rl = (uint32_t)op_ref_len(w), al = (uint32_t)op_qry_len(w);
    uint32_t rs = rl, qs = al;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t a = __shfl_up_sync(HM_FULL, rs, d), e = __shfl_up_sync(HM_FULL, qs, d);
      if (lane >= d) { rs += a; qs += e; }
    }
    const uint32_t t_ex = t_carry + rs - rl, q_ex = q_carry + qs - al;
    if (valid) { if (staged) s_w[wid][k] = w; Tw[k] = t_ex; Qw[k] = q_ex; }
    const bool is_mm = valid && ((kind == HM_OP_SUB && (v & 7u) != HM_BASE_N) || kind == HM_OP_INS || kind == HM_OP_DEL);
    const uint32_t bal = __ballot_sync(HM_FULL, is_mm);
    if (is_mm) Mw[mm_base + __popc(bal & ((1u << lane) - 1u))] = ts + (int32_t)t_ex + 1;
    mm_base += __popc(bal);
    if (valid) {
      if (kind == HM_OP_MATCH) nm += (int)v;
      else if (kind == HM_OP_SUB) ns += 1;
      else if (kind == HM_OP_INS) il += (int)v;
      else dl += (int)v;
    }
    t_carry += __shfl_sync(HM_FULL, rs, 31);
    q_carry += __shfl_sync(HM_FULL, qs, 31);
  }
  __syncwarp();
  nm = __reduce_add_sync(HM_FULL, nm);
  ns = __reduce_add_sync(HM_FULL, ns);
  il = __reduce_add_sync(HM_FULL, il);
  dl = __reduce_add_sync(HM_FULL, dl);
  const int nmm = mm_base;

  // read gates of caller.py:310-317 except the QV gate (k_call_scan has the quality sum), same order of evaluation
  bool pre_ok = true;
  if ((int)__ldg(b.mapq + r) < p.min_mapq) pre_ok = false;
  const double ident = __ddiv_rn((double)nm, (double)(nm + ns + il + dl));
  if (ident < p.min_sequence_identity) pre_ok = false;
  if (!(p.qlen_lower_limit < qlen && qlen < p.qlen_upper_limit)) pre_ok = false;

  const OpView view = {Wr, Tw, Qw, n};
  if (p.phase) {
    const int hap = warp_read_hap_ops<kSeq>(b, r, ts, te, view, ph, ch.phase_set, lane);
    if (lane == 0) pair_hap[pr] = (uint8_t)hap;
    if (hap > 1) return;
  }
  if (!pre_ok) return;
  if (lane == 0) read_counted[r] = 1;

  // sweep 2: bamlib.get_tsbs_candidates (bamlib.py:69-86) for every substitution of the read that lies in the chunk
  const double trim_s = floor(__dmul_rn(p.min_trim, (double)qlen));
  const double trim_e = ceil(__dmul_rn(__dsub_rn(1.0, p.min_trim), (double)qlen));
  const int wsz = p.mismatch_window;
  const uint64_t seg0 = __ldg(seg_off + c);
  uint32_t mm_seen = 0;
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t k = base + lane;
    uint32_t op = 0, v = 0;
    bool is_mm = false;
    if (k < n) {
      op = Wr[k];
      v = op >> 2;
      const uint32_t kind = op & 3u;
      is_mm = (kind == HM_OP_SUB && (v & 7u) != HM_BASE_N) || kind == HM_OP_INS || kind == HM_OP_DEL;
    }
    const uint32_t mbal = __ballot_sync(HM_FULL, is_mm);
    const int rank = (int)(mm_seen + __popc(mbal & ((1u << lane) - 1u))); // this op's own entry in the mismatch list
    mm_seen += __popc(mbal);
    bool emit = false;
    uint32_t key = 0;
    if (k < n && (op & 3u) == HM_OP_SUB && (v & 7u) != HM_BASE_N) {
      const int32_t tpos = ts + (int32_t)Tw[k] + 1;
      const int32_t qpos = (int32_t)Qw[k];
      if (ch.start <= tpos && tpos <= ch.end && !((double)qpos < trim_s) && !((double)qpos > trim_e)) {
        const int qs = qpos - wsz, qe = qpos + wsz; // bamlib.get_mismatch_range
        int u, d;
        if (qs < 0) { u = wsz + qs; d = wsz + (-qs); }
        else if (qe > qlen) { u = wsz + (qe - qlen); d = qlen - qpos; }
        else { u = wsz; d = wsz; }
        int cnt = 0; // mismatches in [tpos - u, tpos + d] other than this one (bamlib.py:266-282)
        for (int j = rank + 1; j < nmm && Mw[j] <= tpos + d; j++) cnt++;
        for (int j = rank - 1; j >= 0 && Mw[j] >= tpos - u; j--) cnt++;
        if (!(cnt > p.max_mismatch_count)) {
          emit = true;
          key = ((uint32_t)(tpos - ch.start) << 4) | ((v & 3u) << 2) | ((v >> 3) & 3u);
        }
      }
    }
    const uint32_t bal = __ballot_sync(HM_FULL, emit);
    if (bal) {
      uint32_t at = 0;
      if (lane == 0) at = atomicAdd(seg_cnt + c, (uint32_t)__popc(bal));
      at = __shfl_sync(HM_FULL, at, 0);
      if (emit) {
        const uint64_t slot = seg0 + at + __popc(bal & ((1u << lane) - 1u));
        seg_keys[slot] = key;
        seg_read[slot] = r;
      }
    }
  }
}

// ============================================================================ k_site_sort
// block-wide exclusive scan of one value per thread (HC_SORT_THREADS threads); returns the exclusive prefix, *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(HM_FULL, incl, d); if (lane >= d) incl += t; }
  __syncthreads(); // s_warp may still be read from a previous scan
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0u, wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(HM_FULL, wi, d); if (lane >= d) wi += t; }
    s_warp[lane] = wi - w;        // exclusive prefix of the warp totals
    if (lane == 31) s_warp[32] = wi; // grand total
  }
  __syncthreads();
  *total = s_warp[32];
  return s_warp[wid] + incl - v;
}

// One CTA per tile.  tile_chunk[t] = chunk of tile t, tile_off[c] = first tile of chunk c (host).  Outputs: the
// tile's sorted distinct keys (chunk-relative: rel << 4 | ref << 2 | alt) at keys_tmp[tile_src[t] ...), tile_cnt[t],
// and for every emitted key its rank among the tile's distinct keys (key_site).
__global__ void __launch_bounds__(HC_SORT_THREADS) k_site_sort(const uint32_t* tile_chunk, const uint32_t* tile_off, const uint64_t* seg_off,
                                                               const uint32_t* seg_cnt, const uint32_t* seg_keys, uint32_t* chunk_cursor,
                                                               uint32_t* chunk_cursor2, uint32_t* gscratch, uint32_t* keys_tmp,
                                                               uint32_t* key_site, uint32_t* tile_src, uint32_t* tile_cnt) {
  extern __shared__ uint32_t sm[];
  uint32_t* s_bits = sm;                          // HC_TILE_WORDS
  uint32_t* s_wpre = s_bits + HC_TILE_WORDS;      // HC_TILE_WORDS: distinct positions before each word
  uint32_t* s_mask = s_wpre + HC_TILE_WORDS;      // HC_CAPD: (ref, alt) combinations seen at the rank-th distinct position
  uint32_t* s_off = s_mask + HC_CAPD;             // HC_CAPD: distinct keys before the rank-th distinct position
  __shared__ uint32_t s_warp[33];
  __shared__ uint32_t s_base[2];
  const uint32_t t = blockIdx.x, tid = threadIdx.x;
  const uint32_t c = tile_chunk[t];
  const uint32_t tile_in_chunk = t - tile_off[c];
  const uint64_t seg0 = seg_off[c];
  const uint32_t nk = seg_cnt[c];
  const uint32_t* keys = seg_keys + seg0;

  for (uint32_t i = tid; i < HC_TILE_WORDS; i += blockDim.x) s_bits[i] = 0u;
  __syncthreads();
  for (uint32_t j = tid; j < nk; j += blockDim.x) {
    const uint32_t rel = __ldg(keys + j) >> 4;
    if ((rel >> HC_TILE_BITS) == tile_in_chunk) atomicOr(&s_bits[(rel & (HC_TILE - 1u)) >> 5], 1u << (rel & 31u));
  }
  __syncthreads();
  // distinct positions before each bitmap word: each thread owns HC_TILE_WORDS / blockDim.x consecutive words
  constexpr uint32_t WPT = HC_TILE_WORDS / HC_SORT_THREADS;
  uint32_t local = 0;
#pragma unroll
  for (uint32_t i = 0; i < WPT; i++) local += __popc(s_bits[tid * WPT + i]);
  uint32_t n_dist;
  uint32_t run = block_excl_scan(local, s_warp, &n_dist);
#pragma unroll
  for (uint32_t i = 0; i < WPT; i++) { s_wpre[tid * WPT + i] = run; run += __popc(s_bits[tid * WPT + i]); }
  // (ref, alt) masks and offsets per distinct position: shared memory, or this tile's own piece of the global scratch
  const bool in_smem = n_dist <= HC_CAPD;
  if (!in_smem && tid == 0) s_base[0] = atomicAdd(chunk_cursor2 + c, 2u * n_dist); // n_dist <= keys of the tile
  __syncthreads();
  uint32_t* mask = in_smem ? s_mask : gscratch + 2 * seg0 + s_base[0];
  uint32_t* off = in_smem ? s_off : mask + n_dist;
  for (uint32_t i = tid; i < n_dist; i += blockDim.x) mask[i] = 0u;
  __syncthreads();
  for (uint32_t j = tid; j < nk; j += blockDim.x) {
    const uint32_t key = __ldg(keys + j), rel = key >> 4;
    if ((rel >> HC_TILE_BITS) != tile_in_chunk) continue;
    const uint32_t p = rel & (HC_TILE - 1u);
    const uint32_t rank = s_wpre[p >> 5] + __popc(s_bits[p >> 5] & ((1u << (p & 31u)) - 1u));
    atomicOr(&mask[rank], 1u << (key & 15u));
  }
  __syncthreads();
  // distinct keys before each distinct position: each thread owns a consecutive run of ranks
  const uint32_t per = (n_dist + blockDim.x - 1) / blockDim.x;
  const uint32_t r0 = min(tid * per, n_dist), r1 = min(r0 + per, n_dist);
  local = 0;
  for (uint32_t i = r0; i < r1; i++) local += __popc(mask[i]);
  uint32_t n_uniq;
  run = block_excl_scan(local, s_warp, &n_uniq);
  for (uint32_t i = r0; i < r1; i++) { off[i] = run; run += __popc(mask[i]); }
  if (tid == 0) {
    const uint32_t at = atomicAdd(chunk_cursor + c, n_uniq); // tiles of a chunk share its segment of keys_tmp
    s_base[1] = at;
    tile_src[t] = (uint32_t)seg0 + at;
    tile_cnt[t] = n_uniq;
  }
  __syncthreads();
  uint32_t* dst = keys_tmp + seg0 + s_base[1];
  // sorted distinct keys: positions ascending (bitmap order), then (ref, alt) ascending (mask bit order)
#pragma unroll 1
  for (uint32_t i = 0; i < WPT; i++) {
    const uint32_t wi = tid * WPT + i;
    uint32_t bits = s_bits[wi];
    uint32_t rank = s_wpre[wi];
    while (bits) {
      const uint32_t bit = __ffs(bits) - 1;
      bits &= bits - 1;
      uint32_t m = mask[rank], o = off[rank];
      const uint32_t rel = (tile_in_chunk << HC_TILE_BITS) | (wi << 5) | bit;
      while (m) { const uint32_t cb = __ffs(m) - 1; m &= m - 1; dst[o++] = (rel << 4) | cb; }
      rank++;
    }
  }
  // where each emitted key went (k_site_valid needs it when a read fails the QV gate)
  for (uint32_t j = tid; j < nk; j += blockDim.x) {
    const uint32_t key = __ldg(keys + j), rel = key >> 4;
    if ((rel >> HC_TILE_BITS) != tile_in_chunk) continue;
    const uint32_t p = rel & (HC_TILE - 1u);
    const uint32_t rank = s_wpre[p >> 5] + __popc(s_bits[p >> 5] & ((1u << (p & 31u)) - 1u));
    key_site[seg0 + j] = off[rank] + __popc(mask[rank] & ((1u << (key & 15u)) - 1u));
  }
}

// exclusive scan of tile_cnt -> tile_dst[0 .. n_tiles]; cnt[1] = number of distinct sites, or 0 with cnt[5] = the
// number needed when it exceeds site_cap (the host grows its buffers and runs the call again)
__global__ void __launch_bounds__(1024) k_tile_scan(const uint32_t* tile_cnt, uint32_t n_tiles, uint32_t* tile_dst, unsigned long long site_cap,
                                                    unsigned long long* cnt) {
  __shared__ uint32_t s_warp[33];
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n_tiles; base += blockDim.x) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n_tiles ? tile_cnt[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_excl_scan(v, s_warp, &total);
    if (i < n_tiles) tile_dst[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) {
    const bool over = (unsigned long long)carry > site_cap;
    tile_dst[n_tiles] = over ? 0u : carry;
    cnt[1] = over ? 0ull : (unsigned long long)carry;
    cnt[5] = over ? (unsigned long long)carry : 0ull;
  }
  if (n_tiles == 0 && threadIdx.x == 0) tile_dst[0] = 0u;
}

// final key array (chunk << 36 | tpos << 4 | ref << 2 | alt, sorted), the file-order read range of every site
// (as k_site_range), its entry slots set to "unwritten", its validity flag cleared
__global__ void __launch_bounds__(256) k_site_range2(DevBatch b, const hm_chunk* chunks, const uint32_t* tile_chunk, const uint32_t* tile_src,
                                                     const uint32_t* tile_dst, uint32_t n_tiles, const uint32_t* keys_tmp,
                                                     const unsigned long long* n_keys_dev, unsigned long long* keys, uint32_t* site_lo,
                                                     uint32_t* site_n, uint32_t* entries, uint64_t stride, uint8_t* site_valid) {
  const uint64_t ki = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ki >= *n_keys_dev) return;
  const uint32_t t = upper_bound_dev(tile_dst, n_tiles + 1, (uint32_t)ki) - 1; // last tile with tile_dst <= ki (empty tiles repeat a value)
  const uint32_t c = __ldg(tile_chunk + t);
  const hm_chunk ch = chunks[c];
  const uint32_t k32 = __ldg(keys_tmp + __ldg(tile_src + t) + ((uint32_t)ki - __ldg(tile_dst + t)));
  const int32_t tpos = ch.start + (int32_t)(k32 >> 4);
  keys[ki] = ((unsigned long long)c << 36) | ((unsigned long long)(uint32_t)tpos << 4) | (unsigned long long)(k32 & 15u);
  const int32_t rpos = tpos - 1;
  const uint32_t n_in = ch.read_hi - ch.read_lo;
  const uint32_t lo = count_le_kary_i32(b.pmax_tend + ch.read_lo, n_in, rpos - 1);
  const uint32_t hi = count_le_kary_i32(b.tstart + ch.read_lo, n_in, rpos);
  const uint32_t n = hi > lo ? hi - lo : 0u;
  site_lo[ki] = ch.read_lo + lo;
  site_n[ki] = n;
  site_valid[ki] = 0;
  const uint32_t ns = min(n, (uint32_t)HM_SITE_SLOTS);
  for (uint32_t s = 0; s < ns; s++) entries[(uint64_t)s * stride + ki] = HM_ENT_UNWRITTEN;
}

// ============================================================================ k_call_scan
// One warp per (chunk, read) pair.  Shared memory per warp: the read's ops with their prefixes, HC_NS quality stages,
// HC_NS mbarriers.  The pair that owns the read (first_pair) and whose read can still count (read_counted) streams the
// read's qualities: lane 0 keeps HC_NS bulk copies in flight; every block is summed from shared memory and then serves
// the sites whose query position lies in it.  Other pairs fetch their few quality bytes directly.
struct __align__(16) ScanWarp {
  uint8_t bq[HC_NS][HC_BLK];
  uint32_t w[HC_MAX_OPS], t[HC_MAX_OPS], q[HC_MAX_OPS];
  uint64_t bar[HC_NS];
  uint64_t pad;
};

template <bool kSeq>
__global__ void __launch_bounds__(32 * HC_WARPS) k_call_scan(DevBatch b, DevParams p, const hm_chunk* chunks, uint32_t n_chunks,
                                                             const uint64_t* pair_off, uint64_t n_pairs, const uint8_t* pair_hap,
                                                             const uint8_t* read_counted, const uint32_t* first_pair,
                                                             const uint32_t* tile_off, const uint32_t* tile_dst,
                                                             const unsigned long long* keys, const uint32_t* site_lo, const uint32_t* site_n,
                                                             uint32_t* entries, uint64_t stride, uint8_t* qv_fail_read,
                                                             unsigned int* qv_fail_any, uint8_t* qname_seen,
                                                             const unsigned long long* n_keys_dev) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint64_t pr = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (pr >= n_pairs) return;
  ScanWarp* S = reinterpret_cast<ScanWarp*>(smem_raw) + wid;
  const uint32_t c = warp_find_chunk(pair_off, n_chunks, pr, lane);
  const hm_chunk ch = chunks[c];
  const uint32_t r = ch.read_lo + (uint32_t)(pr - __ldg(pair_off + c));
  if (__ldg(b.flags + r) & HM_READ_SECONDARY) return;
  const int32_t ts = __ldg(b.tstart + r), te = __ldg(b.tend + r);
  if (!(ts < ch.end && te > ch.start)) return;
  const bool owner = __ldg(first_pair + r) == (uint32_t)pr && __ldg(read_counted + r) != 0;
  const uint32_t n = __ldg(b.n_ops + r);
  // the chunk's sites the read can touch: rpos in [ts, te], i.e. tpos in [ts + 1, te + 1]
  uint32_t s_lo = 0, s_hi = 0;
  if (n && *n_keys_dev) { // 0 also when the site list overflowed its buffers (k_tile_scan): the host runs the call again
    const uint32_t k_lo = __ldg(tile_dst + __ldg(tile_off + c)), k_hi = __ldg(tile_dst + __ldg(tile_off + c + 1));
    if (k_lo < k_hi) {
      const unsigned long long kb = (unsigned long long)c << 36;
      s_lo = warp_lower_bound_u64(keys, k_lo, k_hi, kb | ((unsigned long long)(uint32_t)(ts + 1) << 4), lane);
      s_hi = warp_lower_bound_u64(keys, s_lo, k_hi, kb | ((unsigned long long)(uint32_t)(te + 2) << 4), lane);
    }
  }
  if (!owner && s_lo >= s_hi) return;
  const uint64_t o0 = __ldg(b.op_off + r);
  const bool staged = n <= HC_MAX_OPS;
  if (staged && s_lo < s_hi) { // the prefix scan again (registers only): cheaper than keeping 8 B per op in HBM
    uint32_t t_carry = 0, q_carry = (uint32_t)__ldg(b.qstart + r);
    for (uint32_t base = 0; base < n; base += 32) {
      const uint32_t k = base + lane;
      const uint32_t w = k < n ? __ldg(b.ops + o0 + k) : 0u;
      const uint32_t rl = (uint32_t)op_ref_len(w), al = (uint32_t)op_qry_len(w);
      uint32_t rs = rl, qs = al;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(HM_FULL, rs, d), e = __shfl_up_sync(HM_FULL, qs, d);
        if (lane >= d) { rs += a; qs += e; }
      }
      if (k < n) { S->w[k] = w; S->t[k] = t_carry + rs - rl; S->q[k] = q_carry + qs - al; }
      t_carry += __shfl_sync(HM_FULL, rs, 31);
      q_carry += __shfl_sync(HM_FULL, qs, 31);
    }
    __syncwarp();
  }
  const OpView view = {staged ? S->w : b.ops + o0, staged ? S->t : b.op_t + o0, staged ? S->q : b.op_q + o0, n};
  const uint64_t bq0 = __ldg(b.bq_off + r);
  const uint8_t* bqg = b.bq + bq0;
  const uint64_t sq0 = kSeq ? __ldg(b.seq_off + r) : 0;
  const uint32_t hap = p.phase ? (uint32_t)pair_hap[pr] : 2u;
  const uint32_t qlen = (uint32_t)__ldg(b.qlen + r);

  // site groups: 32 consecutive sites, one per lane.  pending: the lane's site waits for the quality byte at my_q.
  uint32_t g_next = s_lo;
  bool pending = false;
  uint32_t my_q = 0, my_e = 0;
  uint64_t my_at = 0;
  auto load_group = [&]() {
    const uint32_t ki = g_next + (uint32_t)lane;
    g_next += 32;
    if (ki >= s_hi) return;
    const uint32_t slot = r - __ldg(site_lo + ki);
    if (slot >= HM_SITE_SLOTS || slot >= __ldg(site_n + ki)) return; // deep pileups: k_site_reduce computes these itself
    const unsigned long long key = __ldg(keys + ki);
    const int32_t rpos = (int32_t)((key >> 4) & 0xffffffffull) - 1;
    uint32_t q; int ins;
    int a = op_allele(view, (uint32_t)(rpos - ts), &q, &ins);
    const bool has_base = a >= 0 && a != 5;
    if (a == 8) a = kSeq ? (int)((b.seq[sq0 + (q >> 2)] >> (2 * (q & 3u))) & 3u) : (int)((key >> 2) & 3);
    const uint32_t e = (a < 0 ? HM_ENT_NONE : (uint32_t)a) | ((uint32_t)min(ins, 255) << 11) | ((hap & 3u) << 19) |
                       ((te > rpos + 1) ? (1u << 21) : 0u);
    const uint64_t at = (uint64_t)slot * stride + ki;
    if (has_base) { pending = true; my_q = q; my_e = e; my_at = at; }
    else entries[at] = e;
  };

  if (owner) {
    const uint32_t nbytes = (qlen + 15u) & ~15u; // the stream is padded to 16 bytes per read
    const uint32_t nblk = (nbytes + HC_BLK - 1u) / HC_BLK;
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < HC_NS; s++) mbar_init(&S->bar[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
      for (uint32_t s = 0; s < (uint32_t)HC_NS && s < nblk; s++) {
        const uint32_t bytes = min(HC_BLK, nbytes - s * HC_BLK);
        mbar_arrive_expect_tx(&S->bar[s], bytes);
        bulk_g2s(S->bq[s], bqg + (size_t)s * HC_BLK, bytes, &S->bar[s]);
      }
    }
    uint32_t acc = 0;
    for (uint32_t blk = 0; blk < nblk; blk++) {
      const uint32_t st = blk % HC_NS, parity = (blk / HC_NS) & 1u;
      mbar_wait(&S->bar[st], parity);
      const uint32_t blk0 = blk * HC_BLK;
      const uint32_t bytes = min(HC_BLK, nbytes - blk0);
      const uint4* s4 = reinterpret_cast<const uint4*>(S->bq[st]);
      for (uint32_t i = lane; i < (bytes >> 4); i += 32) {
        uint4 v = s4[i];
        const uint32_t g0 = blk0 + (i << 4); // bytes past the read's length are padding: not part of the sum
        if (g0 + 16u > qlen) {
          const uint32_t keep = qlen - g0; // 1 .. 15
          uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int kb = (int)keep - 4 * j;
            wv[j] = kb >= 4 ? wv[j] : kb <= 0 ? 0u : (wv[j] & ((1u << (8 * kb)) - 1u));
          }
          v = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
        acc = sum4(v.x, acc); acc = sum4(v.y, acc); acc = sum4(v.z, acc); acc = sum4(v.w, acc);
      }
      // sites whose quality byte is in this block (query positions ascend with the sites)
      for (;;) {
        if (!__any_sync(HM_FULL, pending)) {
          if (g_next >= s_hi) break;
          load_group();
          continue;
        }
        if (pending && my_q >= blk0 && my_q < blk0 + bytes) {
          entries[my_at] = my_e | ((uint32_t)S->bq[st][my_q - blk0] << 3);
          pending = false;
        }
        if (__any_sync(HM_FULL, pending)) break; // the rest waits for a later block
      }
      __syncwarp(); // every lane is done with this stage
      if (lane == 0 && blk + HC_NS < nblk) {
        const uint32_t nb0 = (blk + HC_NS) * HC_BLK;
        const uint32_t nbts = min(HC_BLK, nbytes - nb0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(&S->bar[st], nbts);
        bulk_g2s(S->bq[st], bqg + nb0, nbts, &S->bar[st]);
      }
    }
    // the QV gate (bamlib.get_qv, caller.py:310): np.mean over every base of the read
    unsigned long long tot = acc;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(HM_FULL, tot, d);
    if (lane == 0) {
      const double qv = __ddiv_rn((double)tot, (double)qlen);
      if (qv < (double)p.min_qv) { qv_fail_read[r] = 1; *qv_fail_any = 1u; }
      else qname_seen[__ldg(b.qname_id + r)] = 1; // m.num_ccs counts it (caller.py:318-320)
    }
  }
  // what is left: every site of a pair that does not stream; nothing, normally, of one that does
  for (;;) {
    if (pending) { entries[my_at] = my_e | ((uint32_t)(my_q < qlen ? bqg[my_q] : 0) << 3); pending = false; }
    if (g_next >= s_hi) break;
    load_group();
  }
}

// sites keep only the candidates of reads that passed the QV gate; runs only when some read failed it
__global__ void __launch_bounds__(256) k_site_valid(const unsigned int* qv_fail_any, const uint8_t* qv_fail_read, const uint64_t* seg_off,
                                                    const uint32_t* seg_cnt, const uint32_t* seg_keys, const uint32_t* seg_read,
                                                    const uint32_t* key_site, const uint32_t* tile_off, const uint32_t* tile_dst,
                                                    const unsigned long long* n_keys_dev, uint8_t* site_valid) {
  if (!*qv_fail_any) return;
  const uint32_t c = blockIdx.x;
  const uint64_t seg0 = seg_off[c];
  const uint32_t nk = seg_cnt[c];
  const unsigned long long n_keys = *n_keys_dev;
  for (uint32_t j = threadIdx.x; j < nk; j += blockDim.x) {
    if (qv_fail_read[seg_read[seg0 + j]]) continue;
    const uint32_t rel = seg_keys[seg0 + j] >> 4;
    const uint64_t site = (uint64_t)tile_dst[tile_off[c] + (rel >> HC_TILE_BITS)] + key_site[seg0 + j];
    if (site < n_keys) site_valid[site] = 1;
  }
}

// ============================================================================ k_compact_sites
// Stable compaction of the records the host wants.  keep_cnt[j] = kept records of k_site_reduce's block j (128 sites);
// a CTA here covers 8 of them.  kpos[i] = where record i went (boundary records are looked up through it).
__global__ void __launch_bounds__(1024) k_compact_sites(const hm_site_record* rec, const unsigned long long* n_dev, const uint32_t* keep_cnt,
                                                        int omit, hm_site_record* out, uint32_t* kpos, unsigned long long* n_kept) {
  __shared__ uint32_t s_warp[33];
  const uint64_t n = *n_dev;
  const uint64_t i0 = (uint64_t)blockIdx.x * 1024;
  if (i0 >= n && blockIdx.x != 0) return;
  const uint32_t n_red = (uint32_t)((n + 127) / 128);
  // records kept before this CTA (and, in CTA 0, in total)
  const uint32_t before = min(blockIdx.x * 8u, n_red);
  uint32_t part = 0, all = 0;
  for (uint32_t j = threadIdx.x; j < (blockIdx.x == 0 ? n_red : before); j += blockDim.x) {
    const uint32_t v = keep_cnt[j];
    all += v;
    if (j < before) part += v;
  }
  uint32_t tot_before, tot_all;
  block_excl_scan(part, s_warp, &tot_before);
  if (blockIdx.x == 0) {
    block_excl_scan(all, s_warp, &tot_all);
    if (threadIdx.x == 0) *n_kept = tot_all;
  }
  const uint64_t i = i0 + threadIdx.x;
  bool keep = false;
  if (i < n) {
    const uint8_t st = rec[i].status;
    keep = st != HM_ST_INTERNAL_DROPPED && !(omit && st >= HM_ST_GERM_HET && st <= HM_ST_GERM_HOMREF);
  }
  uint32_t total;
  const uint32_t ex = block_excl_scan(keep ? 1u : 0u, s_warp, &total);
  if (i < n) {
    kpos[i] = tot_before + ex;
    if (keep) out[tot_before + ex] = rec[i];
  }
}
