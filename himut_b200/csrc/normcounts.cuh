// normcounts.cuh — callable-base half of `himut normcounts` (placeholder until the tile kernel lands)
#pragma once
struct hm_ctx;
static int hm_normcounts_impl(hm_ctx* ctx, const uint8_t* refseq, size_t ref_len, const hm_chunk* chunks, size_t n_chunks,
                              int64_t* ccs_tri, int64_t* ref_tri, int64_t* log, int64_t* n_alt_tie);
