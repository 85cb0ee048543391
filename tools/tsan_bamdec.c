/* ThreadSanitizer driver of the native BAM decoder (csrc/bamdec.c is included whole): region decode with and without
 * the base stream, compact quality stream, window pre-pass and the BAM writer, each on `threads` threads.
 *   gcc -O1 -g -fsanitize=thread -std=c11 -I himut_b200/csrc -I include -o /tmp/tsan_bamdec tools/tsan_bamdec.c -lz -lpthread
 *   setarch x86_64 -R /tmp/tsan_bamdec some_2mb.bam 8        (a BAM written by himut_b200.bamdec.write_batch_bam, contig >= 2 Mb)
 * 2026-10-18: one report (the lazily probed SSSE3 flag was first written by the record-decode threads; now probed in
 * hm_bam_open), none after the fix. */
#include "bamdec.c"
#include <stdio.h>
int main(int argc, char** argv) {
  hm_bam* b = NULL;
  if (hm_bam_open(argv[1], &b)) { printf("open failed\n"); return 1; }
  int threads = argc > 2 ? atoi(argv[2]) : 8;
  unsigned long sum = 0;
  for (int rep = 0; rep < 3; rep++) {
    hm_bam_set_option(b, HM_BAM_OPT_NO_SEQ, rep & 1);
    hm_read_batch rb;
    memset(&rb, 0, sizeof rb);
    int rc = hm_bam_read_batch(b, 0, rep * 100000, 2000000 - rep * 50000, threads, &rb);
    if (rc) { printf("read_batch rc %d: %s\n", rc, hm_bam_error(b)); return 1; }
    sum += rb.n_reads;
    /* compact quality stream on threads too */
    uint8_t* mask = calloc(rb.bq_bytes / 8 + 16, 1); uint64_t* eo = calloc(rb.n_reads + 1, 8); uint8_t* exc = NULL; uint64_t eb = 0; uint8_t modal = 0;
    rc = hm_bq_compact_build(&rb, threads, mask, eo, &exc, &eb, &modal);
    if (rc) { printf("compact rc %d\n", rc); return 1; }
    sum += eb + modal;
    hm_bq_compact_free(exc); free(mask); free(eo);
    int32_t q[100000]; size_t nq = 0;
    rc = hm_bam_window_qlens(b, 0, 300000, 400000, threads, q, 100000, &nq);
    if (rc) { printf("window rc %d\n", rc); return 1; }
    sum += nq;
    if (rep == 2) { rc = hm_bam_write_batch("/tmp/tsan_bamdec_out.bam", "chr1", 2000000, "synth", &rb, 1, threads); if (rc) { printf("write rc %d\n", rc); return 1; } }
  }
  printf("ok %lu qnames %u\n", sum, hm_bam_n_qnames(b));
  hm_bam_close(b);
  return 0;
}
