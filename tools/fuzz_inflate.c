/* Fuzz harness of csrc/inflate_fast.h: pairs of random payloads deflated by zlib at random levels / strategies, then
 * decoded intact (must match) and with flipped bits, truncated input and wrong output sizes (must be rejected or at
 * least stay inside exact-size heap buffers).  Build with the sanitizers:
 *   gcc -O1 -g -std=c11 -fsanitize=address,undefined -I himut_b200/csrc -o fz tools/fuzz_inflate.c -lz && ./fz 3000
 * tests/test_inflate.py runs a short session of it. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <zlib.h>
#include "inflate_fast.h"
static uint64_t rs=88172645463325252ull; static uint32_t rnd(void){rs^=rs<<13;rs^=rs>>7;rs^=rs<<17;return (uint32_t)(rs>>11);}
static size_t make(uint8_t* d, size_t n, int kind){
  for(size_t i=0;i<n;i++){ switch(kind){case 0:d[i]=(uint8_t)rnd();break;case 1:d[i]="ACGT"[rnd()&3];break;case 2:d[i]=(rnd()%10)?93:(uint8_t)(rnd()%93);break;default:d[i]=(uint8_t)((i*7)^(i>>5));} }
  return n;}
int main(int argc, char** argv){
  static uint8_t data[2][70000], comp[2][80000];
  long ok=0,bad=0,iters=0;
  const int n_it = argc > 1 ? atoi(argv[1]) : 3000;
  for(int it=0;it<n_it;it++){
    size_t n[2],c[2];
    for(int k=0;k<2;k++){
      n[k]=1+rnd()%66000; make(data[k],n[k],rnd()%4);
      z_stream z;memset(&z,0,sizeof z);deflateInit2(&z,1+rnd()%9,Z_DEFLATED,-15,8,rnd()%3==0?Z_FIXED:Z_DEFAULT_STRATEGY);
      z.next_in=data[k];z.avail_in=(uInt)n[k];z.next_out=comp[k];z.avail_out=sizeof comp[k];deflate(&z,Z_FINISH);c[k]=z.total_out;deflateEnd(&z);
    }
    for(int m=0;m<8;m++){
      /* exact-size heap copies so that ASAN sees any overrun */
      uint8_t* in0=malloc(c[0]); uint8_t* in1=malloc(c[1]); memcpy(in0,comp[0],c[0]); memcpy(in1,comp[1],c[1]);
      size_t o0=n[0],o1=n[1];
      if(m){ int nf=1+rnd()%3; for(int f=0;f<nf;f++){ if(rnd()&1) in0[rnd()%c[0]]^=(uint8_t)(1u<<(rnd()&7)); else in1[rnd()%c[1]]^=(uint8_t)(1u<<(rnd()&7)); }
             if(rnd()%4==0) o0=o0>10?o0-rnd()%10:o0; if(rnd()%4==0) o1+=rnd()%10; }
      size_t l0=c[0], l1=c[1]; if(m&&rnd()%5==0) l0=rnd()%(c[0]+1); 
      uint8_t* out0=malloc(o0?o0:1); uint8_t* out1=malloc(o1?o1:1);
      int rc[2]; hm_inflate_raw2(in0,l0,out0,o0,in1,l1,out1,o1,rc);
      if(!m){ if(rc[0]||rc[1]||memcmp(out0,data[0],n[0])||memcmp(out1,data[1],n[1])){printf("MISMATCH it %d\n",it);return 1;} }
      int r1=hm_inflate_raw(in0,l0,out0,o0);
      if(!m && (r1||memcmp(out0,data[0],n[0]))){printf("MISMATCH single it %d\n",it);return 1;}
      (rc[0]||rc[1])?bad++:ok++; iters++;
      free(in0);free(in1);free(out0);free(out1);
    }
  }
  printf("iterations %ld ok %ld rejected %ld\n",iters,ok,bad);return 0;}
