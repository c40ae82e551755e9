/*
 * multi_buff.c -- a plain C client of the drop-in library: compresses a file with BZ2_bzBuffToBuffCompress
 * (the reference's one-shot entry point, bzlib.c:1309-1357) and writes the .bz2 to stdout.  Nothing in here knows
 * about GPUs: BZ2_B200_DEVICES=0,1,2,3,4,5,6,7 in the environment makes the library shard the stream by block over
 * those devices (csrc/multi.cu).  Used by tests/test_gpu_multi.py; build: see tests/c/Makefile.
 */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "bzlib.h"

int main(int argc, char** argv)
{
   FILE* f;
   long n;
   char *src, *dst;
   unsigned int dlen;
   int level = 9, rc, reps = 1, k;
   struct timespec t0, t1;
   if (argc < 2) { fprintf(stderr, "usage: %s file [level] [reps]\n", argv[0]); return 2; }
   if (argc > 2) level = atoi(argv[2]);
   if (argc > 3) reps = atoi(argv[3]);
   f = fopen(argv[1], "rb");
   if (!f) { perror(argv[1]); return 2; }
   fseek(f, 0, SEEK_END); n = ftell(f); fseek(f, 0, SEEK_SET);
   src = malloc(n ? n : 1);
   dlen = (unsigned int)(n + n / 50 + 24576u * (n / (100000 * level - 19) + 2) + 1024);
   dst = malloc(dlen);
   if (!src || !dst || fread(src, 1, n, f) != (size_t)n) { fprintf(stderr, "read failed\n"); return 2; }
   fclose(f);
   for (k = 0; k < reps; k++) {
      unsigned int cap = dlen;
      clock_gettime(CLOCK_MONOTONIC, &t0);
      rc = BZ2_bzBuffToBuffCompress(dst, &cap, src, (unsigned int)n, level, 0, 0);
      clock_gettime(CLOCK_MONOTONIC, &t1);
      if (rc != BZ_OK) { fprintf(stderr, "BZ2_bzBuffToBuffCompress: %d\n", rc); return 1; }
      fprintf(stderr, "pass %d: %ld -> %u bytes, %.1f MB/s\n", k, n, cap,
              n / 1e6 / ((t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9));
      if (k == reps - 1) fwrite(dst, 1, cap, stdout);
   }
   return 0;
}
