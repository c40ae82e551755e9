/*
 * oracle/datagen.c -- synthetic workload generators (TEST/BENCH SUPPORT, not product).
 *
 * Implements the input shapes of SURVEY.md section 8(d) with one portable PRNG
 * (xorshift64*), so that C and Python agree byte for byte:
 *   C2  Zipf(s=1) text over a fixed vocabulary, 12 tokens per line
 *   C3  (i) period-1000 random unit tiled, (ii) "aab" tiled, (iii) long a/b/c runs
 *   R   raw PRNG bytes
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t xs_next(uint64_t* s)
{
   uint64_t x = *s;
   x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
   *s = x;
   return x * 0x2545F4914F6CDD1DULL;
}

/* vocab: `nvocab` tokens, token i at vocab + off[i], length len[i].
 * state: in/out PRNG state and position-in-line so streams can be continued. */
uint64_t gen_text(uint8_t* out, uint64_t n, const uint8_t* vocab, const int32_t* off,
                  const int32_t* len, int32_t nvocab, uint64_t* state, int32_t* tok_in_line)
{
   uint64_t* cum = malloc(sizeof(uint64_t) * (size_t)nvocab);
   uint64_t tot = 0, o = 0;
   for (int32_t i = 0; i < nvocab; i++) { tot += (1ULL << 32) / (uint64_t)(i + 1); cum[i] = tot; }
   int32_t til = *tok_in_line;
   while (o < n) {
      uint64_t r = xs_next(state) % tot;
      int32_t lo = 0, hi = nvocab - 1;
      while (lo < hi) { int32_t m = (lo + hi) >> 1; if (r < cum[m]) hi = m; else lo = m + 1; }
      int32_t L = len[lo];
      const uint8_t* t = vocab + off[lo];
      for (int32_t k = 0; k < L && o < n; k++) out[o++] = t[k];
      til++;
      if (o < n) out[o++] = (til == 12) ? '\n' : ' ';
      if (til == 12) til = 0;
   }
   *tok_in_line = til;
   free(cum);
   return o;
}

void gen_random(uint8_t* out, uint64_t n, uint64_t* state)
{
   uint64_t o = 0;
   while (o + 8 <= n) { uint64_t v = xs_next(state); memcpy(out + o, &v, 8); o += 8; }
   if (o < n) { uint64_t v = xs_next(state); memcpy(out + o, &v, (size_t)(n - o)); }
}

void gen_tile(uint8_t* out, uint64_t n, const uint8_t* unit, uint64_t ulen, uint64_t phase)
{
   for (uint64_t i = 0; i < n; i++) out[i] = unit[(phase + i) % ulen];
}

/* long runs of 'a'/'b'/'c' in the spirit of spewG.c:38-54 (run lengths from the PRNG) */
void gen_runs(uint8_t* out, uint64_t n, uint64_t* state)
{
   uint64_t o = 0;
   while (o < n) {
      uint64_t v = xs_next(state);
      uint8_t c = (uint8_t)('a' + (v >> 60) % 3);
      uint64_t L = 1 + (v & 0xffff) % 9000;
      if ((v >> 20 & 7) == 0) L = 1 + (v >> 24 & 7);
      for (uint64_t k = 0; k < L && o < n; k++) out[o++] = c;
   }
}
