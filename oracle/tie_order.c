/*
 * oracle/tie_order.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see bz2_oracle.h).
 *
 * Tie-order emulator for exact-power blocks.
 *
 * When the post-RLE1 block is u^q (q >= 2) every rotation has q equal copies and the
 * reference's origPtr = lo + g, where g is "which of the q equal copies of rotation 0 ended
 * up where" inside the reference's sorter.  Nothing about g is canonical: it is the trace of
 * the swaps the reference's sorter performs on elements that compare equal.  The only way to
 * reproduce it for every (u, q) is to replay those steps.  This file restates, with integer
 * offsets into one work array instead of pointers, the part of the reference's sorter that
 * orders the "B*" suffixes (positions i with T[i] < T[i+1] whose successor run is followed by
 * a descent):
 *
 *   blocksort.c:1316-1401  sort_typeBstar up to and including trsort
 *   blocksort.c:85-148     substring compares (plain, merge, last-suffix with cyclic wrap)
 *   blocksort.c:152-424    insertion sort, heap sort, pivots, multikey introsort
 *   blocksort.c:428-622    block swap, forward / backward merges, ss_merge
 *   blocksort.c:627-664    substringsort: 1024-element chunks, merges, last-suffix re-insertion
 *   blocksort.c:669-806    rank-sort helpers (cyclic key fetch, heap / insertion sort, pivots)
 *   blocksort.c:814-959    lssort (Larsson-Sadakane fallback once the budget is spent)
 *   blocksort.c:966-1309   tr_partition, tr_copy, tr_introsort, trsort
 *
 * The two induced scans (blocksort.c:1441-1501) are NOT replayed: they keep the relative order
 * of the copies of a class, so rotation 0 lands, among its q copies, where the first B*
 * suffix of the block (B* index 0) lands among the copies of ITS class.  The result is that
 * count, g.  Checked against the compiled reference in tests/test_tie_order.py.
 */
#include "bz2_oracle.h"
#include <stdlib.h>
#include <string.h>

#define CHUNK 1024   /* blocksort.c:36 */
#define SMALL 8      /* blocksort.c:37 */
#define STK   64     /* blocksort.c:33 */

typedef struct {
   const uint8_t* T;   /* block bytes; T[n] == T[0] readable (blocksort.c:1540) */
   int32_t* w;         /* work array, n + 256 ints: [0,m) order, [m,..) buffer / ranks, [n-m,n) B* positions */
   int32_t n, m, pa;   /* pa = n - m */
   int32_t isa;        /* offset of the rank array (== m) */
} tctx;

#define PA(i)  (c->w[c->pa + (i)])
#define W      (c->w)
#define SWAPW(x, y) do { int32_t t_ = W[x]; W[x] = W[y]; W[y] = t_; } while (0)

static inline int ilog2(int32_t v) { int r = -1; while (v > 0) { v >>= 1; r++; } return r; }

/* ------------------------------------------------------------- substring compares -- */
/* blocksort.c:85-101 (depth d counted past the two bucket characters) and :105-120 (d = 0). */
static int sub_cmp(const tctx* c, int32_t i1, int32_t i2, int32_t d)
{
   int32_t s1 = PA(i1) + 2 + d, e1 = PA(i1 + 1) + 2;
   int32_t s2 = PA(i2) + 2 + d, e2 = PA(i2 + 1) + 2;
   while (s1 < e1 && s2 < e2 && c->T[s1] == c->T[s2]) { s1++; s2++; }
   if (s1 < e1) return s2 < e2 ? (int)c->T[s1] - (int)c->T[s2] : 1;
   return s2 < e2 ? -1 : 0;
}

/* blocksort.c:124-148: i1 is the last B* suffix of the block; its substring runs to the end of
 * the block and continues from the start up to the first B* suffix. */
static int sub_cmp_last(const tctx* c, int32_t i1, int32_t i2)
{
   int32_t s1 = PA(i1) + 2, e1 = c->n;
   int32_t s2 = PA(i2) + 2, e2 = PA(i2 + 1) + 2;
   while (s1 < e1 && s2 < e2 && c->T[s1] == c->T[s2]) { s1++; s2++; }
   if (s1 < e1) return s2 < e2 ? (int)c->T[s1] - (int)c->T[s2] : 1;
   if (s2 == e2) return 1;
   s1 -= c->n; e1 = PA(0) + 2;
   while (s1 < e1 && s2 < e2 && c->T[s1] == c->T[s2]) { s1++; s2++; }
   if (s1 < e1) return s2 < e2 ? (int)c->T[s1] - (int)c->T[s2] : 1;
   return s2 < e2 ? -1 : 0;
}

/* character of element x at depth d, and the one before it */
#define KEY(x, d)  ((int)c->T[PA(x) + 2 + (d)])

/* ----------------------------------------------------------- small substring sorts -- */
/* blocksort.c:152-166 */
static void ss_isort(tctx* c, int32_t first, int32_t last, int32_t d)
{
   for (int32_t i = last - 2; first <= i; --i) {
      int32_t t = W[i], j = i + 1;
      int r;
      while (0 < (r = sub_cmp(c, t, W[j], d))) {
         do { W[j - 1] = W[j]; } while (++j < last && W[j] < 0);
         if (last <= j) break;
      }
      if (r == 0) W[j] = ~W[j];
      W[j - 1] = t;
   }
}

/* blocksort.c:170-182; base = offset of the heap's element 0 */
static void ss_sift(tctx* c, int32_t d, int32_t base, int32_t i, int32_t size)
{
   int32_t v = W[base + i], j, k;
   int cv = KEY(v, d), x, y;
   for (; (j = 2 * i + 1) < size; W[base + i] = W[base + k], i = k) {
      k = j++;
      x = KEY(W[base + k], d);
      if (x < (y = KEY(W[base + j], d))) { k = j; x = y; }
      if (x <= cv) break;
   }
   W[base + i] = v;
}

/* blocksort.c:186-209 */
static void ss_hsort(tctx* c, int32_t d, int32_t base, int32_t size)
{
   int32_t i, mm = size;
   if ((size % 2) == 0) {
      mm--;
      if (KEY(W[base + mm / 2], d) < KEY(W[base + mm], d)) SWAPW(base + mm, base + mm / 2);
   }
   for (i = mm / 2 - 1; 0 <= i; --i) ss_sift(c, d, base, i, mm);
   if ((size % 2) == 0) { SWAPW(base, base + mm); ss_sift(c, d, base, 0, mm); }
   for (i = mm - 1; 0 < i; --i) {
      int32_t t = W[base];
      W[base] = W[base + i];
      ss_sift(c, d, base, 0, i);
      W[base + i] = t;
   }
}

/* blocksort.c:213-222, :226-236, :240-262: the arguments and results are offsets into W */
static int32_t ss_med3(const tctx* c, int32_t d, int32_t v1, int32_t v2, int32_t v3)
{
   int32_t t;
   if (KEY(W[v1], d) > KEY(W[v2], d)) { t = v1; v1 = v2; v2 = t; }
   if (KEY(W[v2], d) > KEY(W[v3], d)) return KEY(W[v1], d) > KEY(W[v3], d) ? v1 : v3;
   return v2;
}
static int32_t ss_med5(const tctx* c, int32_t d, int32_t v1, int32_t v2, int32_t v3, int32_t v4, int32_t v5)
{
   int32_t t;
   if (KEY(W[v2], d) > KEY(W[v3], d)) { t = v2; v2 = v3; v3 = t; }
   if (KEY(W[v4], d) > KEY(W[v5], d)) { t = v4; v4 = v5; v5 = t; }
   if (KEY(W[v2], d) > KEY(W[v4], d)) { t = v2; v2 = v4; v4 = t; t = v3; v3 = v5; v5 = t; }
   if (KEY(W[v1], d) > KEY(W[v3], d)) { t = v1; v1 = v3; v3 = t; }
   if (KEY(W[v1], d) > KEY(W[v4], d)) { t = v1; v1 = v4; v4 = t; t = v3; v3 = v5; v5 = t; }
   if (KEY(W[v3], d) > KEY(W[v4], d)) return v4;
   return v3;
}
static int32_t ss_pick(const tctx* c, int32_t d, int32_t first, int32_t last)
{
   int32_t t = last - first, mid = first + t / 2;
   if (t <= 512) {
      if (t <= 32) return ss_med3(c, d, first, mid, last - 1);
      t >>= 2;
      return ss_med5(c, d, first, first + t, mid, last - 1 - t, last - 1);
   }
   t >>= 3;
   return ss_med3(c, d, ss_med3(c, d, first, first + t, first + (t << 1)),
                  ss_med3(c, d, mid - t, mid, mid + t),
                  ss_med3(c, d, last - 1 - (t << 1), last - 1 - t, last - 1));
}

/* blocksort.c:284-298: split off (and mark as sorted) the elements whose substring ends here */
static int32_t ss_ended(tctx* c, int32_t first, int32_t last, int32_t d)
{
   int32_t a = first - 1, b = last, t;
   for (;;) {
      for (; ++a < b && (PA(W[a]) + d) >= (PA(W[a] + 1) - 1);) W[a] = ~W[a];
      for (; a < --b && (PA(W[b]) + d) < (PA(W[b] + 1) - 1);) { }
      if (b <= a) break;
      t = ~W[b]; W[b] = W[a]; W[a] = t;
   }
   if (first < a) W[first] = ~W[first];
   return a;
}

/* blocksort.c:302-424 */
static void ss_mkqsort(tctx* c, int32_t first, int32_t last)
{
   struct { int32_t a, b, c, d; } stack[STK];
   int sp = 0;
   int32_t d = 0, a, b, cc, dd, e, f, s, t;
   int limit = ilog2(last - first), v, x = 0;
#define PUSH(A, B, C, D) do { stack[sp].a = (A); stack[sp].b = (B); stack[sp].c = (C); stack[sp].d = (D); sp++; } while (0)
#define POP() do { if (sp == 0) return; sp--; first = stack[sp].a; last = stack[sp].b; d = stack[sp].c; limit = stack[sp].d; } while (0)
   for (;;) {
      if (last - first <= SMALL) {
         if (1 < last - first) ss_isort(c, first, last, d);
         POP();
         continue;
      }
      if (limit-- == 0) ss_hsort(c, d, first, last - first);
      if (limit < 0) {
         for (a = first + 1, v = KEY(W[first], d); a < last; ++a) {
            if ((x = KEY(W[a], d)) != v) {
               if (1 < a - first) break;
               v = x; first = a;
            }
         }
         if (KEY(W[first], d - 1) < v) first = ss_ended(c, first, a, d);
         if (a - first <= last - a) {
            if (1 < a - first) { PUSH(a, last, d, -1); last = a; d += 1; limit = ilog2(a - first); }
            else { first = a; limit = -1; }
         } else {
            if (1 < last - a) { PUSH(first, a, d + 1, ilog2(a - first)); first = a; limit = -1; }
            else { last = a; d += 1; limit = ilog2(a - first); }
         }
         continue;
      }

      a = ss_pick(c, d, first, last);
      v = KEY(W[a], d);
      SWAPW(first, a);

      for (b = first; ++b < last && (x = KEY(W[b], d)) == v;) { }
      if ((a = b) < last && x < v) {
         for (; ++b < last && (x = KEY(W[b], d)) <= v;) if (x == v) { SWAPW(b, a); ++a; }
      }
      for (cc = last; b < --cc && (x = KEY(W[cc], d)) == v;) { }
      if (b < (dd = cc) && x > v) {
         for (; b < --cc && (x = KEY(W[cc], d)) >= v;) if (x == v) { SWAPW(cc, dd); --dd; }
      }
      for (; b < cc;) {
         SWAPW(b, cc);
         for (; ++b < cc && (x = KEY(W[b], d)) <= v;) if (x == v) { SWAPW(b, a); ++a; }
         for (; b < --cc && (x = KEY(W[cc], d)) >= v;) if (x == v) { SWAPW(cc, dd); --dd; }
      }

      if (a <= dd) {
         cc = b - 1;
         if ((s = a - first) > (t = b - a)) s = t;
         for (e = first, f = b - s; 0 < s; --s, ++e, ++f) SWAPW(e, f);
         if ((s = dd - cc) > (t = last - dd - 1)) s = t;
         for (e = b, f = last - s; 0 < s; --s, ++e, ++f) SWAPW(e, f);

         a = first + (b - a); cc = last - (dd - cc);
         b = (v <= KEY(W[a], d - 1)) ? a : ss_ended(c, a, cc, d);

         if (a - first <= last - cc) {
            if (last - cc <= cc - b) {
               PUSH(b, cc, d + 1, ilog2(cc - b)); PUSH(cc, last, d, limit); last = a;
            } else if (a - first <= cc - b) {
               PUSH(cc, last, d, limit); PUSH(b, cc, d + 1, ilog2(cc - b)); last = a;
            } else {
               PUSH(cc, last, d, limit); PUSH(first, a, d, limit);
               first = b; last = cc; d += 1; limit = ilog2(cc - b);
            }
         } else {
            if (a - first <= cc - b) {
               PUSH(b, cc, d + 1, ilog2(cc - b)); PUSH(first, a, d, limit); first = cc;
            } else if (last - cc <= cc - b) {
               PUSH(first, a, d, limit); PUSH(b, cc, d + 1, ilog2(cc - b)); first = cc;
            } else {
               PUSH(first, a, d, limit); PUSH(cc, last, d, limit);
               first = b; last = cc; d += 1; limit = ilog2(cc - b);
            }
         }
      } else {
         limit += 1;
         if (KEY(W[first], d - 1) < v) { first = ss_ended(c, first, last, d); limit = ilog2(last - first); }
         d += 1;
      }
   }
#undef PUSH
#undef POP
}

/* ------------------------------------------------------------------------ merges -- */
/* blocksort.c:428-434 */
static void blk_swap(tctx* c, int32_t x, int32_t y, int32_t size)
{
   for (; 0 < size; --size, ++x, ++y) SWAPW(x, y);
}

/* blocksort.c:438-481 */
static void mrg_fwd(tctx* c, int32_t buf, int32_t first, int32_t middle, int32_t last)
{
   int32_t bufend = buf + (middle - first), i, j, k, t;
   int r;
   blk_swap(c, buf, first, middle - first);
   for (t = W[first], i = first, j = buf, k = middle;;) {
      r = sub_cmp(c, W[j], W[k], 0);
      if (r < 0) {
         do {
            W[i++] = W[j]; W[j++] = W[i];
            if (bufend <= j) { W[bufend - 1] = t; return; }
         } while (W[j] < 0);
      } else if (r > 0) {
         do {
            W[i++] = W[k]; W[k++] = W[i];
            if (last <= k) {
               do { W[i++] = W[j]; W[j++] = W[i]; } while (j < bufend);
               W[bufend - 1] = t;
               return;
            }
         } while (W[k] < 0);
      } else {
         W[k] = ~W[k];
         do {
            W[i++] = W[j]; W[j++] = W[i];
            if (bufend <= j) { W[bufend - 1] = t; return; }
         } while (W[j] < 0);
         do {
            W[i++] = W[k]; W[k++] = W[i];
            if (last <= k) {
               do { W[i++] = W[j]; W[j++] = W[i]; } while (j < bufend);
               W[bufend - 1] = t;
               return;
            }
         } while (W[k] < 0);
      }
   }
}

/* blocksort.c:485-542 */
static void mrg_bwd(tctx* c, int32_t buf, int32_t first, int32_t middle, int32_t last)
{
   int32_t bufend = buf + (last - middle), i, j, k, t, p1, p2;
   int r, x = 0;
   blk_swap(c, buf, middle, last - middle);
   if (W[bufend - 1] < 0) { x |= 1; p1 = ~W[bufend - 1]; } else p1 = W[bufend - 1];
   if (W[middle - 1] < 0) { x |= 2; p2 = ~W[middle - 1]; } else p2 = W[middle - 1];
   for (t = W[last - 1], i = last - 1, j = bufend - 1, k = middle - 1;;) {
      r = sub_cmp(c, p1, p2, 0);
      if (r > 0) {
         if (x & 1) { do { W[i--] = W[j]; W[j--] = W[i]; } while (W[j] < 0); }
         W[i--] = W[j]; W[j--] = W[i];
         if (j < buf) { W[buf] = t; return; }
         if (W[j] < 0) { x |= 1; p1 = ~W[j]; } else { x &= ~1; p1 = W[j]; }
      } else if (r < 0) {
         if (x & 2) { do { W[i--] = W[k]; W[k--] = W[i]; } while (W[k] < 0); }
         W[i--] = W[k]; W[k--] = W[i];
         if (k < first) {
            do { W[i--] = W[j]; W[j--] = W[i]; } while (buf <= j);
            W[buf] = t;
            return;
         }
         if (W[k] < 0) { x |= 2; p2 = ~W[k]; } else { x &= ~2; p2 = W[k]; }
      } else {
         if (x & 1) { do { W[i--] = W[j]; W[j--] = W[i]; } while (W[j] < 0); }
         W[i--] = ~W[j]; W[j--] = W[i];
         if (j < buf) { W[buf] = t; return; }
         if (x & 2) { do { W[i--] = W[k]; W[k--] = W[i]; } while (W[k] < 0); }
         W[i--] = W[k]; W[k--] = W[i];
         if (k < first) {
            while (buf <= j) { W[i--] = W[j]; W[j--] = W[i]; }
            W[buf] = t;
            return;
         }
         if (W[j] < 0) { x |= 1; p1 = ~W[j]; } else { x &= ~1; p1 = W[j]; }
         if (W[k] < 0) { x |= 2; p2 = ~W[k]; } else { x &= ~2; p2 = W[k]; }
      }
   }
}

#define IDX(v) ((0 <= (v)) ? (v) : ~(v))
/* blocksort.c:550-556 */
static void mark_if_equal(tctx* c, int32_t a)
{
   if (0 <= W[a] && sub_cmp(c, IDX(W[a - 1]), W[a], 0) == 0) W[a] = ~W[a];
}

/* blocksort.c:546-619 */
static void ss_mrg(tctx* c, int32_t first, int32_t middle, int32_t last, int32_t buf, int32_t bufsize)
{
   struct { int32_t a, b, c; int d; } stack[STK];
   int sp = 0, check = 0, next;
   int32_t i, j, mm, len, half;
#define POP() do { if (sp == 0) return; sp--; first = stack[sp].a; middle = stack[sp].b; last = stack[sp].c; check = stack[sp].d; } while (0)
#define PUSH(A, B, C, D) do { stack[sp].a = (A); stack[sp].b = (B); stack[sp].c = (C); stack[sp].d = (D); sp++; } while (0)
   for (;;) {
      if (last - middle <= bufsize) {
         if (first < middle && middle < last) mrg_bwd(c, buf, first, middle, last);
         if (check & 1) mark_if_equal(c, first);
         if (check & 2) mark_if_equal(c, last);
         POP();
         continue;
      }
      if (middle - first <= bufsize) {
         if (first < middle) mrg_fwd(c, buf, first, middle, last);
         if (check & 1) mark_if_equal(c, first);
         if (check & 2) mark_if_equal(c, last);
         POP();
         continue;
      }
      len = (middle - first < last - middle) ? middle - first : last - middle;
      for (mm = 0, half = len >> 1; 0 < len; len = half, half >>= 1) {
         if (sub_cmp(c, IDX(W[middle + mm + half]), IDX(W[middle - mm - half - 1]), 0) < 0) {
            mm += half + 1;
            half -= (len & 1) ^ 1;
         }
      }
      if (0 < mm) {
         blk_swap(c, middle - mm, middle, mm);
         i = j = middle; next = 0;
         if (middle + mm < last) {
            if (W[middle + mm] < 0) {
               for (; W[i - 1] < 0; --i) { }
               W[middle + mm] = ~W[middle + mm];
            }
            for (j = middle; W[j] < 0; ++j) { }
            next = 1;
         }
         if (i - first <= last - j) {
            PUSH(j, middle + mm, last, (check & 2) | (next & 1));
            middle -= mm; last = i; check = (check & 1);
         } else {
            if (i == middle && middle == j) next <<= 1;
            PUSH(first, middle - mm, i, (check & 1) | (next & 2));
            first = j; middle += mm; check = (check & 2) | (next & 1);
         }
      } else {
         if (check & 1) mark_if_equal(c, first);
         mark_if_equal(c, middle);
         if (check & 2) mark_if_equal(c, last);
         POP();
      }
   }
#undef PUSH
#undef POP
}

/* blocksort.c:627-664 */
static void sort_bucket(tctx* c, int32_t first, int32_t last, int32_t buf, int32_t bufsize, int lastsuffix)
{
   int32_t a, b, curbuf, curbufsize, i, j, k;
   if (lastsuffix) ++first;
   for (a = first, i = 0; a + CHUNK < last; a += CHUNK, ++i) {
      ss_mkqsort(c, a, a + CHUNK);
      curbuf = a + CHUNK;
      curbufsize = last - (a + CHUNK);
      if (curbufsize <= bufsize) { curbufsize = bufsize; curbuf = buf; }
      for (b = a, k = CHUNK, j = i; j & 1; b -= k, k <<= 1, j >>= 1) ss_mrg(c, b - k, b, b + k, curbuf, curbufsize);
   }
   ss_mkqsort(c, a, last);
   for (k = CHUNK; i != 0; k <<= 1, i >>= 1) {
      if (i & 1) { ss_mrg(c, a - k, a, last, buf, bufsize); a -= k; }
   }
   if (lastsuffix) {
      int r = 1;
      int32_t li = W[first - 1];
      for (a = first; a < last && (W[a] < 0 || 0 < (r = sub_cmp_last(c, li, W[a]))); ++a) W[a - 1] = W[a];
      if (r == 0) W[a] = ~W[a];
      W[a - 1] = li;
   }
}

/* -------------------------------------------------------------------- rank sort -- */
/* blocksort.c:669: rank of the B* suffix dd places after p, cyclically over the m of them */
#define ISA(i)        (W[c->isa + (i)])
static inline int32_t getc_(const tctx* c, int32_t dd, int32_t p)
{
   return (dd + p < c->m) ? W[c->isa + dd + p] : W[c->isa + dd + p - c->m];
}
#define GETC(p) getc_(c, dd, (p))

/* blocksort.c:673-685 */
static void tr_sift(tctx* c, int32_t dd, int32_t base, int32_t i, int32_t size)
{
   int32_t v = W[base + i], cv = GETC(v), j, k, x, y;
   for (; (j = 2 * i + 1) < size; W[base + i] = W[base + k], i = k) {
      k = j++;
      x = GETC(W[base + k]);
      if (x < (y = GETC(W[base + j]))) { k = j; x = y; }
      if (x <= cv) break;
   }
   W[base + i] = v;
}
/* blocksort.c:689-716 */
static void tr_hsort(tctx* c, int32_t dd, int32_t base, int32_t size)
{
   int32_t i, mm = size;
   if ((size % 2) == 0) {
      mm--;
      if (GETC(W[base + mm / 2]) < GETC(W[base + mm])) SWAPW(base + mm, base + mm / 2);
   }
   for (i = mm / 2 - 1; 0 <= i; --i) tr_sift(c, dd, base, i, mm);
   if ((size % 2) == 0) { SWAPW(base, base + mm); tr_sift(c, dd, base, 0, mm); }
   for (i = mm - 1; 0 < i; --i) {
      int32_t t = W[base];
      W[base] = W[base + i];
      tr_sift(c, dd, base, 0, i);
      W[base + i] = t;
   }
}
/* blocksort.c:720-732 */
static void tr_isort(tctx* c, int32_t dd, int32_t first, int32_t last)
{
   for (int32_t a = first + 1; a < last; ++a) {
      int32_t t = W[a], b = a - 1, r;
      while (0 > (r = GETC(t) - GETC(W[b]))) {
         do { W[b + 1] = W[b]; } while (first <= --b && W[b] < 0);
         if (b < first) break;
      }
      if (r == 0) W[b] = ~W[b];
      W[b + 1] = t;
   }
}
/* blocksort.c:758-766, :770-780, :784-807 */
static int32_t tr_med3(const tctx* c, int32_t dd, int32_t v1, int32_t v2, int32_t v3)
{
   int32_t t;
   if (GETC(W[v1]) > GETC(W[v2])) { t = v1; v1 = v2; v2 = t; }
   if (GETC(W[v2]) > GETC(W[v3])) return GETC(W[v1]) > GETC(W[v3]) ? v1 : v3;
   return v2;
}
static int32_t tr_med5(const tctx* c, int32_t dd, int32_t v1, int32_t v2, int32_t v3, int32_t v4, int32_t v5)
{
   int32_t t;
   if (GETC(W[v2]) > GETC(W[v3])) { t = v2; v2 = v3; v3 = t; }
   if (GETC(W[v4]) > GETC(W[v5])) { t = v4; v4 = v5; v5 = t; }
   if (GETC(W[v2]) > GETC(W[v4])) { t = v2; v2 = v4; v4 = t; t = v3; v3 = v5; v5 = t; }
   if (GETC(W[v1]) > GETC(W[v3])) { t = v1; v1 = v3; v3 = t; }
   if (GETC(W[v1]) > GETC(W[v4])) { t = v1; v1 = v4; v4 = t; t = v3; v3 = v5; v5 = t; }
   if (GETC(W[v3]) > GETC(W[v4])) return v4;
   return v3;
}
static int32_t tr_pick(const tctx* c, int32_t dd, int32_t first, int32_t last)
{
   int32_t t = last - first, mid = first + t / 2;
   if (t <= 512) {
      if (t <= 32) return tr_med3(c, dd, first, mid, last - 1);
      t >>= 2;
      return tr_med5(c, dd, first, first + t, mid, last - 1 - t, last - 1);
   }
   t >>= 3;
   return tr_med3(c, dd, tr_med3(c, dd, first, first + t, first + (t << 1)),
                  tr_med3(c, dd, mid - t, mid, mid + t),
                  tr_med3(c, dd, last - 1 - (t << 1), last - 1 - t, last - 1));
}

/* three-way split around v shared by blocksort.c:869-899, :973-1002, :1168-1198: on return the
 * range is [< v | == v | > v] and *pa, *pb bound the middle part; returns 0 when every key equals v
 * (nothing moved apart from what the caller already did). */
static int tr_split3(tctx* c, int32_t dd, int32_t first, int32_t last, int32_t start, int32_t v,
                     int32_t* pa, int32_t* pb)
{
   int32_t a, b, cc, d2, e, f, s, t, x = 0;
   for (b = start; ++b < last && (x = GETC(W[b])) == v;) { }
   if ((a = b) < last && x < v) {
      for (; ++b < last && (x = GETC(W[b])) <= v;) if (x == v) { SWAPW(b, a); ++a; }
   }
   for (cc = last; b < --cc && (x = GETC(W[cc])) == v;) { }
   if (b < (d2 = cc) && x > v) {
      for (; b < --cc && (x = GETC(W[cc])) >= v;) if (x == v) { SWAPW(cc, d2); --d2; }
   }
   for (; b < cc;) {
      SWAPW(b, cc);
      for (; ++b < cc && (x = GETC(W[b])) <= v;) if (x == v) { SWAPW(b, a); ++a; }
      for (; b < --cc && (x = GETC(W[cc])) >= v;) if (x == v) { SWAPW(cc, d2); --d2; }
   }
   if (a <= d2) {
      cc = b - 1;
      if ((s = a - first) > (t = b - a)) s = t;
      for (e = first, f = b - s; 0 < s; --s, ++e, ++f) SWAPW(e, f);
      if ((s = d2 - cc) > (t = last - d2 - 1)) s = t;
      for (e = b, f = last - s; 0 < s; --s, ++e, ++f) SWAPW(e, f);
      *pa = first + (b - a); *pb = last - (d2 - cc);
      return 1;
   }
   *pa = first; *pb = last;
   return 0;
}

/* blocksort.c:814-831; positions are relative to the start of the order array (offset 0) */
static void ls_regroup(tctx* c, int32_t first, int32_t last)
{
   int32_t a, b, t;
   for (a = first; a < last; ++a) {
      if (0 <= W[a]) {
         b = a;
         do { ISA(W[a]) = a; } while (++a < last && 0 <= W[a]);
         W[b] = b - a;
         if (last <= a) break;
      }
      b = a;
      do { W[a] = ~W[a]; } while (W[++a] < 0);
      t = a;
      do { ISA(W[b]) = t; } while (++b <= a);
   }
}

/* blocksort.c:835-924 */
static void ls_qsort(tctx* c, int32_t dd, int32_t first, int32_t last)
{
   struct { int32_t a, b; int c; } stack[STK];
   int sp = 0, limit = ilog2(last - first);
   int32_t a, b, cc, v, x;
#define POP() do { if (sp == 0) return; sp--; first = stack[sp].a; last = stack[sp].b; limit = stack[sp].c; } while (0)
#define PUSH(A, B, C) do { stack[sp].a = (A); stack[sp].b = (B); stack[sp].c = (C); sp++; } while (0)
   for (;;) {
      if (last - first <= SMALL) {
         if (1 < last - first) { tr_isort(c, dd, first, last); ls_regroup(c, first, last); }
         else if (last - first == 1) W[first] = -1;
         POP();
         continue;
      }
      if (limit-- == 0) {
         tr_hsort(c, dd, first, last - first);
         for (a = last - 2, v = GETC(W[last - 1]); first <= a; --a) {
            if ((x = GETC(W[a])) == v) W[a] = ~W[a]; else v = x;
         }
         ls_regroup(c, first, last);
         POP();
         continue;
      }
      a = tr_pick(c, dd, first, last);
      SWAPW(first, a);
      v = GETC(W[first]);
      if (tr_split3(c, dd, first, last, first, v, &a, &b)) {
         for (cc = first, v = a - 1; cc < a; ++cc) ISA(W[cc]) = v;
         if (b < last) { for (cc = a, v = b - 1; cc < b; ++cc) ISA(W[cc]) = v; }
         if (b - a == 1) W[a] = -1;
         if (a - first <= last - b) {
            if (first < a) { PUSH(b, last, limit); last = a; } else first = b;
         } else {
            if (b < last) { PUSH(first, a, limit); first = b; } else last = a;
         }
      } else {
         POP();
      }
   }
#undef PUSH
#undef POP
}

/* blocksort.c:928-959 */
static void ls_sort(tctx* c, int32_t depth)
{
   const int32_t n = c->m;
   int32_t dd, first, last, i, t, skip;
   for (dd = depth; -n < W[0]; dd += dd) {
      first = 0; skip = 0;
      do {
         if ((t = W[first]) < 0) { first -= t; skip += t; }
         else {
            if (skip != 0) { W[first + skip] = skip; skip = 0; }
            last = ISA(t) + 1;
            ls_qsort(c, dd, first, last);
            first = last;
         }
      } while (first < n);
      if (skip != 0) W[first + skip] = skip;
      if (n < dd) {
         first = 0;
         do {
            if ((t = W[first]) < 0) first -= t;
            else {
               last = ISA(t) + 1;
               for (i = first; i < last; ++i) ISA(W[i]) = i;
               first = last;
            }
         } while (first < n);
         break;
      }
   }
}

/* blocksort.c:1008-1029 */
static void tr_spread(tctx* c, int32_t first, int32_t a, int32_t b, int32_t last, int32_t depth)
{
   int32_t cc, d, e, s, v = b - 1;
   for (cc = first, d = a - 1; cc <= d; ++cc) {
      if ((s = W[cc] - depth) < 0) s += c->m;
      if (ISA(s) == v) { W[++d] = s; ISA(s) = d; }
   }
   for (cc = last - 1, e = d + 1, d = b; e < d; --cc) {
      if ((s = W[cc] - depth) < 0) s += c->m;
      if (ISA(s) == v) { W[--d] = s; ISA(s) = d; }
   }
}

/* blocksort.c:1033-1281.  Stack entries carry the key offset dd (-1 stands for the reference's NULL). */
static void tr_qsort(tctx* c, int32_t dd, int32_t first, int32_t last, int64_t* budget)
{
   struct { int32_t a, b, c; int d; } stack[STK];
   int sp = 0;
   int32_t a, b, cc, v, x;
   int limit = ilog2(last - first), next;
#define POP() do { if (sp == 0) return; sp--; dd = stack[sp].a; first = stack[sp].b; last = stack[sp].c; limit = stack[sp].d; } while (0)
#define PUSH(A, B, C, D) do { stack[sp].a = (A); stack[sp].b = (B); stack[sp].c = (C); stack[sp].d = (D); sp++; } while (0)
   for (;;) {
      if (limit < 0) {
         if (limit == -1) {
            /* tandem repeat: split by the rank one place earlier around "my own group" */
            tr_split3(c, dd - 1, first, last, first - 1, last - 1, &a, &b);
            if (first < a || b < last) {
               if (a < last) { for (cc = first, v = a - 1; cc < a; ++cc) ISA(W[cc]) = v; }
               if (b < last) { for (cc = a, v = b - 1; cc < b; ++cc) ISA(W[cc]) = v; }
               if (1 < b - a) { PUSH(-1, a, b, 0); PUSH(dd - 1, first, last, -2); }
               if (a - first <= last - b) {
                  if (1 < a - first) { PUSH(dd, b, last, ilog2(last - b)); last = a; limit = ilog2(a - first); }
                  else if (1 < last - b) { first = b; limit = ilog2(last - b); }
                  else POP();
               } else {
                  if (1 < last - b) { PUSH(dd, first, a, ilog2(a - first)); first = b; limit = ilog2(last - b); }
                  else if (1 < a - first) { last = a; limit = ilog2(a - first); }
                  else POP();
               }
            } else {
               for (cc = first; cc < last; ++cc) ISA(W[cc]) = cc;
               POP();
            }
         } else if (limit == -2) {
            sp--; a = stack[sp].b; b = stack[sp].c;
            tr_spread(c, first, a, b, last, dd);
            POP();
         } else {
            if (0 <= W[first]) {
               a = first;
               do { ISA(W[a]) = a; } while (++a < last && 0 <= W[a]);
               first = a;
            }
            if (first < last) {
               b = first; do { W[b] = ~W[b]; } while (W[++b] < 0);
               a = b + 1;
               next = (ISA(W[b]) != GETC(W[b])) ? ilog2(a - first) : -1;
               if (a < last) { for (b = first, v = a - 1; b < a; ++b) ISA(W[b]) = v; }
               if (a - first <= last - a) {
                  PUSH(dd, a, last, -3);
                  dd += 1; last = a; limit = next;
               } else {
                  if (1 < last - a) { PUSH(dd + 1, first, a, next); first = a; limit = -3; }
                  else { dd += 1; last = a; limit = next; }
               }
            } else POP();
         }
         continue;
      }

      if (last - first <= SMALL) {
         *budget -= last - first;
         tr_isort(c, dd, first, last);
         for (;;) {
            if (0 <= W[first]) {
               a = first;
               do { ISA(W[a]) = a; } while (++a < last && 0 <= W[a]);
               first = a;
            }
            if (first < last) {
               b = first; do { W[b] = ~W[b]; } while (W[++b] < 0);
               a = b + 1;
               if (ISA(W[b]) == GETC(W[b])) limit = -1;
               if (a < last) { for (b = first, v = a - 1; b < a; ++b) ISA(W[b]) = v; }
               if (1 < last - a) PUSH(dd, a, last, -4);
               dd += 1; last = a;
               if (limit == -1) break;
               *budget -= last - first;
               tr_isort(c, dd, first, last);
            } else {
               POP();
               if (limit != -4) break;
            }
         }
         continue;
      }

      if (limit-- == 0) {
         *budget -= last - first;
         tr_hsort(c, dd, first, last - first);
         for (a = last - 2, v = GETC(W[last - 1]); first <= a; --a) {
            if ((x = GETC(W[a])) == v) W[a] = ~W[a]; else v = x;
         }
         limit = -3;
         continue;
      }

      a = tr_pick(c, dd, first, last);
      SWAPW(first, a);
      v = GETC(W[first]);
      if (tr_split3(c, dd, first, last, first, v, &a, &b)) {
         next = (ISA(W[a]) == GETC(W[a])) ? -1 : ilog2(b - a);
         for (cc = first, v = a - 1; cc < a; ++cc) ISA(W[cc]) = v;
         if (b < last) { for (cc = a, v = b - 1; cc < b; ++cc) ISA(W[cc]) = v; }
         *budget -= last - first;
         if (a - first <= last - b) {
            if (last - b <= b - a) {
               if (1 < a - first) { PUSH(dd + 1, a, b, next); PUSH(dd, b, last, limit); last = a; }
               else if (1 < last - b) { PUSH(dd + 1, a, b, next); first = b; }
               else if (1 < b - a) { dd += 1; first = a; last = b; limit = next; }
               else POP();
            } else if (a - first <= b - a) {
               if (1 < a - first) { PUSH(dd, b, last, limit); PUSH(dd + 1, a, b, next); last = a; }
               else if (1 < b - a) { PUSH(dd, b, last, limit); dd += 1; first = a; last = b; limit = next; }
               else first = b;
            } else {
               if (1 < b - a) { PUSH(dd, b, last, limit); PUSH(dd, first, a, limit); dd += 1; first = a; last = b; limit = next; }
               else { PUSH(dd, b, last, limit); last = a; }
            }
         } else {
            if (a - first <= b - a) {
               if (1 < last - b) { PUSH(dd + 1, a, b, next); PUSH(dd, first, a, limit); first = b; }
               else if (1 < a - first) { PUSH(dd + 1, a, b, next); last = a; }
               else if (1 < b - a) { dd += 1; first = a; last = b; limit = next; }
               else POP();
            } else if (last - b <= b - a) {
               if (1 < last - b) { PUSH(dd, first, a, limit); PUSH(dd + 1, a, b, next); first = b; }
               else if (1 < b - a) { PUSH(dd, first, a, limit); dd += 1; first = a; last = b; limit = next; }
               else last = a;
            } else {
               if (1 < b - a) { PUSH(dd, first, a, limit); PUSH(dd, b, last, limit); dd += 1; first = a; last = b; limit = next; }
               else { PUSH(dd, first, a, limit); first = b; }
            }
         }
      } else {
         limit = (ISA(W[first]) == GETC(W[first])) ? -1 : (limit + 1);
         dd += 1;
         *budget -= last - first;
      }
   }
#undef PUSH
#undef POP
}

/* blocksort.c:1285-1309 */
static void rank_sort(tctx* c, int32_t depth)
{
   const int32_t n = c->m;
   int32_t first, last, t;
   int64_t budget;
   if (-n < W[0]) {
      first = 0;
      budget = (int64_t)(ilog2(n) * 2 / 3 + 1) * n;
      do {
         if ((t = W[first]) < 0) first -= t;
         else {
            last = ISA(t) + 1;
            tr_qsort(c, depth, first, last, &budget);
            first = last;
            if (budget <= 0) {
               W[0] = -first;
               ls_sort(c, depth);
               break;
            }
         }
      } while (first < n);
   }
}

/* -------------------------------------------------------------------- driver -- */
/* blocksort.c:1316-1401.  work: n + 256 ints.  bstar: 65536 ints.  T[n] must equal T[0].
 * Returns m (number of B* suffixes); ranks of the B* suffixes (by B* index, ascending position)
 * are left in work[m .. 2m). */
int32_t orc_bstar_ranks(const uint8_t* T, int32_t n, int32_t* work, int32_t* bstar)
{
   tctx ctx, *c = &ctx;
   int32_t i, j, m, t, c0, c1, buf, bufsize;
   int flag;
   c->T = T; c->w = work; c->n = n;
   memset(bstar, 0, sizeof(int32_t) * 65536);
#define BSTAR(x, y) bstar[((x) << 8) + (y)]

   /* classify from the end; only the B* counts matter here (:1329-1347) */
   for (i = 1, c1 = T[0]; i < n && T[i] == c1; ++i) { }
   flag = c1 <= T[i];
   i = n - 1; m = n;
   if ((c0 = T[i]) < c1 || (c0 == c1 && flag)) {
      if (!flag) { ++BSTAR(c0, c1); work[--m] = i; }
      for (; c1 = c0, 0 <= --i && (c0 = T[i]) <= c1;) { }
   }
   for (; 0 <= i;) {
      do { c1 = c0; } while (0 <= --i && (c0 = T[i]) >= c1);
      if (0 <= i) {
         ++BSTAR(c0, c1);
         work[--m] = i;
         for (; c1 = c0, 0 <= --i && (c0 = T[i]) <= c1;) { }
      }
   }
   m = n - m;
   c->m = m; c->pa = n - m; c->isa = m;
   if (m == 0) return 0;

   /* bucket ends (:1351-1360), then the fill: ascending B* index inside a bucket, the block's last
    * B* suffix placed last, i.e. at the front of its bucket (:1362-1368) */
   for (c0 = 0, j = 0; c0 < 256; ++c0)
      for (c1 = c0 + 1; c1 < 256; ++c1) { j += BSTAR(c0, c1); BSTAR(c0, c1) = j; }
   for (i = m - 2; 0 <= i; --i) { t = PA(i); W[--BSTAR(T[t], T[t + 1])] = i; }
   t = PA(m - 1); W[--BSTAR(T[t], T[t + 1])] = m - 1;

   /* substring sort per bucket, highest bucket first (:1375-1385) */
   buf = m; bufsize = n - 2 * m;
   if (bufsize <= 256) { buf = n; bufsize = 256; }
   for (c0 = 255, j = m; 0 < j; --c0) {
      for (c1 = 255; c0 < c1; j = i, --c1) {
         i = BSTAR(c0, c1);
         if (1 < j - i) sort_bucket(c, i, j, buf, bufsize, W[i] == m - 1);
      }
   }

   /* ranks from the marked order (:1387-1398) */
   for (i = m - 1; 0 <= i; --i) {
      if (0 <= (t = W[i])) {
         j = i;
         do { ISA(t) = i; } while (0 <= --i && 0 <= (t = W[i]));
         W[i + 1] = i - j;
         if (i <= 0) break;
      }
      j = i;
      do { ISA(W[i] = ~t) = j; } while ((t = W[--i]) < 0);
      ISA(t) = j;
   }

   rank_sort(c, 1);   /* :1401 */
   return m;
#undef BSTAR
}

/* g of origPtr = lo + g on blk = u^q, q >= 2 (see the header of this file). */
int32_t orc_tie_offset(const uint8_t* blk, int32_t n, int32_t q)
{
   if (q < 2 || n % q) return -1;
   uint8_t* T = malloc((size_t)n + 2);
   int32_t* work = malloc(sizeof(int32_t) * ((size_t)n + 256));
   int32_t* bstar = malloc(sizeof(int32_t) * 65536);
   memcpy(T, blk, (size_t)n);
   T[n] = blk[0];
   int32_t m = orc_bstar_ranks(T, n, work, bstar);
   int32_t g = 0;
   if (m > 0) {
      const int32_t per_unit = m / q;           /* B* suffixes per copy of u */
      const int32_t* isa = work + m;
      int32_t lo = isa[0];
      for (int32_t k = 1; k < q; k++) if (isa[k * per_unit] < lo) lo = isa[k * per_unit];
      g = isa[0] - lo;
   }
   free(T); free(work); free(bstar);
   return g;
}
