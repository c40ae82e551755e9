/*
 * bzlib_decode.c -- host-side .bz2 decoder behind libbz2's decompression API.
 *
 * Decompression is outside the accelerated path; it is
 * provided so that the library exports the complete libbz2 interface (libbz2.def:4-27) and
 * existing clients -- including the reference's own bzip2.c -- relink unchanged.  Written from
 * the stream format, not from the reference's decompress.c: the parser
 * is "retry from the start of the block" instead of a resumable switch, and output is
 * generated lazily from the inverse-BWT chain.
 *
 *   BZ2_bzDecompressInit / BZ2_bzDecompress / BZ2_bzDecompressEnd   (API of bzlib.c:482-940)
 *   BZ2_bzBuffToBuffDecompress                                      (bzlib.c:1360-1415)
 * Not supported: the `randomised` block flag of pre-0.9.5 streams (BZ_DATA_ERROR); `small` is
 * accepted and ignored.  Input is consumed exactly up to the end of the stream, so
 * strm->next_in/avail_in describe the unused tail after BZ_STREAM_END as in the reference.
 */
#include "../../include/bzlib.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { D_HEADER = 0, D_ITEM, D_OUTPUT, D_DONE, D_ERROR };
enum { P_OK = 0, P_MORE = 1, P_BAD = 2, P_MAGIC = 3, P_END = 4 };

typedef struct {
   bz_stream* strm;
   int phase;
   int level;                      /* blockSize100k from the header                       */
   /* unparsed compressed bytes carried between calls (start of the current item)          */
   unsigned char* ibuf; size_t ilen, icap;
   unsigned ibit;                  /* bits of ibuf[0] already consumed (0..7)              */
   uint64_t scan;                  /* next bit position of ibuf to test for the following magic */
   /* decoded block                                                                          */
   uint32_t* tt; unsigned char* ll; int nblock, orig_ptr;
   uint32_t block_crc_stored, block_crc, combined_crc;
   /* lazy output state: inverse BWT chain + RLE1 expansion                                  */
   uint32_t tpos; int used; int run_ch, run_len, pending_ch, pending_rep;
   int err;
} dstate;

static uint32_t crc_tab[256];
static int crc_ready;
static void crc_init(void)
{
   uint32_t b, r; int k;
   if (crc_ready) return;
   for (b = 0; b < 256; b++) {
      r = b << 24;
      for (k = 0; k < 8; k++) r = (r & 0x80000000u) ? (r << 1) ^ 0x04C11DB7u : (r << 1);
      crc_tab[b] = r;
   }
   crc_ready = 1;
}

/* ---- bit reader over a byte range; sets `short_` instead of reading past the end ---------- */
typedef struct { const unsigned char* p; size_t n; uint64_t pos; int short_; } bitr;
static uint32_t get_bits(bitr* r, int nb)
{
   uint32_t v = 0;
   if (r->pos + (uint64_t)nb > (uint64_t)r->n * 8) { r->short_ = 1; r->pos += (uint64_t)nb; return 0; }
   while (nb > 0) {
      const unsigned byte = r->p[r->pos >> 3];
      const int avail = 8 - (int)(r->pos & 7);
      const int take = nb < avail ? nb : avail;
      v = (v << take) | ((byte >> (avail - take)) & ((1u << take) - 1u));
      r->pos += (uint64_t)take; nb -= take;
   }
   return v;
}

/* Parse one block (after its 48-bit magic) into s->ll / s->tt.  P_MORE if data ran out. */
static int parse_block(dstate* s, bitr* r)
{
   unsigned char seq2unseq[256], sel[32768], order[6], mtf[256];
   unsigned char lens[6][258];
   int32_t limit[6][22], base[6][22], perm[6][258], minlen[6];
   int32_t cftab[257];
   int n_in_use = 0, alpha, ngroups, nsel, i, j, t, eob, nblock = 0;
   const int nmax = 100000 * s->level;
   uint32_t used16;

   s->block_crc_stored = get_bits(r, 32);
   if (get_bits(r, 1)) { if (r->short_) return P_MORE; return P_BAD; }       /* randomised blocks unsupported */
   s->orig_ptr = (int)get_bits(r, 24);
   used16 = get_bits(r, 16);
   if (r->short_) return P_MORE;
   for (i = 0; i < 16; i++) if (used16 & (0x8000u >> i)) {
      const uint32_t m = get_bits(r, 16);
      for (j = 0; j < 16; j++) if (m & (0x8000u >> j)) seq2unseq[n_in_use++] = (unsigned char)(i * 16 + j);
   }
   if (r->short_) return P_MORE;
   if (n_in_use == 0) return P_BAD;
   alpha = n_in_use + 2;
   ngroups = (int)get_bits(r, 3);
   nsel = (int)get_bits(r, 15);
   if (r->short_) return P_MORE;
   if (ngroups < 2 || ngroups > 6 || nsel < 1) return P_BAD;
   for (i = 0; i < 6; i++) order[i] = (unsigned char)i;
   for (i = 0; i < nsel; i++) {
      unsigned char v;
      j = 0;
      while (get_bits(r, 1)) { j++; if (j >= ngroups) { if (r->short_) return P_MORE; return P_BAD; } }
      if (r->short_) return P_MORE;
      v = order[j];
      for (; j > 0; j--) order[j] = order[j - 1];
      order[0] = v;
      sel[i] = v;
   }
   for (t = 0; t < ngroups; t++) {
      int cur = (int)get_bits(r, 5);
      for (i = 0; i < alpha; i++) {
         for (;;) {
            if (r->short_) return P_MORE;
            if (cur < 1 || cur > 20) return P_BAD;
            if (!get_bits(r, 1)) break;
            cur += get_bits(r, 1) ? -1 : 1;
         }
         lens[t][i] = (unsigned char)cur;
      }
   }
   if (r->short_) return P_MORE;
   /* canonical decode tables */
   for (t = 0; t < ngroups; t++) {
      int mn = 32, mx = 0, pp = 0, vec = 0;
      for (i = 0; i < alpha; i++) { if (lens[t][i] > mx) mx = lens[t][i]; if (lens[t][i] < mn) mn = lens[t][i]; }
      for (i = mn; i <= mx; i++) for (j = 0; j < alpha; j++) if (lens[t][j] == i) perm[t][pp++] = j;
      for (i = 0; i < 22; i++) { base[t][i] = 0; limit[t][i] = 0; }
      for (i = 0; i < alpha; i++) base[t][lens[t][i] + 1]++;
      for (i = 1; i < 22; i++) base[t][i] += base[t][i - 1];
      for (i = mn; i <= mx; i++) {
         vec += base[t][i + 1] - base[t][i];
         limit[t][i] = vec - 1;
         vec <<= 1;
      }
      for (i = mn + 1; i <= mx; i++) base[t][i] = ((limit[t][i - 1] + 1) << 1) - base[t][i];
      minlen[t] = mn;
      limit[t][mx + 1] = 0x7fffffff;     /* sentinel: any code stops at mx + 1 at the latest */
   }
   /* symbols */
   eob = n_in_use + 1;
   for (i = 0; i < 256; i++) mtf[i] = (unsigned char)i;
   memset(cftab, 0, sizeof cftab);
   {
      int group = -1, left = 0, gt = 0;
      int64_t run = 0; int run_shift = 0;
      for (;;) {
         int zn, zvec, sym;
         if (left == 0) { group++; if (group >= nsel) return r->short_ ? P_MORE : P_BAD; gt = sel[group]; left = 50; }
         left--;
         zn = minlen[gt];
         zvec = (int)get_bits(r, zn);
         while (zvec > limit[gt][zn]) {
            zn++;
            if (zn > 20) return r->short_ ? P_MORE : P_BAD;
            zvec = (zvec << 1) | (int)get_bits(r, 1);
         }
         if (r->short_) return P_MORE;
         if (zvec - base[gt][zn] < 0 || zvec - base[gt][zn] >= alpha) return P_BAD;
         sym = perm[gt][zvec - base[gt][zn]];
         if (sym <= 1) {                               /* RUNA / RUNB: bijective base 2 */
            run += (int64_t)(sym + 1) << run_shift;
            run_shift++;
            if (run_shift > 21) return P_BAD;
            continue;
         }
         if (run) {
            const unsigned char ch = seq2unseq[mtf[0]];
            if (nblock + run > nmax) return P_BAD;
            cftab[ch + 1] += (int32_t)run;
            memset(s->ll + nblock, ch, (size_t)run);
            nblock += (int)run; run = 0; run_shift = 0;
         }
         if (sym == eob) break;
         {
            const int p = sym - 1;
            const unsigned char v = mtf[p];
            if (p >= n_in_use) return P_BAD;
            memmove(mtf + 1, mtf, (size_t)p);
            mtf[0] = v;
            if (nblock >= nmax) return P_BAD;
            s->ll[nblock++] = seq2unseq[v];
            cftab[seq2unseq[v] + 1]++;
         }
      }
   }
   if (s->orig_ptr < 0 || s->orig_ptr >= nblock) return P_BAD;
   for (i = 1; i <= 256; i++) cftab[i] += cftab[i - 1];
   for (i = 0; i < nblock; i++) { const unsigned char ch = s->ll[i]; s->tt[cftab[ch]++] = (uint32_t)i; }
   s->nblock = nblock;
   return P_OK;
}

/* ---- lazy output: walk the inverse BWT, undo RLE1, update the CRC -------------------------- */
static void start_output(dstate* s)
{
   s->tpos = s->tt[s->orig_ptr];
   s->used = 0; s->run_ch = -1; s->run_len = 0; s->pending_rep = 0;
   s->block_crc = 0xFFFFFFFFu;
}
static int next_byte(dstate* s)
{
   const int ch = s->ll[s->tpos];
   s->tpos = s->tt[s->tpos];
   s->used++;
   return ch;
}
/* returns 1 when the block is exhausted */
static int produce(dstate* s)
{
   bz_stream* z = s->strm;
   for (;;) {
      while (s->pending_rep > 0) {
         if (z->avail_out == 0) return 0;
         *z->next_out++ = (char)s->pending_ch; z->avail_out--;
         if (++z->total_out_lo32 == 0) z->total_out_hi32++;
         s->block_crc = (s->block_crc << 8) ^ crc_tab[(s->block_crc >> 24) ^ (unsigned)s->pending_ch];
         s->pending_rep--;
      }
      if (s->used >= s->nblock) return 1;
      {
         const int ch = next_byte(s);
         if (s->run_len == 4) {                      /* this byte is a repeat count */
            s->pending_ch = s->run_ch; s->pending_rep = ch; s->run_len = 0; s->run_ch = -1;
            continue;
         }
         if (ch == s->run_ch) s->run_len++; else { s->run_ch = ch; s->run_len = 1; }
         s->pending_ch = ch; s->pending_rep = 1;
      }
   }
}

/* ---------------------------------------------------------------------------------------------- */
static void* dflt_alloc(void* o, int a, int b) { (void)o; return malloc((size_t)a * (size_t)b); }
static void dflt_free(void* o, void* p) { (void)o; if (p) free(p); }

int BZ2_bzDecompressInit(bz_stream* strm, int verbosity, int small)
{
   dstate* s;
   if (strm == NULL || (small != 0 && small != 1) || verbosity < 0 || verbosity > 4) return BZ_PARAM_ERROR;
   if (strm->bzalloc == NULL) strm->bzalloc = dflt_alloc;
   if (strm->bzfree == NULL) strm->bzfree = dflt_free;
   s = (dstate*)strm->bzalloc(strm->opaque, (int)sizeof(dstate), 1);
   if (!s) return BZ_MEM_ERROR;
   memset(s, 0, sizeof *s);
   s->strm = strm;
   s->phase = D_HEADER;
   crc_init();
   strm->state = s;
   strm->total_in_lo32 = strm->total_in_hi32 = 0;
   strm->total_out_lo32 = strm->total_out_hi32 = 0;
   return BZ_OK;
}

int BZ2_bzDecompressEnd(bz_stream* strm)
{
   dstate* s;
   if (strm == NULL || (s = (dstate*)strm->state) == NULL || s->strm != strm) return BZ_PARAM_ERROR;
   free(s->ibuf); free(s->tt); free(s->ll);
   strm->bzfree(strm->opaque, s);
   strm->state = NULL;
   return BZ_OK;
}

/* ---- input: consumed one byte at a time, exactly up to the end of the stream -------------------
 * Items (blocks, the trailer) are contiguous in the bit stream and every item starts with a
 * 48-bit magic, so the end of a block is found by watching for the next magic; only then is
 * the block parsed, once, from the bytes gathered in ibuf.  A magic look-alike inside coded data
 * is recognised because the parse does not end exactly there. */
#define MAGIC_BLOCK 0x314159265359ULL
#define MAGIC_END   0x177245385090ULL

static int push_byte(dstate* s, unsigned char b)
{
   if (s->ilen + 1 > s->icap) {
      size_t nc = s->icap ? s->icap * 2 : 65536;
      unsigned char* nb = (unsigned char*)realloc(s->ibuf, nc);
      if (!nb) return 0;
      s->ibuf = nb; s->icap = nc;
   }
   s->ibuf[s->ilen++] = b;
   return 1;
}

static int take_byte(dstate* s)
{
   bz_stream* z = s->strm;
   if (z->avail_in == 0) return 0;
   if (!push_byte(s, (unsigned char)*z->next_in)) { s->phase = D_ERROR; s->err = BZ_MEM_ERROR; return 0; }
   z->next_in++; z->avail_in--;
   if (++z->total_in_lo32 == 0) z->total_in_hi32++;
   return 1;
}

/* bits of the current item available in ibuf */
static uint64_t have_bits(const dstate* s) { return (uint64_t)s->ilen * 8 - s->ibit; }

/* 48 bits starting at absolute bit position `at` of ibuf */
static uint64_t peek48(const dstate* s, uint64_t at)
{
   bitr r;
   uint64_t hi, lo;
   r.p = s->ibuf; r.n = s->ilen; r.pos = at; r.short_ = 0;
   hi = get_bits(&r, 24); lo = get_bits(&r, 24);
   return (hi << 24) | lo;
}

/* First bit position p >= s->scan of ibuf at which a block or end magic starts, testing only
 * positions whose 48 bits have arrived; advances s->scan past what was tested. */
static int find_next_magic(dstate* s, uint64_t* at_out)
{
   size_t byte = (size_t)((s->scan + 47) >> 3);        /* byte holding the last bit of the first candidate */
   for (; byte < s->ilen; byte++) {
      uint64_t w = 0;
      int i, k;
      for (i = 7; i >= 0; i--) w = (w << 8) | ((size_t)i <= byte ? s->ibuf[byte - (size_t)i] : 0);
      for (k = 0; k < 8; k++) {
         const uint64_t end = (uint64_t)byte * 8 + (uint64_t)k + 1;
         uint64_t v;
         if (end < 48 || end - 48 < s->scan) continue;
         v = (w >> (7 - k)) & 0xFFFFFFFFFFFFULL;
         if (v == MAGIC_BLOCK || v == MAGIC_END) { *at_out = end - 48; s->scan = end - 48; return 1; }
      }
      s->scan = (uint64_t)byte * 8 + 8 - 47;
   }
   return 0;
}

/* the next item starts at absolute bit `bits_abs`: drop whole bytes from the front of ibuf */
static void advance_item(dstate* s, uint64_t bits_abs)
{
   const size_t full = (size_t)(bits_abs >> 3);
   memmove(s->ibuf, s->ibuf + full, s->ilen - full);
   s->ilen -= full;
   s->ibit = (unsigned)(bits_abs & 7);
   s->scan = (uint64_t)s->ibit + 48;       /* first position where the FOLLOWING magic may start */
}

int BZ2_bzDecompress(bz_stream* strm)
{
   dstate* s;
   if (strm == NULL || (s = (dstate*)strm->state) == NULL || s->strm != strm) return BZ_PARAM_ERROR;
   for (;;) {
      if (s->phase == D_ERROR) return s->err;
      if (s->phase == D_DONE) return BZ_SEQUENCE_ERROR;
      if (s->phase == D_OUTPUT) {
         if (!produce(s)) return BZ_OK;                                   /* output space exhausted */
         s->block_crc = ~s->block_crc;
         if (s->block_crc != s->block_crc_stored) { s->phase = D_ERROR; return s->err = BZ_DATA_ERROR; }
         s->combined_crc = ((s->combined_crc << 1) | (s->combined_crc >> 31)) ^ s->block_crc;
         s->phase = D_ITEM;
      }
      if (s->phase == D_HEADER) {
         while (s->ilen < 4) {
            if (!take_byte(s)) return s->phase == D_ERROR ? s->err : BZ_OK;
            /* reject a wrong magic as early as the bytes allow */
            if ((s->ilen == 1 && s->ibuf[0] != 'B') || (s->ilen == 2 && s->ibuf[1] != 'Z') || (s->ilen == 3 && s->ibuf[2] != 'h'))
               { s->phase = D_ERROR; return s->err = BZ_DATA_ERROR_MAGIC; }
         }
         if (s->ibuf[3] < '1' || s->ibuf[3] > '9') { s->phase = D_ERROR; return s->err = BZ_DATA_ERROR_MAGIC; }
         s->level = s->ibuf[3] - '0';
         s->tt = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(100000 * s->level));
         s->ll = (unsigned char*)malloc((size_t)(100000 * s->level));
         if (!s->tt || !s->ll) { s->phase = D_ERROR; return s->err = BZ_MEM_ERROR; }
         s->ilen = 0; s->ibit = 0; s->scan = 48;
         s->phase = D_ITEM;
      }
      /* D_ITEM: the item's own magic first */
      while (have_bits(s) < 48) {
         uint64_t hb;
         if (!take_byte(s)) return s->phase == D_ERROR ? s->err : BZ_OK;
         hb = have_bits(s);
         if (hb > 0 && hb < 48) {          /* a wrong magic is an error as soon as one byte of it is wrong */
            bitr r;
            uint64_t got;
            r.p = s->ibuf; r.n = s->ilen; r.pos = s->ibit; r.short_ = 0;
            got = hb > 24 ? ((uint64_t)get_bits(&r, 24) << (hb - 24)) | get_bits(&r, (int)(hb - 24)) : get_bits(&r, (int)hb);
            if (got != (MAGIC_BLOCK >> (48 - hb)) && got != (MAGIC_END >> (48 - hb))) { s->phase = D_ERROR; return s->err = BZ_DATA_ERROR; }
         }
      }
      {
         const uint64_t magic = peek48(s, s->ibit);
         if (magic == MAGIC_END) {
            bitr r;
            uint32_t stored;
            while (have_bits(s) < 80) if (!take_byte(s)) return s->phase == D_ERROR ? s->err : BZ_OK;
            r.p = s->ibuf; r.n = s->ilen; r.pos = (uint64_t)s->ibit + 48; r.short_ = 0;
            stored = get_bits(&r, 32);
            s->phase = D_DONE;
            s->ilen = 0;
            if (stored != s->combined_crc) { s->phase = D_ERROR; return s->err = BZ_DATA_ERROR; }
            return BZ_STREAM_END;
         }
         if (magic != MAGIC_BLOCK) { s->phase = D_ERROR; return s->err = BZ_DATA_ERROR; }
      }
      /* a block: gather bytes until the next magic shows up, then parse once */
      for (;;) {
         int found = 0;
         uint64_t at = 0;
         found = find_next_magic(s, &at);
         if (!found) {
            /* a legal block of a foreign encoder may spend up to 20 bits on every symbol (the format's code-length limit):
             * 2.5 bytes per block byte plus selectors and tables */
            if ((uint64_t)s->ilen > (uint64_t)s->level * 250000 + 100000) { s->phase = D_ERROR; return s->err = BZ_DATA_ERROR; }
            if (!take_byte(s)) return s->phase == D_ERROR ? s->err : BZ_OK;
            continue;
         }
         {
            bitr r;
            int rc;
            r.p = s->ibuf; r.n = (size_t)((at + 7) >> 3); r.pos = (uint64_t)s->ibit + 48; r.short_ = 0;
            rc = parse_block(s, &r);
            if (rc == P_OK && !r.short_ && r.pos == at) {
               advance_item(s, at);
               s->phase = D_OUTPUT;
               start_output(s);
               break;
            }
            if (rc == P_OK && !r.short_ && r.pos < at) { s->phase = D_ERROR; return s->err = BZ_DATA_ERROR; }
            if (rc == P_BAD && !r.short_) { s->phase = D_ERROR; return s->err = BZ_DATA_ERROR; }
            s->scan = at + 1;            /* a magic look-alike inside the block: keep looking */
         }
      }
   }
}

int BZ2_bzBuffToBuffDecompress(char* dest, unsigned int* destLen, char* source, unsigned int sourceLen, int small, int verbosity)
{
   bz_stream strm;
   int ret;
   if (dest == NULL || destLen == NULL || source == NULL || (small != 0 && small != 1) || verbosity < 0 || verbosity > 4)
      return BZ_PARAM_ERROR;
   memset(&strm, 0, sizeof strm);
   ret = BZ2_bzDecompressInit(&strm, verbosity, small);
   if (ret != BZ_OK) return ret;
   strm.next_in = source; strm.avail_in = sourceLen;
   strm.next_out = dest; strm.avail_out = *destLen;
   ret = BZ2_bzDecompress(&strm);
   if (ret == BZ_OK) {                       /* stopped for lack of input or of room: say which */
      BZ2_bzDecompressEnd(&strm);
      return strm.avail_out > 0 ? BZ_UNEXPECTED_EOF : BZ_OUTBUFF_FULL;
   }
   if (ret != BZ_STREAM_END) { BZ2_bzDecompressEnd(&strm); return ret; }
   *destLen -= strm.avail_out;
   BZ2_bzDecompressEnd(&strm);
   return BZ_OK;
}
