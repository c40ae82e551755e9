/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Builds the UNMODIFIED reference (aeb1787/bzip2) as one translation unit by
 * #include-ing its sources from where they lie (-I/root/reference, see
 * oracle/Makefile) and appends a few entry points that expose the reference's
 * per-stage functions (several are `static` in the reference) so that tests
 * can compare stage by stage.  No reference source text is copied here.
 *
 * Output: oracle/_ref/libbz2_ref.so  (git-ignored; travels with gpurun).
 */
#include "bzlib.c"
#include "compress.c"
#include "huffman.c"
#include "crctable.c"
#include "randtable.c"
#include "decompress.c"
#include "blocksort.c"

#include <stdint.h>
#include <string.h>
#include <stdlib.h>

/* One record per emitted block, filled by ref_trace(). */
typedef struct {
   uint64_t in_end;      /* input bytes consumed when the block was closed      */
   int32_t  nblock;      /* post-RLE1 length                                    */
   uint32_t block_crc;   /* finalised block CRC                                 */
   uint32_t comb_crc;    /* combined CRC after folding this block               */
   int32_t  orig_ptr;
   int32_t  n_mtf;
   int32_t  n_in_use;
   int32_t  num_z;       /* bytes appended to the stream by this block          */
} ref_block_rec;

/*
 * Run the reference compressor over `in` exactly as BZ2_bzCompress(BZ_FINISH)
 * would (bzlib.c:361-396), recording per-block facts.  If want_block >= 0 the
 * post-RLE1 bytes of that block are copied to blk_out (cap 900000+16).
 * Returns the number of blocks, or <0 on error.  Stream bytes go to out/out_len.
 */
int ref_trace(const uint8_t* in, uint64_t n, int level,
              ref_block_rec* recs, int max_recs,
              int want_block, uint8_t* blk_out,
              uint8_t* out, uint64_t out_cap, uint64_t* out_len)
{
   bz_stream strm;
   EState* s;
   int nb = 0;
   uint64_t pos = 0, olen = 0;
   memset(&strm, 0, sizeof strm);
   if (BZ2_bzCompressInit(&strm, level, 0, 0) != BZ_OK) return -1;
   s = (EState*)strm.state;
   s->mode = BZ_M_FINISHING;
   for (;;) {
      uint32_t chunk = (n - pos > 0x40000000u) ? 0x40000000u : (uint32_t)(n - pos);
      int final_chunk = (pos + chunk == n);
      strm.next_in = (char*)(in + pos);
      strm.avail_in = chunk;
      /* only the final chunk may flush the pending run */
      s->mode = final_chunk ? BZ_M_FINISHING : BZ_M_RUNNING;
      copy_input_until_stop(s);
      pos = (uint64_t)((uint8_t*)strm.next_in - in);
      if (s->nblock >= s->nblockMAX || (final_chunk && strm.avail_in == 0)) {
         int is_last = (final_chunk && strm.avail_in == 0);
         if (s->mode == BZ_M_RUNNING) s->avail_in_expect = 1;
         if (nb == want_block && blk_out) memcpy(blk_out, s->block, (size_t)s->nblock);
         {
            int32_t nblock = s->nblock;
            BZ2_compressBlock(s, (Bool)is_last);
            if (nblock > 0 || 1) {
               if (nb < max_recs && recs) {
                  recs[nb].in_end = pos;
                  recs[nb].nblock = nblock;
                  recs[nb].block_crc = s->blockCRC;
                  recs[nb].comb_crc = s->combinedCRC;
                  recs[nb].orig_ptr = nblock > 0 ? s->origPtr : 0;
                  recs[nb].n_mtf = nblock > 0 ? s->nMTF : 0;
                  recs[nb].n_in_use = nblock > 0 ? s->nInUse : 0;
                  recs[nb].num_z = s->numZ;
               }
               nb++;
            }
         }
         if (out) {
            if (olen + (uint64_t)s->numZ > out_cap) { BZ2_bzCompressEnd(&strm); return -2; }
            memcpy(out + olen, s->zbits, (size_t)s->numZ);
         }
         olen += (uint64_t)s->numZ;
         if (is_last) break;
         prepare_new_block(s);
      }
   }
   if (out_len) *out_len = olen;
   BZ2_bzCompressEnd(&strm);
   return nb;
}

/* BZ2_blockSort on one post-RLE1 block (blocksort.c:1534). bwt_out: n+1 u32. */
int ref_bwt(const uint8_t* blk, int n, uint32_t* bwt_out, int* orig_ptr)
{
   bz_stream strm;
   EState* s;
   if (n < 1 || n > 900000 - 1) return -1;
   memset(&strm, 0, sizeof strm);
   if (BZ2_bzCompressInit(&strm, 9, 0, 0) != BZ_OK) return -1;
   s = (EState*)strm.state;
   memcpy(s->block, blk, (size_t)n);
   s->nblock = n;
   BZ2_blockSort(s);
   memcpy(bwt_out, s->ptr, (size_t)n * sizeof(uint32_t));
   *orig_ptr = s->origPtr;
   BZ2_bzCompressEnd(&strm);
   return 0;
}

/* generateMTFValues (compress.c:93) on BWT bytes given as one byte each. */
int ref_mtf(const uint8_t* bwt, int n, const uint8_t* in_use,
            uint16_t* mtfv_out, int* n_mtf, int32_t* mtf_freq /*258*/, int* n_in_use)
{
   bz_stream strm;
   EState* s;
   int i;
   if (n < 1 || n > 900000 - 1) return -1;
   memset(&strm, 0, sizeof strm);
   if (BZ2_bzCompressInit(&strm, 9, 0, 0) != BZ_OK) return -1;
   s = (EState*)strm.state;
   for (i = 0; i < n; i++) s->ptr[i] = bwt[i];
   for (i = 0; i < 256; i++) s->inUse[i] = in_use[i] ? True : False;
   s->nblock = n;
   generateMTFValues(s);
   memcpy(mtfv_out, s->mtfv, (size_t)s->nMTF * sizeof(uint16_t));
   *n_mtf = s->nMTF;
   *n_in_use = s->nInUse;
   memcpy(mtf_freq, s->mtfFreq, sizeof(int32_t) * BZ_MAX_ALPHA_SIZE);
   BZ2_bzCompressEnd(&strm);
   return 0;
}

/*
 * sendMTFValues (compress.c:250) in isolation: emits only the table/selector/
 * payload section for the given MTF symbols, starting from an empty bit buffer.
 * Returns number of bytes written to out (bit buffer flushed with zero padding),
 * and the exact bit count in *nbits.
 */
int ref_send_mtf(const uint16_t* mtfv, int n_mtf, const uint8_t* in_use,
                 uint8_t* out, int out_cap, int64_t* nbits)
{
   bz_stream strm;
   EState* s;
   int i, ninuse = 0, nz;
   memset(&strm, 0, sizeof strm);
   if (BZ2_bzCompressInit(&strm, 9, 0, 0) != BZ_OK) return -1;
   s = (EState*)strm.state;
   memcpy(s->mtfv, mtfv, (size_t)n_mtf * sizeof(uint16_t));
   for (i = 0; i < 256; i++) { s->inUse[i] = in_use[i] ? True : False; ninuse += in_use[i] ? 1 : 0; }
   s->nInUse = ninuse;
   s->nMTF = n_mtf;
   for (i = 0; i < BZ_MAX_ALPHA_SIZE; i++) s->mtfFreq[i] = 0;
   for (i = 0; i < n_mtf; i++) s->mtfFreq[mtfv[i]]++;
   s->nblock = 1;
   BZ2_bsInitWrite(s);
   s->numZ = 0;
   sendMTFValues(s);
   *nbits = (int64_t)s->numZ * 8 + (64 - s->bsLive);
   bsFinishWrite(s);
   nz = s->numZ;
   if (nz > out_cap) { BZ2_bzCompressEnd(&strm); return -2; }
   memcpy(out, s->zbits, (size_t)nz);
   BZ2_bzCompressEnd(&strm);
   return nz;
}

void ref_make_code_lengths(int32_t* len, int32_t* freq, int alpha, int max_len)
{
   BZ2_hbMakeCodeLengths(len, freq, alpha, max_len);
}

void ref_assign_codes(int32_t* code, int32_t* len, int min_len, int max_len, int alpha)
{
   BZ2_hbAssignCodes(code, len, min_len, max_len, alpha);
}

uint32_t ref_crc(const uint8_t* p, uint64_t n)
{
   uint32_t c = 0xffffffffu;
   uint64_t i;
   for (i = 0; i < n; i++) { BZ_UPDATE_CRC(c, p[i]); }
   return ~c;
}
