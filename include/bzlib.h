/*
 * bzlib.h -- libbz2-compatible public interface of the B200 bzip2 compressor.
 *
 * Drop-in for the reference's bzlib.h (aeb1787/bzip2 bzlib.h:29-66 constants
 * and bz_stream, :100-128 streaming calls, :204-221 one-shot calls, :134-199
 * stdio calls, :238-271 zlib-flavoured calls).  Names, argument meaning, return
 * codes and the bz_stream layout are the reference's, so existing callers
 * relink unchanged; the work behind them runs on the GPU (see bz2_b200.h).
 * Decompression entry points are provided by a small host decoder (used for
 * round-trip checks); they are outside the accelerated path.  It does not decode
 * the "randomised" blocks that only pre-0.9.5 encoders wrote (BZ_DATA_ERROR).
 *
 * Memory: the stream state and the queue of compressed bytes waiting for the caller
 * are obtained through bzalloc / bzfree (bzlib.c:104-115, :164-175).  Device memory
 * and the pinned staging buffers of the engine cannot be: they come from the CUDA
 * driver (cudaMalloc / cudaMallocHost).
 */
#ifndef BZ2_B200_BZLIB_H
#define BZ2_B200_BZLIB_H

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* actions for BZ2_bzCompress and return codes: preprocessor constants, as in the reference (bzlib.h:29-46), so that
 * clients testing them with #ifdef / #if compile unchanged */
#define BZ_RUN               0
#define BZ_FLUSH             1
#define BZ_FINISH            2

#define BZ_OK                0
#define BZ_RUN_OK            1
#define BZ_FLUSH_OK          2
#define BZ_FINISH_OK         3
#define BZ_STREAM_END        4
#define BZ_SEQUENCE_ERROR    (-1)
#define BZ_PARAM_ERROR       (-2)
#define BZ_MEM_ERROR         (-3)
#define BZ_DATA_ERROR        (-4)
#define BZ_DATA_ERROR_MAGIC  (-5)
#define BZ_IO_ERROR          (-6)
#define BZ_UNEXPECTED_EOF    (-7)
#define BZ_OUTBUFF_FULL      (-8)
#define BZ_CONFIG_ERROR      (-9)

#define BZ_MAX_UNUSED 5000

/* Caller-owned stream cursor; field order and types are ABI (reference bzlib.h:48-66). */
typedef struct {
   char*        next_in;
   unsigned int avail_in;
   unsigned int total_in_lo32;
   unsigned int total_in_hi32;

   char*        next_out;
   unsigned int avail_out;
   unsigned int total_out_lo32;
   unsigned int total_out_hi32;

   void*        state;

   void* (*bzalloc)(void*, int, int);
   void  (*bzfree)(void*, void*);
   void*        opaque;
} bz_stream;

typedef void BZFILE;

/* streaming compression (reference bzlib.c:144-207, :400-454, :458-474) */
int BZ2_bzCompressInit(bz_stream* strm, int blockSize100k, int verbosity, int workFactor);
int BZ2_bzCompress(bz_stream* strm, int action);
int BZ2_bzCompressEnd(bz_stream* strm);

/* one-shot compression (reference bzlib.c:1309-1357) */
int BZ2_bzBuffToBuffCompress(char* dest, unsigned int* destLen, char* source, unsigned int sourceLen,
                             int blockSize100k, int verbosity, int workFactor);

/* stdio write side (reference bzlib.c:978-1146) */
BZFILE* BZ2_bzWriteOpen(int* bzerror, FILE* f, int blockSize100k, int verbosity, int workFactor);
void    BZ2_bzWrite(int* bzerror, BZFILE* b, void* buf, int len);
void    BZ2_bzWriteClose(int* bzerror, BZFILE* b, int abandon, unsigned int* nbytes_in, unsigned int* nbytes_out);
void    BZ2_bzWriteClose64(int* bzerror, BZFILE* b, int abandon,
                           unsigned int* nbytes_in_lo32, unsigned int* nbytes_in_hi32,
                           unsigned int* nbytes_out_lo32, unsigned int* nbytes_out_hi32);

const char* BZ2_bzlibVersion(void);

/* streaming and one-shot decompression (reference bzlib.c:488-530, :801-900, :1361-1409).
 * Host decoder (csrc/bzlib_decode.c); `small` is accepted and ignored. */
int BZ2_bzDecompressInit(bz_stream* strm, int verbosity, int small);
int BZ2_bzDecompress(bz_stream* strm);
int BZ2_bzDecompressEnd(bz_stream* strm);
int BZ2_bzBuffToBuffDecompress(char* dest, unsigned int* destLen, char* source, unsigned int sourceLen,
                               int small, int verbosity);

/* stdio read side (reference bzlib.c:1150-1301) */
BZFILE* BZ2_bzReadOpen(int* bzerror, FILE* f, int verbosity, int small, void* unused, int nUnused);
void    BZ2_bzReadClose(int* bzerror, BZFILE* b);
void    BZ2_bzReadGetUnused(int* bzerror, BZFILE* b, void** unused, int* nUnused);
int     BZ2_bzRead(int* bzerror, BZFILE* b, void* buf, int len);

/* zlib-flavoured layer (reference bzlib.c:1448-1629) */
BZFILE*     BZ2_bzopen(const char* path, const char* mode);
BZFILE*     BZ2_bzdopen(int fd, const char* mode);
int         BZ2_bzread(BZFILE* b, void* buf, int len);
int         BZ2_bzwrite(BZFILE* b, void* buf, int len);
int         BZ2_bzflush(BZFILE* b);
void        BZ2_bzclose(BZFILE* b);
const char* BZ2_bzerror(BZFILE* b, int* errnum);

#ifdef __cplusplus
}
#endif
#endif
