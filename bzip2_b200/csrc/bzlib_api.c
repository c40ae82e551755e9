/*
 * bzlib_api.c -- libbz2-compatible compression front end (plain C) over the GPU engine.
 *
 * Mirrors the reference's public behaviour for the compression path:
 *   BZ2_bzCompressInit / BZ2_bzCompress / BZ2_bzCompressEnd   (bzlib.c:144-207, :400-454, :458-474)
 *   BZ2_bzBuffToBuffCompress                                   (bzlib.c:1309-1357)
 *   BZ2_bzWriteOpen / BZ2_bzWrite / BZ2_bzWriteClose[64]       (bzlib.c:978-1146)
 * Return codes, mode transitions (RUNNING / FLUSHING / FINISHING / IDLE), total_in/out
 * accounting and parameter checks follow those functions.  What differs, within the
 * freedom the API leaves: input is accepted as fast as it is offered and compressed on
 * the GPU a window (~100 blocks) at a time, so output becomes available in bursts.
 * There is no CPU codec here: if no CUDA device is usable, Init returns BZ_CONFIG_ERROR.
 */
#include "../../include/bzlib.h"
#include "../../include/bz2_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

enum { M_IDLE = 1, M_RUNNING = 2, M_FLUSHING = 3, M_FINISHING = 4 };

typedef struct {
   bz_stream* strm;               /* back pointer, checked on every call (bzlib.c:404) */
   bz2b200_engine* eng;
   int level;
   int mode;
   unsigned int avail_in_expect;
   int closing_fed;               /* the flush/finish marker has reached the engine */
   /* compressed bytes produced by the engine but not yet handed to the caller */
   unsigned char* obuf;
   size_t ocap, olen, opos;
   int sticky_err;
} cstate;

/* ---- a small pool of idle engines: creating one allocates several GB of HBM ---------------- */
#define POOL_MAX 3
#define SMALL_INPUT ((size_t)8 << 20)      /* one-shot calls up to this size use an engine sized to the input */
static pthread_mutex_t pool_mu = PTHREAD_MUTEX_INITIALIZER;
static struct { bz2b200_engine* e; int level; int device; int window_mb; } pool[POOL_MAX];

static int env_int(const char* name, int dflt)
{
   const char* v = getenv(name);
   return (v && *v) ? atoi(v) : dflt;
}

/* window key of the pool: the configured window in MB, or -k for an engine bounded to inputs of at most k MiB */
static int window_key(size_t bounded_bytes)
{
   if (bounded_bytes) { int k = 1; while (((size_t)k << 20) < bounded_bytes) k <<= 1; return -k; }
   return env_int("BZ2_B200_WINDOW_MB", 0);
}

static int engine_acquire(bz2b200_engine** out, int level, size_t bounded_bytes)
{
   const int device = env_int("BZ2_B200_DEVICE", 0);
   const int wkey = window_key(bounded_bytes);
   int i;
   pthread_mutex_lock(&pool_mu);
   for (i = 0; i < POOL_MAX; i++) {
      if (pool[i].e && pool[i].level == level && pool[i].device == device && pool[i].window_mb == wkey) {
         *out = pool[i].e; pool[i].e = NULL;
         pthread_mutex_unlock(&pool_mu);
         return 0;
      }
   }
   pthread_mutex_unlock(&pool_mu);
   if (wkey < 0) return bz2b200_engine_create_bounded(out, device, level, (size_t)(-wkey) << 20);
   return bz2b200_engine_create(out, device, level, (size_t)wkey << 20);
}

static void engine_release(bz2b200_engine* e, int level, size_t bounded_bytes)
{
   const int device = env_int("BZ2_B200_DEVICE", 0);
   const int wkey = window_key(bounded_bytes);
   int i;
   bz2b200_engine* victim = e;
   pthread_mutex_lock(&pool_mu);
   for (i = 0; i < POOL_MAX; i++) {
      if (!pool[i].e) { pool[i].e = e; pool[i].level = level; pool[i].device = device; pool[i].window_mb = wkey; victim = NULL; break; }
   }
   if (victim) { victim = pool[0].e; pool[0].e = e; pool[0].level = level; pool[0].device = device; pool[0].window_mb = wkey; }
   pthread_mutex_unlock(&pool_mu);
   if (victim) bz2b200_engine_destroy(victim);
}

/* ---- several engines on one stream (bz2b200_multi_*): BZ2_B200_DEVICES=0,1,2,3 shards the one-shot call by block
 * over those GPUs; a GPU listed twice (0,0) keeps two windows in flight on it.  One set of engines is kept per
 * process and reused while level and device list stay the same. */
static struct { bz2b200_multi* m; int level; char devs[128]; } mpool;

static int parse_devices(const char* v, int* out, int max)
{
   int n = 0;
   while (v && *v && n < max) {
      char* end;
      long d = strtol(v, &end, 10);
      if (end == v) break;
      out[n++] = (int)d;
      if (*end != ',') break;
      v = end + 1;
   }
   return n;
}

static bz2b200_multi* multi_acquire(int level, int* rc_out)
{
   const char* v = getenv("BZ2_B200_DEVICES");
   int devs[16], n;
   bz2b200_multi* m = NULL;
   *rc_out = 0;
   if (!v || !*v) return NULL;
   n = parse_devices(v, devs, 16);
   if (n < 2) return NULL;
   pthread_mutex_lock(&pool_mu);
   if (mpool.m && mpool.level == level && !strncmp(mpool.devs, v, sizeof mpool.devs)) { m = mpool.m; mpool.m = NULL; }
   pthread_mutex_unlock(&pool_mu);
   if (m) return m;
   *rc_out = bz2b200_multi_create(&m, devs, n, level, (size_t)env_int("BZ2_B200_WINDOW_MB", 0) << 20);
   return *rc_out ? NULL : m;
}

static void multi_release(bz2b200_multi* m, int level)
{
   const char* v = getenv("BZ2_B200_DEVICES");
   bz2b200_multi* victim;
   pthread_mutex_lock(&pool_mu);
   victim = mpool.m;
   mpool.m = m; mpool.level = level;
   snprintf(mpool.devs, sizeof mpool.devs, "%s", v ? v : "");
   pthread_mutex_unlock(&pool_mu);
   if (victim) bz2b200_multi_destroy(victim);
}

/* Called at process exit or by tests that want the HBM back. */
void bz2b200_pool_clear(void)
{
   int i;
   bz2b200_multi* m;
   pthread_mutex_lock(&pool_mu);
   for (i = 0; i < POOL_MAX; i++) if (pool[i].e) { bz2b200_engine_destroy(pool[i].e); pool[i].e = NULL; }
   m = mpool.m; mpool.m = NULL;
   pthread_mutex_unlock(&pool_mu);
   if (m) bz2b200_multi_destroy(m);
}

static int map_engine_error(int rc)
{
   switch (rc) {
      case BZ2B200_OK: return BZ_OK;
      case BZ2B200_ENOMEM: return BZ_MEM_ERROR;
      case BZ2B200_EOUTFULL: return BZ_OUTBUFF_FULL;
      case BZ2B200_EPARAM: return BZ_PARAM_ERROR;
      default: return BZ_CONFIG_ERROR;          /* no device / CUDA failure: never silent */
   }
}

static void* dflt_alloc(void* opaque, int items, int size) { (void)opaque; return malloc((size_t)items * (size_t)size); }
static void dflt_free(void* opaque, void* p) { (void)opaque; if (p) free(p); }

/* ------------------------------------------------------------------------------------------ */
int BZ2_bzCompressInit(bz_stream* strm, int blockSize100k, int verbosity, int workFactor)
{
   cstate* s;
   int rc;
   /* verbosity is not range-checked here (bzlib.c:155-158) */
   if (sizeof(int) != 4 || sizeof(short) != 2 || sizeof(char) != 1) return BZ_CONFIG_ERROR;
   if (strm == NULL || blockSize100k < 1 || blockSize100k > 9 || workFactor < 0 || workFactor > 250)
      return BZ_PARAM_ERROR;
   if (strm->bzalloc == NULL) strm->bzalloc = dflt_alloc;
   if (strm->bzfree == NULL) strm->bzfree = dflt_free;
   s = (cstate*)strm->bzalloc(strm->opaque, (int)sizeof(cstate), 1);
   if (s == NULL) return BZ_MEM_ERROR;
   memset(s, 0, sizeof *s);
   s->strm = strm;
   s->level = blockSize100k;
   rc = engine_acquire(&s->eng, blockSize100k, 0);
   if (rc == 0) rc = bz2b200_stream_begin(s->eng);
   if (rc == 0) bz2b200_engine_set_verbosity(s->eng, verbosity);
   if (rc) {
      if (s->eng) bz2b200_engine_destroy(s->eng);
      strm->bzfree(strm->opaque, s);
      return map_engine_error(rc) == BZ_MEM_ERROR ? BZ_MEM_ERROR : BZ_CONFIG_ERROR;
   }
   s->mode = M_RUNNING;
   strm->state = s;
   strm->total_in_lo32 = strm->total_in_hi32 = 0;
   strm->total_out_lo32 = strm->total_out_hi32 = 0;
   return BZ_OK;
}

static int sink_append(void* user, const void* bytes, size_t n)
{
   cstate* s = (cstate*)user;
   if (s->opos == s->olen) s->opos = s->olen = 0;
   if (s->olen + n > s->ocap) {
      /* the caller's allocator (bzlib.c:104-115) has no realloc: new buffer, copy, free */
      size_t nc = s->ocap ? s->ocap : (1u << 20);
      unsigned char* nb;
      while (nc < s->olen + n) nc *= 2;
      if (nc > 0x7fffffffu) return BZ2B200_ENOMEM;
      nb = (unsigned char*)s->strm->bzalloc(s->strm->opaque, (int)nc, 1);
      if (!nb) return BZ2B200_ENOMEM;
      if (s->olen) memcpy(nb, s->obuf, s->olen);
      if (s->obuf) s->strm->bzfree(s->strm->opaque, s->obuf);
      s->obuf = nb; s->ocap = nc;
   }
   memcpy(s->obuf + s->olen, bytes, n);
   s->olen += n;
   return 0;
}

/* hand the engine everything the caller offers; returns bytes consumed or <0 */
static long take_input(cstate* s, int end_mode)
{
   bz_stream* strm = s->strm;
   const unsigned int n = strm->avail_in;
   unsigned int t;
   int rc;
   if (n == 0 && end_mode == 0) return 0;
   rc = bz2b200_stream_feed(s->eng, strm->next_in, n, end_mode, sink_append, s);
   if (rc) { s->sticky_err = map_engine_error(rc); return -1; }
   strm->next_in += n;
   strm->avail_in = 0;
   t = strm->total_in_lo32 + n;
   if (t < strm->total_in_lo32) strm->total_in_hi32++;
   strm->total_in_lo32 = t;
   return (long)n;
}

static unsigned int give_output(cstate* s)
{
   bz_stream* strm = s->strm;
   size_t have = s->olen - s->opos;
   unsigned int cnt = strm->avail_out, t;
   if (have < cnt) cnt = (unsigned int)have;
   if (cnt) {
      memcpy(strm->next_out, s->obuf + s->opos, cnt);
      s->opos += cnt;
      strm->next_out += cnt;
      strm->avail_out -= cnt;
      t = strm->total_out_lo32 + cnt;
      if (t < strm->total_out_lo32) strm->total_out_hi32++;
      strm->total_out_lo32 = t;
   }
   return cnt;
}

int BZ2_bzCompress(bz_stream* strm, int action)
{
   cstate* s;
   long in;
   unsigned int out;
   if (strm == NULL || (s = (cstate*)strm->state) == NULL || s->strm != strm) return BZ_PARAM_ERROR;
   if (s->sticky_err) return s->sticky_err;

   for (;;) switch (s->mode) {
      case M_IDLE:
         return BZ_SEQUENCE_ERROR;

      case M_RUNNING:
         if (action == BZ_RUN) {
            in = take_input(s, 0);
            if (in < 0) return s->sticky_err;
            out = give_output(s);
            return (in > 0 || out > 0) ? BZ_RUN_OK : BZ_PARAM_ERROR;
         }
         if (action == BZ_FLUSH)  { s->avail_in_expect = strm->avail_in; s->mode = M_FLUSHING;  s->closing_fed = 0; continue; }
         if (action == BZ_FINISH) { s->avail_in_expect = strm->avail_in; s->mode = M_FINISHING; s->closing_fed = 0; continue; }
         return BZ_PARAM_ERROR;

      case M_FLUSHING:
         if (action != BZ_FLUSH) return BZ_SEQUENCE_ERROR;
         if (s->avail_in_expect != strm->avail_in) return BZ_SEQUENCE_ERROR;
         if (!s->closing_fed) {
            in = take_input(s, 1);
            if (in < 0) return s->sticky_err;
            s->closing_fed = 1; s->avail_in_expect = 0;
         }
         give_output(s);
         if (s->opos < s->olen) return BZ_FLUSH_OK;
         s->mode = M_RUNNING;
         return BZ_RUN_OK;

      case M_FINISHING:
         if (action != BZ_FINISH) return BZ_SEQUENCE_ERROR;
         if (s->avail_in_expect != strm->avail_in) return BZ_SEQUENCE_ERROR;
         in = 0;
         if (!s->closing_fed) {
            in = take_input(s, 2);
            if (in < 0) return s->sticky_err;
            s->closing_fed = 1; s->avail_in_expect = 0;
            in = 1;                                  /* the trailer counts as progress */
         }
         out = give_output(s);
         if (in == 0 && out == 0) return BZ_SEQUENCE_ERROR;
         if (s->opos < s->olen) return BZ_FINISH_OK;
         s->mode = M_IDLE;
         return BZ_STREAM_END;

      default:
         return BZ_OK;
   }
}

int BZ2_bzCompressEnd(bz_stream* strm)
{
   cstate* s;
   if (strm == NULL) return BZ_PARAM_ERROR;
   s = (cstate*)strm->state;
   if (s == NULL || s->strm != strm) return BZ_PARAM_ERROR;
   if (s->eng) engine_release(s->eng, s->level, 0);
   if (s->obuf) strm->bzfree(strm->opaque, s->obuf);
   strm->bzfree(strm->opaque, s);
   strm->state = NULL;
   return BZ_OK;
}

int BZ2_bzBuffToBuffCompress(char* dest, unsigned int* destLen, char* source, unsigned int sourceLen,
                             int blockSize100k, int verbosity, int workFactor)
{
   bz2b200_engine* eng = NULL;
   size_t dlen;
   int rc;
   if (dest == NULL || destLen == NULL || source == NULL || blockSize100k < 1 || blockSize100k > 9 ||
       verbosity < 0 || verbosity > 4 || workFactor < 0 || workFactor > 250)
      return BZ_PARAM_ERROR;
   {
      bz2b200_multi* m = multi_acquire(blockSize100k, &rc);
      if (rc) return map_engine_error(rc) == BZ_MEM_ERROR ? BZ_MEM_ERROR : BZ_CONFIG_ERROR;
      if (m) {
         dlen = *destLen;
         rc = bz2b200_multi_compress(m, source, NULL, sourceLen, dest, &dlen, 0, NULL);
         multi_release(m, blockSize100k);
         if (rc == BZ2B200_EOUTFULL) return BZ_OUTBUFF_FULL;
         if (rc) return map_engine_error(rc);
         *destLen = (unsigned int)dlen;
         return BZ_OK;
      }
   }
   {
      /* a small input gets an engine sized to it (a few hundred MB of HBM instead of ~9 GB) */
      const size_t bounded = (sourceLen <= SMALL_INPUT) ? (sourceLen ? sourceLen : 1) : 0;
      rc = engine_acquire(&eng, blockSize100k, bounded);
      if (rc) return map_engine_error(rc) == BZ_MEM_ERROR ? BZ_MEM_ERROR : BZ_CONFIG_ERROR;
      bz2b200_engine_set_verbosity(eng, verbosity);
      dlen = *destLen;
      rc = bz2b200_compress_host(eng, source, sourceLen, dest, &dlen, 0, NULL);
      engine_release(eng, blockSize100k, bounded);
   }
   if (rc == BZ2B200_EOUTFULL) return BZ_OUTBUFF_FULL;
   if (rc) return map_engine_error(rc);
   *destLen = (unsigned int)dlen;
   return BZ_OK;
}

const char* BZ2_bzlibVersion(void) { return "1.0.6x-b200, 18-Oct-2026"; }

/* ---- stdio write side ---------------------------------------------------------------------- */
typedef struct {
   FILE* handle;
   char buf[BZ_MAX_UNUSED];
   bz_stream strm;
   int last_err;
   int writing;
} wfile;

#define SETERR(v) do { if (bzerror) *bzerror = (v); if (bzf) bzf->last_err = (v); } while (0)

BZFILE* BZ2_bzWriteOpen(int* bzerror, FILE* f, int blockSize100k, int verbosity, int workFactor)
{
   wfile* bzf = NULL;
   int ret;
   SETERR(BZ_OK);
   if (f == NULL || blockSize100k < 1 || blockSize100k > 9 || workFactor < 0 || workFactor > 250 ||
       verbosity < 0 || verbosity > 4) { SETERR(BZ_PARAM_ERROR); return NULL; }
   if (ferror(f)) { SETERR(BZ_IO_ERROR); return NULL; }
   bzf = (wfile*)malloc(sizeof(wfile));
   if (!bzf) { SETERR(BZ_MEM_ERROR); return NULL; }
   memset(bzf, 0, sizeof *bzf);
   bzf->handle = f; bzf->writing = 1; bzf->last_err = BZ_OK;
   ret = BZ2_bzCompressInit(&bzf->strm, blockSize100k, verbosity, workFactor);
   if (ret != BZ_OK) { wfile* t = bzf; bzf = NULL; SETERR(ret); free(t); return NULL; }
   bzf->strm.avail_in = 0;
   return bzf;
}

static int pump(wfile* bzf, int action, int* ret_out)
{
   size_t n;
   int ret;
   bzf->strm.avail_out = BZ_MAX_UNUSED;
   bzf->strm.next_out = bzf->buf;
   ret = BZ2_bzCompress(&bzf->strm, action);
   *ret_out = ret;
   if (bzf->strm.avail_out < BZ_MAX_UNUSED) {
      n = BZ_MAX_UNUSED - bzf->strm.avail_out;
      if (fwrite(bzf->buf, 1, n, bzf->handle) != n || ferror(bzf->handle)) return BZ_IO_ERROR;
   }
   return BZ_OK;
}

void BZ2_bzWrite(int* bzerror, BZFILE* b, void* buf, int len)
{
   wfile* bzf = (wfile*)b;
   int ret;
   SETERR(BZ_OK);
   if (bzf == NULL || buf == NULL || len < 0) { SETERR(BZ_PARAM_ERROR); return; }
   if (!bzf->writing) { SETERR(BZ_SEQUENCE_ERROR); return; }
   if (ferror(bzf->handle)) { SETERR(BZ_IO_ERROR); return; }
   if (len == 0) { SETERR(BZ_OK); return; }
   bzf->strm.avail_in = (unsigned int)len;
   bzf->strm.next_in = (char*)buf;
   for (;;) {
      int io = pump(bzf, BZ_RUN, &ret);
      if (ret != BZ_RUN_OK) { SETERR(ret); return; }
      if (io != BZ_OK) { SETERR(io); return; }
      /* unlike the reference we also drain whatever the engine has already produced */
      if (bzf->strm.avail_in == 0) {
         cstate* cs = (cstate*)bzf->strm.state;
         if (cs->opos == cs->olen) { SETERR(BZ_OK); return; }
      }
   }
}

void BZ2_bzWriteClose64(int* bzerror, BZFILE* b, int abandon, unsigned int* in_lo, unsigned int* in_hi,
                        unsigned int* out_lo, unsigned int* out_hi)
{
   wfile* bzf = (wfile*)b;
   int ret;
   if (bzf == NULL) { SETERR(BZ_OK); return; }
   if (!bzf->writing) { SETERR(BZ_SEQUENCE_ERROR); return; }
   if (ferror(bzf->handle)) { SETERR(BZ_IO_ERROR); return; }
   if (in_lo) *in_lo = 0;
   if (in_hi) *in_hi = 0;
   if (out_lo) *out_lo = 0;
   if (out_hi) *out_hi = 0;
   if (!abandon && bzf->last_err == BZ_OK) {
      for (;;) {
         int io = pump(bzf, BZ_FINISH, &ret);
         if (ret != BZ_FINISH_OK && ret != BZ_STREAM_END) { SETERR(ret); return; }
         if (io != BZ_OK) { SETERR(io); return; }
         if (ret == BZ_STREAM_END) break;
      }
   }
   if (!abandon && !ferror(bzf->handle)) {
      fflush(bzf->handle);
      if (ferror(bzf->handle)) { SETERR(BZ_IO_ERROR); return; }
   }
   if (in_lo) *in_lo = bzf->strm.total_in_lo32;
   if (in_hi) *in_hi = bzf->strm.total_in_hi32;
   if (out_lo) *out_lo = bzf->strm.total_out_lo32;
   if (out_hi) *out_hi = bzf->strm.total_out_hi32;
   SETERR(BZ_OK);
   BZ2_bzCompressEnd(&bzf->strm);
   free(bzf);
}

void BZ2_bzWriteClose(int* bzerror, BZFILE* b, int abandon, unsigned int* nbytes_in, unsigned int* nbytes_out)
{
   BZ2_bzWriteClose64(bzerror, b, abandon, nbytes_in, NULL, nbytes_out, NULL);
}

/* ---- stdio read side (bzlib.c:1150-1301); decoding runs on the host (bzlib_decode.c) --------- */
static int at_eof(FILE* f)
{
   int c = fgetc(f);
   if (c == EOF) return 1;
   ungetc(c, f);
   return 0;
}

BZFILE* BZ2_bzReadOpen(int* bzerror, FILE* f, int verbosity, int small, void* unused, int nUnused)
{
   wfile* bzf = NULL;
   int ret;
   SETERR(BZ_OK);
   if (f == NULL || (small != 0 && small != 1) || verbosity < 0 || verbosity > 4 ||
       (unused == NULL && nUnused != 0) || (unused != NULL && (nUnused < 0 || nUnused > BZ_MAX_UNUSED))) {
      SETERR(BZ_PARAM_ERROR); return NULL;
   }
   if (ferror(f)) { SETERR(BZ_IO_ERROR); return NULL; }
   bzf = (wfile*)malloc(sizeof(wfile));
   if (!bzf) { SETERR(BZ_MEM_ERROR); return NULL; }
   memset(bzf, 0, sizeof *bzf);
   bzf->handle = f; bzf->writing = 0; bzf->last_err = BZ_OK;
   if (nUnused > 0) memcpy(bzf->buf, unused, (size_t)nUnused);
   ret = BZ2_bzDecompressInit(&bzf->strm, verbosity, small);
   if (ret != BZ_OK) { wfile* t = bzf; bzf = NULL; SETERR(ret); free(t); return NULL; }
   bzf->strm.avail_in = (unsigned int)nUnused;
   bzf->strm.next_in = bzf->buf;
   return bzf;
}

void BZ2_bzReadClose(int* bzerror, BZFILE* b)
{
   wfile* bzf = (wfile*)b;
   SETERR(BZ_OK);
   if (bzf == NULL) return;
   if (bzf->writing) { SETERR(BZ_SEQUENCE_ERROR); return; }
   (void)BZ2_bzDecompressEnd(&bzf->strm);
   free(bzf);
}

int BZ2_bzRead(int* bzerror, BZFILE* b, void* buf, int len)
{
   wfile* bzf = (wfile*)b;
   int ret;
   SETERR(BZ_OK);
   if (bzf == NULL || buf == NULL || len < 0) { SETERR(BZ_PARAM_ERROR); return 0; }
   if (bzf->writing) { SETERR(BZ_SEQUENCE_ERROR); return 0; }
   if (len == 0) return 0;
   bzf->strm.avail_out = (unsigned int)len;
   bzf->strm.next_out = (char*)buf;
   for (;;) {
      if (ferror(bzf->handle)) { SETERR(BZ_IO_ERROR); return 0; }
      if (bzf->strm.avail_in == 0 && !at_eof(bzf->handle)) {
         size_t n = fread(bzf->buf, 1, BZ_MAX_UNUSED, bzf->handle);
         if (ferror(bzf->handle)) { SETERR(BZ_IO_ERROR); return 0; }
         bzf->strm.avail_in = (unsigned int)n;
         bzf->strm.next_in = bzf->buf;
      }
      ret = BZ2_bzDecompress(&bzf->strm);
      if (ret != BZ_OK && ret != BZ_STREAM_END) { SETERR(ret); return 0; }
      if (ret == BZ_OK && at_eof(bzf->handle) && bzf->strm.avail_in == 0 && bzf->strm.avail_out > 0) {
         SETERR(BZ_UNEXPECTED_EOF); return 0;
      }
      if (ret == BZ_STREAM_END) { SETERR(BZ_STREAM_END); return len - (int)bzf->strm.avail_out; }
      if (bzf->strm.avail_out == 0) { SETERR(BZ_OK); return len; }
   }
}

void BZ2_bzReadGetUnused(int* bzerror, BZFILE* b, void** unused, int* nUnused)
{
   wfile* bzf = (wfile*)b;
   if (bzf == NULL) { SETERR(BZ_PARAM_ERROR); return; }
   if (bzf->last_err != BZ_STREAM_END) { SETERR(BZ_SEQUENCE_ERROR); return; }
   if (unused == NULL || nUnused == NULL) { SETERR(BZ_PARAM_ERROR); return; }
   SETERR(BZ_OK);
   *nUnused = (int)bzf->strm.avail_in;
   *unused = bzf->strm.next_in;
}

/* ---- zlib-flavoured convenience layer (bzlib.c:1448-1629) ------------------------------------ */
static BZFILE* open_common(const char* path, int fd, const char* mode, int by_fd)
{
   int err, level = 9, writing = 0, small = 0;
   FILE* fp;
   BZFILE* h;
   if (mode == NULL) return NULL;
   for (; *mode; mode++) {
      if (*mode == 'r') writing = 0;
      else if (*mode == 'w') writing = 1;
      else if (*mode == 's') small = 1;
      else if (*mode >= '0' && *mode <= '9') level = *mode - '0';
   }
   if (by_fd) fp = fdopen(fd, writing ? "wb" : "rb");
   else if (path == NULL || path[0] == 0) fp = writing ? stdout : stdin;
   else fp = fopen(path, writing ? "wb" : "rb");
   if (fp == NULL) return NULL;
   if (writing) {
      if (level < 1) level = 1;
      if (level > 9) level = 9;
      h = BZ2_bzWriteOpen(&err, fp, level, 0, 30);
   } else {
      h = BZ2_bzReadOpen(&err, fp, 0, small, NULL, 0);
   }
   if (h == NULL && fp != stdin && fp != stdout) fclose(fp);
   return h;
}

BZFILE* BZ2_bzopen(const char* path, const char* mode) { return open_common(path, -1, mode, 0); }
BZFILE* BZ2_bzdopen(int fd, const char* mode) { return open_common(NULL, fd, mode, 1); }

int BZ2_bzread(BZFILE* b, void* buf, int len)
{
   int err, n;
   if (((wfile*)b)->last_err == BZ_STREAM_END) return 0;
   n = BZ2_bzRead(&err, b, buf, len);
   return (err == BZ_OK || err == BZ_STREAM_END) ? n : -1;
}

int BZ2_bzwrite(BZFILE* b, void* buf, int len)
{
   int err;
   BZ2_bzWrite(&err, b, buf, len);
   return err == BZ_OK ? len : -1;
}

int BZ2_bzflush(BZFILE* b) { (void)b; return 0; }

void BZ2_bzclose(BZFILE* b)
{
   int err;
   FILE* fp;
   if (b == NULL) return;
   fp = ((wfile*)b)->handle;
   if (((wfile*)b)->writing) {
      BZ2_bzWriteClose(&err, b, 0, NULL, NULL);
      if (err != BZ_OK) BZ2_bzWriteClose(NULL, b, 1, NULL, NULL);
   } else {
      BZ2_bzReadClose(&err, b);
   }
   if (fp != stdin && fp != stdout) fclose(fp);
}

const char* BZ2_bzerror(BZFILE* b, int* errnum)
{
   static const char* const names[] = { "OK", "SEQUENCE_ERROR", "PARAM_ERROR", "MEM_ERROR", "DATA_ERROR",
      "DATA_ERROR_MAGIC", "IO_ERROR", "UNEXPECTED_EOF", "OUTBUFF_FULL", "CONFIG_ERROR" };
   int err = ((wfile*)b)->last_err;
   if (err > 0) err = 0;
   *errnum = err;
   return (-err < 10) ? names[-err] : "???";
}
