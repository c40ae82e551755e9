/*
 * bzip2_cli.c -- `bzip2` command-line front end, compression side, over the GPU library.
 *
 * Mirrors the user-visible behaviour of the reference CLI in compress mode
 * (bzip2.c: flag parsing :1870-1934, compressStream :328-427, compress() :1132-1309):
 *   -1..-9 / --fast / --best, -z, -c, -k, -f, -q, -v, -s (clamps the level to 2,
 *   bzip2.c:1937-1938), BZIP2 / BZIP environment flags, "file" -> "file.bz2", refusal to
 *   overwrite without -f, refusal to write compressed data to a terminal, input removed
 *   after success unless -k / -c.
 * It feeds the libbz2-compatible stdio API (BZ2_bzWriteOpen / BZ2_bzWrite /
 * BZ2_bzWriteClose64), with 4 MiB reads instead of the reference's 5000-byte trickle
 * (bzip2.c:350-358) -- the stream bytes do not depend on the chunking.
 * Decompression (-d, -t) is outside this build: use the reference's bzip2 for that.
 */
#include "../../include/bzlib.h"
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

static int level = 9, to_stdout = 0, keep = 0, force = 0, quiet = 0, verbose = 0, small = 0;
static const char* prog = "bzip2";

static void usage(void)
{
   fprintf(stderr,
      "bzip2-b200, a GPU (sm_100a) block-sorting file compressor, stream-compatible with bzip2 1.0.6.\n\n"
      "   usage: %s [flags and input files in any order]\n\n"
      "   -z --compress       compress (the only mode of this build)\n"
      "   -k --keep           keep (don't delete) input files\n"
      "   -f --force          overwrite existing output files\n"
      "   -c --stdout         output to standard out\n"
      "   -q --quiet          suppress noncritical error messages\n"
      "   -v --verbose        be verbose\n"
      "   -s --small          use block size 200k at most\n"
      "   -1 .. -9            set block size to 100k .. 900k\n"
      "   --fast / --best     alias for -1 / -9\n\n"
      "   If no file names are given, compresses standard input to standard output.\n", prog);
}

static int compress_stream(FILE* in, FILE* out, const char* name, unsigned long long* nin, unsigned long long* nout)
{
   enum { CHUNK = 4 << 20 };
   int bzerr = BZ_OK;
   unsigned int in_lo = 0, in_hi = 0, out_lo = 0, out_hi = 0;
   char* buf = (char*)malloc(CHUNK);
   BZFILE* bz;
   if (!buf) { fprintf(stderr, "%s: out of memory\n", prog); return 1; }
   bz = BZ2_bzWriteOpen(&bzerr, out, small && level > 2 ? 2 : level, verbose, 30);
   if (bzerr != BZ_OK) {
      fprintf(stderr, "%s: cannot start the GPU compressor (libbz2 error %d%s)\n", prog, bzerr,
              bzerr == BZ_CONFIG_ERROR ? ": no usable CUDA device; this build has no CPU path" : "");
      free(buf);
      return 1;
   }
   for (;;) {
      size_t n = fread(buf, 1, CHUNK, in);
      if (ferror(in)) { fprintf(stderr, "%s: %s: read error: %s\n", prog, name, strerror(errno)); BZ2_bzWriteClose64(&bzerr, bz, 1, 0, 0, 0, 0); free(buf); return 1; }
      if (n > 0) {
         BZ2_bzWrite(&bzerr, bz, buf, (int)n);
         if (bzerr != BZ_OK) { fprintf(stderr, "%s: %s: compression failed (libbz2 error %d)\n", prog, name, bzerr); BZ2_bzWriteClose64(&bzerr, bz, 1, 0, 0, 0, 0); free(buf); return 1; }
      }
      if (n < CHUNK) break;
   }
   BZ2_bzWriteClose64(&bzerr, bz, 0, &in_lo, &in_hi, &out_lo, &out_hi);
   free(buf);
   if (bzerr != BZ_OK) { fprintf(stderr, "%s: %s: compression failed (libbz2 error %d)\n", prog, name, bzerr); return 1; }
   if (fflush(out) != 0) { fprintf(stderr, "%s: write error: %s\n", prog, strerror(errno)); return 1; }
   *nin = ((unsigned long long)in_hi << 32) | in_lo;
   *nout = ((unsigned long long)out_hi << 32) | out_lo;
   return 0;
}

static void report(const char* name, unsigned long long nin, unsigned long long nout)
{
   if (!verbose) return;
   if (nin == 0) { fprintf(stderr, "  %s: no data compressed.\n", name); return; }
   fprintf(stderr, "  %s: %6.3f:1, %6.3f bits/byte, %5.2f%% saved, %llu in, %llu out.\n", name,
           (double)nin / (double)nout, 8.0 * (double)nout / (double)nin, 100.0 * (1.0 - (double)nout / (double)nin), nin, nout);
}

static int do_file(const char* name)
{
   char outname[4096];
   struct stat st;
   FILE *in, *out;
   unsigned long long nin = 0, nout = 0;
   int rc;
   size_t L = strlen(name);
   if (L + 5 > sizeof outname) { fprintf(stderr, "%s: file name too long: %s\n", prog, name); return 1; }
   if (stat(name, &st) != 0) { fprintf(stderr, "%s: Can't open input file %s: %s.\n", prog, name, strerror(errno)); return 1; }
   if (S_ISDIR(st.st_mode)) { fprintf(stderr, "%s: Input file %s is a directory.\n", prog, name); return 1; }
   if (L > 4 && strcmp(name + L - 4, ".bz2") == 0 && !force) {
      if (!quiet) fprintf(stderr, "%s: Input file %s already has .bz2 suffix.\n", prog, name);
      return 1;
   }
   in = fopen(name, "rb");
   if (!in) { fprintf(stderr, "%s: Can't open input file %s: %s.\n", prog, name, strerror(errno)); return 1; }
   if (to_stdout) {
      rc = compress_stream(in, stdout, name, &nin, &nout);
      fclose(in);
      if (!rc) report(name, nin, nout);
      return rc;
   }
   snprintf(outname, sizeof outname, "%s.bz2", name);
   if (!force && access(outname, F_OK) == 0) {
      fprintf(stderr, "%s: Output file %s already exists.\n", prog, outname);
      fclose(in);
      return 1;
   }
   out = fopen(outname, "wb");
   if (!out) { fprintf(stderr, "%s: Can't create output file %s: %s.\n", prog, outname, strerror(errno)); fclose(in); return 1; }
   rc = compress_stream(in, out, name, &nin, &nout);
   fclose(in);
   if (fclose(out) != 0) rc = 1;
   if (rc) { remove(outname); return rc; }
   chmod(outname, st.st_mode & 07777);
   if (!keep) remove(name);
   report(name, nin, nout);
   return 0;
}

static int handle_flag(const char* a)
{
   if (a[1] == '-') {
      if (!strcmp(a, "--stdout")) to_stdout = 1;
      else if (!strcmp(a, "--compress")) {}
      else if (!strcmp(a, "--keep")) keep = 1;
      else if (!strcmp(a, "--force")) force = 1;
      else if (!strcmp(a, "--quiet")) quiet = 1;
      else if (!strcmp(a, "--verbose")) verbose++;
      else if (!strcmp(a, "--small")) small = 1;
      else if (!strcmp(a, "--fast")) level = 1;
      else if (!strcmp(a, "--best")) level = 9;
      else if (!strcmp(a, "--repetitive-fast") || !strcmp(a, "--repetitive-best") || !strcmp(a, "--exponential")) {}
      else if (!strcmp(a, "--help")) { usage(); exit(0); }
      else if (!strcmp(a, "--decompress") || !strcmp(a, "--test")) { fprintf(stderr, "%s: this build only compresses; use the reference bzip2 to decompress\n", prog); exit(1); }
      else { fprintf(stderr, "%s: Bad flag `%s'\n", prog, a); usage(); exit(1); }
      return 0;
   }
   for (const char* p = a + 1; *p; p++) {
      switch (*p) {
         case 'c': to_stdout = 1; break;
         case 'z': break;
         case 'k': keep = 1; break;
         case 'f': force = 1; break;
         case 'q': quiet = 1; break;
         case 'v': verbose++; break;
         case 's': small = 1; break;
         case 'h': usage(); exit(0);
         case 'd': case 't': fprintf(stderr, "%s: this build only compresses; use the reference bzip2 to decompress\n", prog); exit(1);
         case '1': case '2': case '3': case '4': case '5': case '6': case '7': case '8': case '9': level = *p - '0'; break;
         default: fprintf(stderr, "%s: Bad flag `%s'\n", prog, a); usage(); exit(1);
      }
   }
   return 0;
}

static void env_flags(const char* var)
{
   const char* v = getenv(var);
   char tmp[1024];
   char* tok;
   if (!v) return;
   snprintf(tmp, sizeof tmp, "%s", v);
   for (tok = strtok(tmp, " \t"); tok; tok = strtok(NULL, " \t")) if (tok[0] == '-' && tok[1]) handle_flag(tok);
}

int main(int argc, char** argv)
{
   int i, nfiles = 0, rc = 0, dashdash = 0;
   const char* slash = strrchr(argv[0], '/');
   prog = slash ? slash + 1 : argv[0];
   env_flags("BZIP2");
   env_flags("BZIP");
   for (i = 1; i < argc; i++) {
      if (!dashdash && !strcmp(argv[i], "--")) { dashdash = 1; continue; }
      if (!dashdash && argv[i][0] == '-' && argv[i][1]) handle_flag(argv[i]); else nfiles++;
   }
   if (verbose > 4) verbose = 4;
   if (nfiles == 0) {
      unsigned long long nin = 0, nout = 0;
      if (isatty(fileno(stdout)) && !force) {
         fprintf(stderr, "%s: I won't write compressed data to a terminal.\n%s: For help, type: `%s --help'.\n", prog, prog, prog);
         return 1;
      }
      rc = compress_stream(stdin, stdout, "(stdin)", &nin, &nout);
      if (!rc) report("(stdin)", nin, nout);
      return rc;
   }
   dashdash = 0;
   for (i = 1; i < argc; i++) {
      if (!dashdash && !strcmp(argv[i], "--")) { dashdash = 1; continue; }
      if (!dashdash && argv[i][0] == '-' && argv[i][1]) continue;
      rc |= do_file(argv[i]);
   }
   return rc;
}
