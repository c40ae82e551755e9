// stage3_mtf.cu -- S3: move-to-front + zero-run (RUNA/RUNB) coding of the BWT output.
//
// Replaces generateMTFValues (reference compress.c:93-229).  The reference is a serial
// recurrence over the block; here the block is cut into tiles of MTF_TILE symbols:
//   k_mtf_summary  (warp/tile)  distinct symbols of the tile, most recent first.  The
//                               effect of a tile on the MTF list is "move these to the
//                               front in this order", which composes left to right.
//   k_mtf_lists    (warp/block) folds the summaries, storing the list each tile starts with
//   k_mtf_encode   (thread/tile) the real MTF with the tile's list in shared memory (column
//                               layout, one column per thread); 32 tiles advance per warp
//                               instruction.  Also records the tile's zero-run shape.
//   k_rle2_scan / k_rle2_emit   runs crossing tiles are stitched by a per-block warp scan of
//                               (lead, trail, inner) summaries, then symbols are written at
//                               their final offsets and mtfFreq is accumulated.
// Output per block: mtfv[] (u16), nMTF, mtfFreq[258]; symbol values as in the reference
// (RUNA=0, RUNB=1, position p>0 -> p+1, EOB = nInUse+1).
#include "engine.h"
#include <stdlib.h>

namespace bz {

struct S3Params {
   const u8* bwt;
   const u32* X;
   const u8* inuse;
   const u32* ninuse;
   u8* lists;          // [nb * tiles_max * 256]
   u32* tilecnt;       // [nb * tiles_max]
   u32* mode;          // [nb] 1: warp-cooperative MTF (high-entropy block), 0: thread per tile
   u32 warp_threshold; // mean distinct symbols per tile above which the warp variant is used
   u32* tmeta;         // [nb * tiles_max * 4]: lead, trail|allzero<<31, inner, out_base
   u32* tcarry;        // [nb * tiles_max]
   u8* z;
   u16* mtfv;
   u32* nmtf;
   i32* mtffreq;
   u32 tiles_max;
};

__global__ void __launch_bounds__(256) k_mtf_summary(S3Params p)
{
   __shared__ u32 seen[8][8];
   const u32 w = threadIdx.x >> 5, l = lane_id();
   const u32 b = blockIdx.y;
   const u32 t = blockIdx.x * 8 + w;
   const u32 xb = p.X[b], n = p.X[b + 1] - xb;
   if ((u64)t * MTF_TILE >= n) return;
   const u32 start = xb + t * MTF_TILE;
   const u32 size = min((u32)MTF_TILE, n - t * MTF_TILE);
   const u32 nu = p.ninuse[b];
   if (l < 8) seen[w][l] = 0;
   __syncwarp();
   u8* out = p.lists + ((size_t)b * p.tiles_max + t) * 256;
   u32 count = 0;
   for (u32 off = size; off > 0 && count < nu; off = (off > 32) ? off - 32 : 0) {
      const bool valid = off > l;
      const u32 s = valid ? p.bwt[start + off - 1 - l] : (0x100u + l);
      const u32 m = __match_any_sync(FULL, s);
      const bool first = (l == (u32)(__ffs(m) - 1));
      const bool fresh = valid && first && !((seen[w][(s >> 5) & 7] >> (s & 31)) & 1u);
      const u32 bal = __ballot_sync(FULL, fresh);
      __syncwarp();
      if (fresh) {
         out[count + __popc(bal & lanemask_lt())] = (u8)s;
         atomicOr(&seen[w][s >> 5], 1u << (s & 31));
      }
      count += __popc(bal);
      __syncwarp();
   }
   if (l == 0) p.tilecnt[(size_t)b * p.tiles_max + t] = count;
}

// One WARP per block folds the tile summaries in order and stores the list every tile starts with.  Lane l
// holds list entries 8l..8l+7; the tile's symbol set is a 256-byte membership row in shared memory; the
// survivors are compacted behind the tile's own recency order through a 256-byte shared-memory row.  Only
// warp-level synchronisation per tile (the CTA-wide version needed four barriers per tile and ran 879 of them
// back to back at 12 % occupancy).
constexpr int ML_WARPS = 4;
__global__ void __launch_bounds__(ML_WARPS * 32) k_mtf_lists(S3Params p, u32 nb)
{
   __shared__ __align__(8) u8 row[ML_WARPS][256];
   __shared__ __align__(8) u8 inset[ML_WARPS][256];
   const u32 w = threadIdx.x >> 5, l = lane_id();
   const u32 b = blockIdx.x * ML_WARPS + w;
   if (b >= nb) return;
   u8* const in = inset[w];
   const u32 n = p.X[b + 1] - p.X[b];
   const u32 ntile = (n + MTF_TILE - 1) / MTF_TILE;
   const u32 nu = p.ninuse[b];
   u8* const my = row[w];
   // initial list: the in-use byte values in ascending order (compress.c:124-129), zero padded
   {
      u32 cntl = 0;
      u8 used[8];
#pragma unroll
      for (int k = 0; k < 8; k++) { used[k] = p.inuse[(size_t)b * 256 + l * 8 + k]; cntl += used[k] ? 1u : 0u; }
      u32 base = warp_incl_sum(cntl) - cntl;
      reinterpret_cast<uint2*>(my)[l] = make_uint2(0u, 0u);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; k++) if (used[k]) my[base++] = (u8)(l * 8 + k);
      __syncwarp();
   }
   uint2 cur = reinterpret_cast<const uint2*>(my)[l];
   u8* slot = p.lists + (size_t)b * p.tiles_max * 256;
   const u32* tc = p.tilecnt + (size_t)b * p.tiles_max;
   {
      // mean number of distinct symbols per tile is a proxy for the mean MTF position
      u32 sum = 0;
      for (u32 t = l; t < ntile; t += 32) sum += tc[t];
      sum = warp_sum(sum);
      if (l == 0) p.mode[b] = (sum > p.warp_threshold * ntile) ? 1u : 0u;
   }
   // the fold is a dependent chain; the summaries it consumes are fetched a group of four tiles ahead so that
   // no step waits on HBM
   constexpr int PF = 4;
   uint2 cs[PF]; u32 cc[PF];
#pragma unroll
   for (int j = 0; j < PF; j++) {
      cs[j] = make_uint2(0u, 0u); cc[j] = 0;
      if ((u32)j < ntile) { cs[j] = reinterpret_cast<const uint2*>(slot + (size_t)j * 256)[l]; cc[j] = tc[j]; }
   }
   for (u32 t0 = 0; t0 < ntile; t0 += PF) {
      uint2 ns[PF]; u32 nc[PF];
#pragma unroll
      for (int j = 0; j < PF; j++) {
         ns[j] = make_uint2(0u, 0u); nc[j] = 0;
         const u32 tn = t0 + PF + (u32)j;
         if (tn < ntile) { ns[j] = reinterpret_cast<const uint2*>(slot + (size_t)tn * 256)[l]; nc[j] = tc[tn]; }
      }
#pragma unroll
      for (int j = 0; j < PF; j++) {
         const u32 t = t0 + (u32)j;
         if (t >= ntile) break;
         const uint2 sy = cs[j];                       // the tile's distinct symbols, most recent first (cnt of them)
         const u32 cnt = cc[j];
         reinterpret_cast<uint2*>(slot + (size_t)t * 256)[l] = cur;     // the list this tile starts with
         // membership table of the tile's symbols (one byte per value, this warp's row)
         u32 sym[8];
#pragma unroll
         for (int k = 0; k < 8; k++) sym[k] = ((k < 4 ? sy.x : sy.y) >> (8 * (k & 3))) & 0xff;
         reinterpret_cast<uint2*>(in)[l] = make_uint2(0u, 0u);
         __syncwarp();
#pragma unroll
         for (int k = 0; k < 8; k++) if (l * 8 + k < cnt) in[sym[k]] = 1;
         __syncwarp();
         // survivors of the old list keep their order behind the tile's symbols
         u32 keepm = 0, nk = 0;
         u32 c[8];
#pragma unroll
         for (int k = 0; k < 8; k++) {
            c[k] = ((k < 4 ? cur.x : cur.y) >> (8 * (k & 3))) & 0xff;
            const bool keep = (l * 8 + k < nu) && !in[c[k]];
            if (keep) { keepm |= 1u << k; nk++; }
         }
         u32 dst = cnt + warp_incl_sum(nk) - nk;
         __syncwarp();
#pragma unroll
         for (int k = 0; k < 8; k++) if ((keepm >> k) & 1u) my[dst++] = (u8)c[k];
#pragma unroll
         for (int k = 0; k < 8; k++) if (l * 8 + k < cnt) my[l * 8 + k] = (u8)sym[k];
         __syncwarp();
         cur = reinterpret_cast<const uint2*>(my)[l];
      }
#pragma unroll
      for (int j = 0; j < PF; j++) { cs[j] = ns[j]; cc[j] = nc[j]; }
   }
}

__device__ __forceinline__ u32 run_digits(u32 r) { return 31 - __clz(r + 1); }

// One THREAD per tile.  The 256-entry list of thread t lives in shared memory at
// lst[j * MTF_CTA + t] (conflict-free for any mix of j across lanes).  All tiles of a block
// share the same misalignment, so the tile is read as aligned 32-bit words without divergence.
// Also records the zero-run shape of the tile (lead / trail / inner symbol count) for RLE2.
constexpr int MTF_CTA = 128;

// move-to-front inside one list word: the entry in byte `bsel` leaves, the entries below it move up one byte,
// `carry` (the entry pushed out of the previous word, or the symbol itself for word 0) enters at byte 0
__device__ __forceinline__ u32 mtf_word_hit(u32 wv, u32 bsel, u32 carry)
{
   const u32 low = (1u << (8 * bsel)) - 1u;
   return (wv & ~((low << 8) | 0xffu)) | ((wv & low) << 8) | carry;
}

__global__ void __launch_bounds__(MTF_CTA) k_mtf_encode(S3Params p)
{
   // Column layout of 16-byte groups: one uint4 = sixteen consecutive list entries (entry i = byte i&3 of word
   // (i>>2)&3 of group i>>4).  A lookup tests sixteen entries with four SIMD byte compares and the move-to-front
   // shifts sixteen entries with four funnel shifts, so deep positions (binary data) cost ~1 instruction per entry.
   __shared__ uint4 lst[16 * MTF_CTA];
   const u32 b = blockIdx.y;
   if (p.mode[b]) return;
   const u32 tid = threadIdx.x;
   const u32 t = blockIdx.x * MTF_CTA + tid;
   const u32 xb = p.X[b], n = p.X[b + 1] - xb;
   const u32 ntile = (n + MTF_TILE - 1) / MTF_TILE;
   const u32 t0 = blockIdx.x * MTF_CTA;
   if (t0 >= ntile) return;
   // stage the start lists of this CTA's tiles (coalesced 256-byte rows -> interleaved 16-byte columns)
   {
      const uint4* src = reinterpret_cast<const uint4*>(p.lists + ((size_t)b * p.tiles_max + t0) * 256);
      const u32 nt = min((u32)MTF_CTA, ntile - t0);
      for (u32 q = tid; q < nt * 16; q += MTF_CTA) {
         const u32 tile = q >> 4, g = q & 15;
         lst[g * MTF_CTA + tile] = src[q];
      }
   }
   __syncthreads();
   if (t >= ntile) return;
   const u32 start = xb + t * MTF_TILE;
   const u32 size = min((u32)MTF_TILE, n - t * MTF_TILE);
   const u32 a = start & 3u;
   const u32* in32 = reinterpret_cast<const u32*>(p.bwt + (start - a));
   u32* out32 = reinterpret_cast<u32*>(p.z + (start - a));
   uint4* my = lst + tid;
   uint4 G = my[0];                      // entries 0..15 live in registers; shared-memory row 0 is dead from here on
   u32 lead = 0, run = 0, inner = 0;
   bool seen_nz = false;
   const u32 nwords = (size + a + 3) >> 2;
   for (u32 w = 0; w < nwords; w++) {
      const u32 word = in32[w];
      u32 zword = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
         const i32 i = (i32)(w * 4 + k) - (i32)a;
         if (i < 0 || i >= (i32)size) continue;
         const u32 c = (word >> (8 * k)) & 0xff;
         const u32 cc = c * 0x01010101u;
         u32 pos;
         u32 hit = __vcmpeq4(G.x, cc);          // 0xff in every byte that equals c
         if (hit) {
            pos = (u32)(__ffs(hit) - 1) >> 3;
            G.x = mtf_word_hit(G.x, pos, c);
         } else {
            const u32 h1 = __vcmpeq4(G.y, cc), h2 = __vcmpeq4(G.z, cc), h3 = __vcmpeq4(G.w, cc);
            if (h1 | h2 | h3) {
               const u32 c0 = G.x >> 24;
               G.x = (G.x << 8) | c;
               if (h1) { pos = 4 + ((u32)(__ffs(h1) - 1) >> 3); G.y = mtf_word_hit(G.y, pos - 4, c0); }
               else {
                  const u32 c1 = G.y >> 24;
                  G.y = (G.y << 8) | c0;
                  if (h2) { pos = 8 + ((u32)(__ffs(h2) - 1) >> 3); G.z = mtf_word_hit(G.z, pos - 8, c1); }
                  else {
                     const u32 c2 = G.z >> 24;
                     G.z = (G.z << 8) | c1;
                     pos = 12 + ((u32)(__ffs(h3) - 1) >> 3);
                     G.w = mtf_word_hit(G.w, pos - 12, c2);
                  }
               }
            } else {
               // not among the first sixteen: shift the register group, then walk the shared-memory groups
               u32 carry = G.w >> 24;
               G.w = __funnelshift_l(G.z, G.w, 8);
               G.z = __funnelshift_l(G.y, G.z, 8);
               G.y = __funnelshift_l(G.x, G.y, 8);
               G.x = (G.x << 8) | c;
               u32 j = 1;
               for (;;) {
                  uint4 q = my[j * MTF_CTA];
                  const u32 m0 = __vcmpeq4(q.x, cc), m1 = __vcmpeq4(q.y, cc), m2 = __vcmpeq4(q.z, cc), m3 = __vcmpeq4(q.w, cc);
                  if (m0 | m1 | m2 | m3) {
                     if (m0) { pos = (u32)(__ffs(m0) - 1) >> 3; q.x = mtf_word_hit(q.x, pos, carry); }
                     else {
                        const u32 d0 = q.x >> 24;
                        q.x = (q.x << 8) | carry;
                        if (m1) { pos = 4 + ((u32)(__ffs(m1) - 1) >> 3); q.y = mtf_word_hit(q.y, pos - 4, d0); }
                        else {
                           const u32 d1 = q.y >> 24;
                           q.y = (q.y << 8) | d0;
                           if (m2) { pos = 8 + ((u32)(__ffs(m2) - 1) >> 3); q.z = mtf_word_hit(q.z, pos - 8, d1); }
                           else {
                              const u32 d2 = q.z >> 24;
                              q.z = (q.z << 8) | d1;
                              pos = 12 + ((u32)(__ffs(m3) - 1) >> 3);
                              q.w = mtf_word_hit(q.w, pos - 12, d2);
                           }
                        }
                     }
                     my[j * MTF_CTA] = q;
                     pos += 16 * j;
                     break;
                  }
                  const u32 out = q.w >> 24;
                  q.w = __funnelshift_l(q.z, q.w, 8);
                  q.z = __funnelshift_l(q.y, q.z, 8);
                  q.y = __funnelshift_l(q.x, q.y, 8);
                  q.x = (q.x << 8) | carry;
                  my[j * MTF_CTA] = q;
                  carry = out;
                  j++;
               }
            }
         }
         zword |= pos << (8 * k);
         if (pos == 0) run++;
         else {
            if (!seen_nz) { lead = run; seen_nz = true; }
            else if (run) inner += run_digits(run);
            inner += 1;
            run = 0;
         }
      }
      // interior words are owned by this tile; edge words are shared with the neighbours
      const bool full = (w * 4 >= a) && (w * 4 + 4 <= a + size);
      if (full) out32[w] = zword;
      else {
#pragma unroll
         for (int k = 0; k < 4; k++) {
            const i32 i = (i32)(w * 4 + k) - (i32)a;
            if (i >= 0 && i < (i32)size) p.z[start + i] = (u8)(zword >> (8 * k));
         }
      }
   }
   u32* m = p.tmeta + ((size_t)b * p.tiles_max + t) * 4;
   if (!seen_nz) { m[0] = size; m[1] = size | 0x80000000u; m[2] = 0; }
   else { m[0] = lead; m[1] = run; m[2] = inner; }
}


// Zero-run stitching across tiles: one WARP per block.  A range of tiles is summarised as
// (all-zero?, leading zeros, trailing zeros, symbols emitted after the leading run), which
// composes associatively, so the carry into every tile is an exclusive scan.
struct RunSum { u32 allz, lead, trail, inner; };
__device__ __forceinline__ RunSum rs_comb(RunSum A, RunSum B)
{
   RunSum r;
   if (A.allz) { r = B; r.lead = A.lead + B.lead; if (B.allz) r.trail = r.lead; return r; }
   if (B.allz) { r = A; r.trail = A.trail + B.lead; return r; }
   const u32 mid = A.trail + B.lead;
   r.allz = 0; r.lead = A.lead; r.trail = B.trail;
   r.inner = A.inner + B.inner + (mid ? run_digits(mid) : 0);
   return r;
}
__device__ __forceinline__ RunSum rs_load(const u32* m)
{
   RunSum r;
   const u32 tr = m[1];
   r.allz = tr >> 31; r.lead = m[0]; r.trail = tr & 0x7fffffffu; r.inner = m[2];
   return r;
}

constexpr int RS_PER_LANE = 32;            // up to 1024 tiles per block

__global__ void __launch_bounds__(256) k_rle2_scan(S3Params p, u32 nb)
{
   const u32 b = blockIdx.x * 8 + (threadIdx.x >> 5);
   if (b >= nb) return;
   const u32 l = lane_id();
   const u32 n = p.X[b + 1] - p.X[b];
   const u32 ntile = (n + MTF_TILE - 1) / MTF_TILE;
   const u32 per = (ntile + 31) / 32;
   u32* m = p.tmeta + (size_t)b * p.tiles_max * 4;
   u32* cy = p.tcarry + (size_t)b * p.tiles_max;
   const RunSum ident = {1u, 0u, 0u, 0u};
   RunSum mine = ident;
   const u32 lo = l * per;
   for (u32 k = 0; k < per; k++) if (lo + k < ntile) mine = rs_comb(mine, rs_load(m + 4 * (lo + k)));
   // inclusive warp scan
   RunSum inc = mine;
#pragma unroll
   for (int d = 1; d < 32; d <<= 1) {
      RunSum o;
      o.allz = __shfl_up_sync(FULL, inc.allz, d); o.lead = __shfl_up_sync(FULL, inc.lead, d);
      o.trail = __shfl_up_sync(FULL, inc.trail, d); o.inner = __shfl_up_sync(FULL, inc.inner, d);
      if (l >= (u32)d) inc = rs_comb(o, inc);
   }
   RunSum pre;
   pre.allz = __shfl_up_sync(FULL, inc.allz, 1); pre.lead = __shfl_up_sync(FULL, inc.lead, 1);
   pre.trail = __shfl_up_sync(FULL, inc.trail, 1); pre.inner = __shfl_up_sync(FULL, inc.inner, 1);
   if (l == 0) pre = ident;
   for (u32 k = 0; k < per; k++) {
      const u32 t = lo + k;
      if (t >= ntile) break;
      cy[t] = pre.allz ? pre.lead : pre.trail;                                     // zero run entering tile t
      m[4 * t + 3] = pre.allz ? 0u : (pre.inner + (pre.lead ? run_digits(pre.lead) : 0u));   // symbols emitted before it
      pre = rs_comb(pre, rs_load(m + 4 * t));
   }
   if (l == 31) {
      // `inc` of the last lane covers the whole block
      u32 o;
      if (inc.allz) o = run_digits(inc.lead);
      else o = inc.inner + (inc.lead ? run_digits(inc.lead) : 0u) + (inc.trail ? run_digits(inc.trail) : 0u);
      p.nmtf[b] = o + 1;                   // + EOB
   }
}

__global__ void __launch_bounds__(128) k_rle2_emit(S3Params p)
{
   __shared__ u32 hist[BZ_MAX_ALPHA + 6];
   const u32 b = blockIdx.y;
   const u32 t = blockIdx.x * 128 + threadIdx.x;
   const u32 xb = p.X[b], n = p.X[b + 1] - xb;
   for (u32 k = threadIdx.x; k < BZ_MAX_ALPHA; k += 128) hist[k] = 0;
   __syncthreads();
   const u32 ntile = (n + MTF_TILE - 1) / MTF_TILE;
   if (t < ntile) {
      const u32 start = xb + t * MTF_TILE;
      const u32 size = min((u32)MTF_TILE, n - t * MTF_TILE);
      const u32* m = p.tmeta + ((size_t)b * p.tiles_max + t) * 4;
      // symbols are gathered four at a time and stored as one aligned 64-bit word: a thread's 16-bit stores land
      // in 32 different sectors per warp instruction otherwise.  Slots of a boundary word that belong to the
      // neighbouring tile are never touched.
      const size_t base = (size_t)xb + b + m[3];             // this tile's first slot in mtfv
      u16* const mt = p.mtfv;
      u64 acc = 0;
      u32 run = p.tcarry[(size_t)b * p.tiles_max + t];
      u32 o = 0;
      auto put = [&](u32 sym) {
         const size_t gi = base + o;
         const u32 slot = (u32)gi & 3u;
         acc |= (u64)sym << (16 * slot);
         o++;
         if (slot == 3) {
            const size_t g0 = gi - 3;
            if (g0 >= base) *reinterpret_cast<u64*>(mt + g0) = acc;
            else {
#pragma unroll
               for (int j = 0; j < 4; j++) if (g0 + j >= base) mt[g0 + j] = (u16)(acc >> (16 * j));
            }
            acc = 0;
         }
      };
      u32 c0 = 0, c1 = 0, c2 = 0, c3 = 0;          // the hottest symbols are counted in registers
      // the tile is read as aligned 32-bit words (all tiles of a block share the misalignment)
      const u32 a = start & 3u;
      const u32* in32 = reinterpret_cast<const u32*>(p.z + (start - a));
      const u32 nwords = (size + a + 3) >> 2;
      for (u32 w = 0; w < nwords; w++) {
         const u32 word = in32[w];
#pragma unroll
         for (int k = 0; k < 4; k++) {
            const i32 i = (i32)(w * 4 + k) - (i32)a;
            if (i < 0 || i >= (i32)size) continue;
            const u32 v = (word >> (8 * k)) & 0xff;
            if (v == 0) { run++; continue; }
            if (run) {
               u32 zz = run - 1;
               for (;;) { const u32 s = zz & 1; put(s); if (s) c1++; else c0++; if (zz < 2) break; zz = (zz - 2) >> 1; }
               run = 0;
            }
            put(v + 1);
            if (v == 1) c2++; else if (v == 2) c3++; else atomicAdd(&hist[v + 1], 1u);
         }
      }
      if (t == ntile - 1) {
         if (run) {
            u32 zz = run - 1;
            for (;;) { const u32 s = zz & 1; put(s); if (s) c1++; else c0++; if (zz < 2) break; zz = (zz - 2) >> 1; }
         }
         const u32 eob = p.ninuse[b] + 1;
         put(eob);
         atomicAdd(&hist[eob], 1u);
      }
      {
         // the unfinished word
         const size_t gi = base + o;
         const u32 slot = (u32)gi & 3u;
         const size_t g0 = gi - slot;
         for (u32 j = 0; j < slot; j++) if (g0 + j >= base) mt[g0 + j] = (u16)(acc >> (16 * j));
      }
      if (c0) atomicAdd(&hist[0], c0);
      if (c1) atomicAdd(&hist[1], c1);
      if (c2) atomicAdd(&hist[2], c2);
      if (c3) atomicAdd(&hist[3], c3);
   }
   __syncthreads();
   for (u32 k = threadIdx.x; k < BZ_MAX_ALPHA; k += 128)
      if (hist[k]) atomicAdd(&p.mtffreq[(size_t)b * BZ_MAX_ALPHA + k], (i32)hist[k]);
}

int stage3_run(Engine* e, u32 nb, u32 E)
{
   (void)E;
   cudaStream_t st = e->stream;
   const u32 tiles_max = (e->nmax + 16 + MTF_TILE - 1) / MTF_TILE;
   S3Params p;
   p.bwt = e->bwt; p.X = e->bt.X; p.inuse = e->bt.inuse; p.ninuse = e->bt.ninuse;
   p.lists = e->mtf_summary; p.tilecnt = e->mtf_tilecnt; p.mode = e->mtf_mode;
   // (a warp-cooperative encoder for high-entropy blocks was measured in round 1: the thread-per-tile kernel wins even on
   // uniform random bytes, 120 vs 137 ms/GB; it is no longer in the tree and no block is ever marked for it)
   p.warp_threshold = 1000u; p.tmeta = e->mtf_tilemeta;
   p.tcarry = e->mtf_tilemeta + (size_t)e->blk_cap * tiles_max * 4;
   p.z = e->z; p.mtfv = e->mtfv; p.nmtf = e->bt.nmtf; p.mtffreq = e->bt.mtffreq; p.tiles_max = tiles_max;
   BZ_CUDA(e, cudaMemsetAsync(e->bt.mtffreq, 0, sizeof(i32) * BZ_MAX_ALPHA * nb, st));
   const dim3 gw((tiles_max + 7) / 8, nb);
   const dim3 gt((tiles_max + 127) / 128, nb);
   k_mtf_summary<<<gw, 256, 0, st>>>(p);                BZ_KCHECK(e);
   k_mtf_lists<<<(nb + ML_WARPS - 1) / ML_WARPS, ML_WARPS * 32, 0, st>>>(p, nb);   BZ_KCHECK(e);
   k_mtf_encode<<<gt, MTF_CTA, 0, st>>>(p);             BZ_KCHECK(e);
   k_rle2_scan<<<(nb + 7) / 8, 256, 0, st>>>(p, nb);    BZ_KCHECK(e);
   k_rle2_emit<<<gt, 128, 0, st>>>(p);                  BZ_KCHECK(e);
   return 0;
}

} // namespace bz
