// stage2_bwt.cu -- S2: Burrows-Wheeler transform of every block of a window.
//
// Replaces BZ2_blockSort and everything under it (reference blocksort.c:1534-1545,
// divsufsort :1503-1514, sort_typeBstar :1314-1437, ss_* :83-664, tr_* / ls_* :669-1309,
// construct_SA :1439-1501).  Same contract: rotations of the block in lexicographic
// order; output byte k is the byte preceding rotation k; origPtr is the rank of
// rotation 0.  The method is not the reference's (induced sorting is serial):
//
//   1. bigram bucket sort   every rotation is counted into one of 65536 buckets by
//                           its first two bytes (global atomics), buckets are
//                           scanned, rotations scattered -> h-order with h = 2.
//   2. prefix doubling      each unresolved bucket ("segment") is sorted by the rank
//                           of the rotation h further on (cyclic), which doubles the
//                           sorted depth; equal keys stay one segment.  Segments are
//                           handled by size class, all blocks of the window at once:
//                             2..32       sub-warp bitonic network in registers
//                             33..4096    one CTA, bitonic sort in shared memory
//                             > 4096      one CTA, stable 8-bit LSD radix passes
//                                         staged through shared-memory histograms
//                           New ranks are written by sorted position and applied in
//                           a second kernel so that a round only ever reads h-order
//                           ranks.
//   3. last column          bwt[k] = T[(sa[k]-1) mod n]; origPtr = k with sa[k]==0.
//
// A segment that survives to depth >= n holds equal rotations: the block is an exact
// power u^q; q is recorded in power_q[b] (the BWT bytes do not depend on their order).
#include "engine.h"

namespace bz {

struct ListsDev {
   u32* small_items[N_SMALL_CLASSES];
   u64* big_items[3];
   u32* counts;
   u32  small_cap[N_SMALL_CLASSES];
   u32  big_cap[3];
   u32* overflow;
};

struct S2Params {
   const u8* T;
   const u32* X;
   u32 nb;
   u32* sa; u32* rank; u32* nrank;
   u32* keyA; u32* keyB; u32* idxB;
   u32* hist;
   const u32* blockmap;
   u32* power_q;
   u8* inuse; u32* ninuse;
};

__device__ __forceinline__ int seg_class(u32 len)
{
   if (len <= 32) return 31 - __clz(len - 1);
   if (len <= MED1_MAX) return CLS_MED1;
   if (len <= MED2_MAX) return CLS_MED2;
   return CLS_LARGE;
}

// All 32 lanes of the warp must call this together.
__device__ __forceinline__ void push_seg(const ListsDev& L, bool valid, u32 pos, u32 blk, u32 len)
{
   const int cls = valid ? seg_class(len) : -1;
   const u32 m = __match_any_sync(FULL, cls);
   const u32 leader = __ffs(m) - 1;
   u32 base = 0;
   if (valid && lane_id() == leader) base = atomicAdd(&L.counts[cls], __popc(m));
   base = __shfl_sync(FULL, base, leader);
   if (valid) {
      const u32 slot = base + __popc(m & lanemask_lt());
      if (cls < N_SMALL_CLASSES) {
         if (slot < L.small_cap[cls]) L.small_items[cls][slot] = pos | ((len - 1) << 27);
         else atomicOr(L.overflow, 1u);
      } else {
         const int c = cls - N_SMALL_CLASSES;
         if (slot < L.big_cap[c]) L.big_items[c][slot] = ((u64)pos << 32) | ((u64)blk << 20) | (u64)len;
         else atomicOr(L.overflow, 1u);
      }
   }
}

__device__ __forceinline__ u32 block_of(const S2Params& p, u32 pos)
{
   u32 b = p.blockmap[pos >> 12];
   while (b + 1 < p.nb && pos >= p.X[b + 1]) b++;
   return b;
}

__global__ void k_blockmap(const u32* X, u32 nb, u32* blockmap, u32 nchunks)
{
   u32 c = blockIdx.x * blockDim.x + threadIdx.x;
   if (c >= nchunks) return;
   u32 pos = c << 12;
   u32 lo = 0, hi = nb - 1;
   while (lo < hi) { u32 mid = (lo + hi + 1) >> 1; if (X[mid] <= pos) lo = mid; else hi = mid - 1; }
   blockmap[c] = lo;
}

// ---- 1. bigram bucket sort -----------------------------------------------------
constexpr int BG_THREADS = 256;
constexpr int BG_ITEMS = 16;
constexpr int BG_TILE = BG_THREADS * BG_ITEMS;

enum { BG_HIST = 0, BG_RANK = 1, BG_SCATTER = 2 };

template <int MODE>
__global__ void __launch_bounds__(BG_THREADS) k_bigram(S2Params p)
{
   const u32 b = blockIdx.y;
   const u32 xb = p.X[b], n = p.X[b + 1] - xb;
   const u32 t0 = blockIdx.x * BG_TILE;
   if (t0 >= n) return;
   const u8* T = p.T + xb;
   u32* hist = p.hist + (size_t)b * 65536;
#pragma unroll 4
   for (int k = 0; k < BG_ITEMS; k++) {
      const u32 i = t0 + k * BG_THREADS + threadIdx.x;
      if (i < n) {
         const u32 c0 = T[i];
         const u32 c1 = T[(i + 1 == n) ? 0 : i + 1];
         const u32 bin = (c0 << 8) | c1;
         if (MODE == BG_HIST) atomicAdd(&hist[bin], 1u);
         else if (MODE == BG_RANK) p.rank[xb + i] = hist[bin];
         else { const u32 pos = atomicAdd(&hist[bin], 1u); p.sa[xb + pos] = i; }
      }
   }
}

// exclusive scan of the 65536 bucket counts of one block; also the in-use byte map
__global__ void __launch_bounds__(1024) k_bigram_scan(S2Params p)
{
   __shared__ u32 ssm[34];
   __shared__ u32 rows[256];
   const u32 b = blockIdx.x;
   u32* hist = p.hist + (size_t)b * 65536;
   if (threadIdx.x < 256) rows[threadIdx.x] = 0;
   u32 v[64];
   u32 sum = 0;
   uint4* h4 = reinterpret_cast<uint4*>(hist + threadIdx.x * 64);
#pragma unroll
   for (int k = 0; k < 16; k++) {
      uint4 q = h4[k];
      v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
      sum += q.x + q.y + q.z + q.w;
   }
   u32 ex = block_excl_sum<1024>(sum, ssm, nullptr);   // syncs inside also publish rows[] zeroing
   if (sum) atomicAdd(&rows[threadIdx.x >> 2], sum);
#pragma unroll
   for (int k = 0; k < 16; k++) {
      uint4 q;
      q.x = ex; ex += v[4 * k];
      q.y = ex; ex += v[4 * k + 1];
      q.z = ex; ex += v[4 * k + 2];
      q.w = ex; ex += v[4 * k + 3];
      h4[k] = q;
   }
   __syncthreads();
   const u32 used = (threadIdx.x < 256 && rows[threadIdx.x]) ? 1u : 0u;
   if (threadIdx.x < 256) p.inuse[(size_t)b * 256 + threadIdx.x] = (u8)used;
   const u32 cnt = __syncthreads_count(used);
   if (threadIdx.x == 0) p.ninuse[b] = cnt;
}

// after the scatter hist[bin] is the END of bucket `bin`; emit the initial segments
__global__ void __launch_bounds__(256) k_seg_init(S2Params p, ListsDev L)
{
   const u32 b = blockIdx.y;
   const u32 xb = p.X[b], n = p.X[b + 1] - xb;
   const u32* hist = p.hist + (size_t)b * 65536;
   const u32 bin = blockIdx.x * 256 + threadIdx.x;
   const u32 end = hist[bin];
   const u32 start = bin ? hist[bin - 1] : 0;
   const u32 len = end - start;
   const bool multi = len >= 2;
   const bool deep = (2u >= n);                       // depth 2 already covers the whole rotation
   if (multi && deep) atomicMax(&p.power_q[b], len);
   push_seg(L, multi && !deep, xb + start, b, len);
}

// ---- 2a. small segments: sub-warp bitonic network ------------------------------------
template <int LANES>
__global__ void __launch_bounds__(256) k_refine_small(S2Params p, ListsDev Lout, const u32* items, u32 count, u32 h)
{
   const u32 gid = blockIdx.x * blockDim.x + threadIdx.x;
   const u32 seg = gid / LANES;
   const u32 sub = gid % LANES;
   const bool vseg = seg < count;
   const u32 entry = vseg ? items[seg] : 0;
   const u32 pos = entry & 0x7ffffffu;
   const u32 len = (entry >> 27) + 1;
   u32 b = 0, xb = 0, n = 1;
   if (vseg) { b = block_of(p, pos); xb = p.X[b]; n = p.X[b + 1] - xb; }
   const bool active = vseg && sub < len;
   u32 idx = 0, key = 0xffffffffu;
   if (active) {
      idx = p.sa[pos + sub];
      u32 t = idx + h; if (t >= n) t -= n;
      key = p.rank[xb + t];
   }
#pragma unroll
   for (int k = 2; k <= LANES; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
         const u32 ok = __shfl_xor_sync(FULL, key, j);
         const u32 oi = __shfl_xor_sync(FULL, idx, j);
         const bool asc = ((sub & k) == 0);
         const bool low = ((sub & j) == 0);
         const bool take = (asc == low) ? (ok < key) : (ok > key);
         if (take) { key = ok; idx = oi; }
      }
   }
   const u32 pk = __shfl_up_sync(FULL, key, 1);
   const u32 nk = __shfl_down_sync(FULL, key, 1);
   const bool head = active && (sub == 0 || pk != key);
   const u32 bal = __ballot_sync(FULL, head);
   const u32 sh = lane_id() & ~(u32)(LANES - 1);
   const u32 mask = (LANES == 32) ? bal : ((bal >> sh) & ((1u << (LANES & 31)) - 1u));
   const u32 below = mask & ((2u << sub) - 1u);
   const u32 gstart = active ? (31 - __clz(below)) : 0;
   const bool is_end = active && (sub == len - 1 || nk != key);
   const u32 size = sub - gstart + 1;
   if (active) {
      p.sa[pos + sub] = idx;
      p.nrank[pos + sub] = (pos - xb) + gstart;
   }
   const bool multi = is_end && size >= 2;
   const bool deep = (2u * h >= n);
   if (multi && deep) atomicMax(&p.power_q[b], size);
   push_seg(Lout, multi && !deep, pos + gstart, b, size);
}

template <int LANES>
__global__ void __launch_bounds__(256) k_apply_small(S2Params p, const u32* items, u32 count)
{
   const u32 gid = blockIdx.x * blockDim.x + threadIdx.x;
   const u32 seg = gid / LANES;
   const u32 sub = gid % LANES;
   if (seg >= count) return;
   const u32 entry = items[seg];
   const u32 pos = entry & 0x7ffffffu;
   const u32 len = (entry >> 27) + 1;
   if (sub >= len) return;
   const u32 b = block_of(p, pos);
   const u32 xb = p.X[b];
   p.rank[xb + p.sa[pos + sub]] = p.nrank[pos + sub];
}

// ---- 2b. medium segments: one CTA, bitonic sort in shared memory ---------------------
template <int CAP, int THREADS>
__global__ void __launch_bounds__(THREADS) k_refine_medium(S2Params p, ListsDev Lout, const u64* items, u32 h)
{
   constexpr int ITEMS = CAP / THREADS;
   __shared__ u32 skey[CAP];
   __shared__ u32 sidx[CAP];
   __shared__ u32 ssm[34];
   const u64 entry = items[blockIdx.x];
   const u32 pos = (u32)(entry >> 32);
   const u32 b = (u32)(entry >> 20) & 0xfffu;
   const u32 len = (u32)entry & 0xfffffu;
   const u32 xb = p.X[b], n = p.X[b + 1] - xb;
   u32 n2 = 64;
   while (n2 < len) n2 <<= 1;
   for (u32 i = threadIdx.x; i < n2; i += THREADS) {
      u32 key = 0xffffffffu, idx = 0;
      if (i < len) {
         idx = p.sa[pos + i];
         u32 t = idx + h; if (t >= n) t -= n;
         key = p.rank[xb + t];
      }
      skey[i] = key; sidx[i] = idx;
   }
   __syncthreads();
   for (u32 k = 2; k <= n2; k <<= 1) {
      for (u32 j = k >> 1; j > 0; j >>= 1) {
         for (u32 t = threadIdx.x; t < (n2 >> 1); t += THREADS) {
            const u32 lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            const u32 hi = lo | j;
            const bool asc = ((lo & k) == 0);
            const u32 a = skey[lo], c = skey[hi];
            if ((a > c) == asc && a != c) {
               skey[lo] = c; skey[hi] = a;
               const u32 ia = sidx[lo]; sidx[lo] = sidx[hi]; sidx[hi] = ia;
            }
         }
         __syncthreads();
      }
   }
   // group starts: inclusive max-scan of (head ? i+1 : 0)
   const u32 base = threadIdx.x * ITEMS;
   u32 last = 0;
   u32 gs[ITEMS];
#pragma unroll
   for (int k = 0; k < ITEMS; k++) {
      const u32 i = base + k;
      if (i < len) {
         const bool head = (i == 0) || (skey[i] != skey[i - 1]);
         if (head) last = i + 1;
      }
      gs[k] = last;
   }
   const u32 incl = block_incl_max<THREADS>(last, ssm);
   // exclusive value for this thread = max over previous threads; recover from warp shuffle of inclusive
   u32 prev = __shfl_up_sync(FULL, incl, 1);
   if (lane_id() == 0) {
      // previous warp's inclusive max lives in ssm[w] (exclusive prefix of warps)
      prev = ssm[threadIdx.x >> 5];
   }
   const bool deep = (2u * h >= n);
#pragma unroll
   for (int k = 0; k < ITEMS; k++) {
      const u32 i = base + k;
      const bool in = i < len;
      u32 g = gs[k] ? gs[k] : prev;          // 1-based head index
      g = g ? g - 1 : 0;
      bool is_end = false;
      if (in) {
         p.sa[pos + i] = sidx[i];
         p.nrank[pos + i] = (pos - xb) + g;
         is_end = (i == len - 1) || (skey[i + 1] != skey[i]);
      }
      const u32 size = i - g + 1;
      const bool multi = in && is_end && size >= 2;
      if (multi && deep) atomicMax(&p.power_q[b], size);
      push_seg(Lout, multi && !deep, pos + g, b, size);
   }
}

// ---- 2c. large segments: one CTA, stable LSD radix passes ----------------------------
constexpr int LG_THREADS = 1024;
constexpr int LG_ITEMS = 4;
constexpr int LG_TILE = LG_THREADS * LG_ITEMS;
constexpr int LG_WARPS = LG_THREADS / 32;

__global__ void __launch_bounds__(LG_THREADS) k_refine_large(S2Params p, ListsDev Lout, const u64* items, u32 h)
{
   __shared__ u32 whist[LG_WARPS][256];
   __shared__ u32 binbase[3][256];
   __shared__ u32 ssm[34];
   __shared__ u32 s_carry;
   const u64 entry = items[blockIdx.x];
   const u32 pos = (u32)(entry >> 32);
   const u32 b = (u32)(entry >> 20) & 0xfffu;
   const u32 len = (u32)entry & 0xfffffu;
   const u32 xb = p.X[b], n = p.X[b + 1] - xb;
   const int npass = (n > 65536u) ? 3 : 2;
   const u32 w = threadIdx.x >> 5, l = lane_id();

   for (u32 i = threadIdx.x; i < 3 * 256; i += LG_THREADS) (&binbase[0][0])[i] = 0;
   __syncthreads();
   // phase A: gather keys, digit histograms
   for (u32 i = threadIdx.x; i < len; i += LG_THREADS) {
      const u32 idx = p.sa[pos + i];
      u32 t = idx + h; if (t >= n) t -= n;
      const u32 key = p.rank[xb + t];
      p.keyA[pos + i] = key;
      atomicAdd(&binbase[0][key & 255], 1u);
      atomicAdd(&binbase[1][(key >> 8) & 255], 1u);
      if (npass == 3) atomicAdd(&binbase[2][(key >> 16) & 255], 1u);
   }
   __syncthreads();
   if (w < 3) {
      // exclusive scan of 256 bins by one warp: 8 per lane
      u32 v[8], s = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) { v[k] = binbase[w][l * 8 + k]; s += v[k]; }
      u32 ex = warp_incl_sum(s) - s;
#pragma unroll
      for (int k = 0; k < 8; k++) { binbase[w][l * 8 + k] = ex; ex += v[k]; }
   }
   __syncthreads();

   // phase B: passes
   for (int pass = 0; pass < npass; pass++) {
      const u32* ksrc = (pass & 1) ? p.keyB : p.keyA;
      const u32* isrc = (pass & 1) ? p.idxB : p.sa;
      u32* kdst = (pass & 1) ? p.keyA : p.keyB;
      u32* idst = (pass & 1) ? p.sa : p.idxB;
      const int shift = pass * 8;
      for (u32 tb = 0; tb < len; tb += LG_TILE) {
#pragma unroll
         for (int k = 0; k < 8; k++) whist[w][l * 8 + k] = 0;
         __syncwarp();
         u32 key[LG_ITEMS], idx[LG_ITEMS], rk[LG_ITEMS];
#pragma unroll
         for (int k = 0; k < LG_ITEMS; k++) {
            const u32 i = tb + w * (32 * LG_ITEMS) + k * 32 + l;
            const bool valid = i < len;
            key[k] = valid ? ksrc[pos + i] : 0;
            idx[k] = valid ? isrc[pos + i] : 0;
            const u32 d = (key[k] >> shift) & 255;
            const u32 m = __match_any_sync(FULL, valid ? d : 0x1000u);
            const u32 old = whist[w][d];
            __syncwarp();
            if (valid && l == (u32)(__ffs(m) - 1)) whist[w][d] = old + __popc(m);
            __syncwarp();
            rk[k] = old + __popc(m & lanemask_lt());
         }
         __syncthreads();
         if (threadIdx.x < 256) {
            u32 run = binbase[pass][threadIdx.x];
#pragma unroll 8
            for (int ww = 0; ww < LG_WARPS; ww++) {
               const u32 c = whist[ww][threadIdx.x];
               whist[ww][threadIdx.x] = run;
               run += c;
            }
            binbase[pass][threadIdx.x] = run;
         }
         __syncthreads();
#pragma unroll
         for (int k = 0; k < LG_ITEMS; k++) {
            const u32 i = tb + w * (32 * LG_ITEMS) + k * 32 + l;
            if (i < len) {
               const u32 d = (key[k] >> shift) & 255;
               const u32 dst = whist[w][d] + rk[k];
               kdst[pos + dst] = key[k];
               idst[pos + dst] = idx[k];
            }
         }
         __syncthreads();
      }
      __threadfence_block();
   }
   // phase C: group boundaries, new ranks, next-round segments
   const u32* kfin = (npass & 1) ? p.keyB : p.keyA;
   const u32* ifin = (npass & 1) ? p.idxB : p.sa;
   const bool deep = (2u * h >= n);
   if (threadIdx.x == 0) s_carry = 0;
   __syncthreads();
   for (u32 tb = 0; tb < len; tb += LG_TILE) {
      const u32 base = tb + threadIdx.x * LG_ITEMS;
      u32 kk[LG_ITEMS + 2];
#pragma unroll
      for (int k = 0; k < LG_ITEMS + 2; k++) {
         const i64 i = (i64)base + k - 1;
         kk[k] = (i >= 0 && i < (i64)len) ? kfin[pos + (u32)i] : 0xffffffffu;
      }
      u32 last = 0, gs[LG_ITEMS];
#pragma unroll
      for (int k = 0; k < LG_ITEMS; k++) {
         const u32 i = base + k;
         if (i < len) { if (i == 0 || kk[k + 1] != kk[k]) last = i + 1; }
         gs[k] = last;
      }
      const u32 incl = block_incl_max<LG_THREADS>(last, ssm);
      u32 prev = __shfl_up_sync(FULL, incl, 1);
      if (l == 0) prev = ssm[w];
      const u32 carry = s_carry;
      if (prev == 0) prev = carry;
#pragma unroll
      for (int k = 0; k < LG_ITEMS; k++) {
         const u32 i = base + k;
         const bool in = i < len;
         u32 g = gs[k] ? gs[k] : prev;
         g = g ? g - 1 : 0;
         bool is_end = false;
         if (in) {
            if (npass & 1) p.sa[pos + i] = ifin[pos + i];
            p.nrank[pos + i] = (pos - xb) + g;
            is_end = (i == len - 1) || (kk[k + 2] != kk[k + 1]);
         }
         const u32 size = i - g + 1;
         const bool multi = in && is_end && size >= 2;
         if (multi && deep) atomicMax(&p.power_q[b], size);
         push_seg(Lout, multi && !deep, pos + g, b, size);
      }
      __syncthreads();
      if (threadIdx.x == LG_THREADS - 1) s_carry = incl ? incl : carry;
      __syncthreads();
   }
}

__global__ void __launch_bounds__(256) k_apply_big(S2Params p, const u64* items)
{
   const u64 entry = items[blockIdx.x];
   const u32 pos = (u32)(entry >> 32);
   const u32 b = (u32)(entry >> 20) & 0xfffu;
   const u32 len = (u32)entry & 0xfffffu;
   const u32 xb = p.X[b];
   for (u32 i = threadIdx.x; i < len; i += 256) p.rank[xb + p.sa[pos + i]] = p.nrank[pos + i];
}

// ---- 3. last column -------------------------------------------------------------------
__global__ void __launch_bounds__(BG_THREADS) k_bwt_out(S2Params p, u8* bwt, u32* origptr)
{
   const u32 b = blockIdx.y;
   const u32 xb = p.X[b], n = p.X[b + 1] - xb;
   const u32 t0 = blockIdx.x * BG_TILE;
   if (t0 >= n) return;
   const u8* T = p.T + xb;
   // rank of rotation 0: its final position, or the start of its tie group for exact powers
   if (blockIdx.x == 0 && threadIdx.x == 0) origptr[b] = p.rank[xb];
#pragma unroll 4
   for (int k = 0; k < BG_ITEMS; k++) {
      const u32 i = t0 + k * BG_THREADS + threadIdx.x;
      if (i < n) {
         const u32 s = p.sa[xb + i];
         bwt[xb + i] = T[s ? s - 1 : n - 1];
      }
   }
}

// --------------------------------------------------------------------------------------
static ListsDev lists_dev(Engine* e, int which)
{
   ListsDev L;
   for (int c = 0; c < N_SMALL_CLASSES; c++) { L.small_items[c] = e->lists.small_items[which][c]; L.small_cap[c] = e->lists.small_cap[c]; }
   for (int c = 0; c < 3; c++) { L.big_items[c] = e->lists.big_items[which][c]; L.big_cap[c] = e->lists.big_cap[c]; }
   L.counts = e->lists.counts[which];
   L.overflow = e->s1_scalars + 4;
   return L;
}

template <int LANES>
static void launch_small(Engine* e, const S2Params& p, const ListsDev& Lout, const u32* items, u32 count, u32 h)
{
   const u64 threads = (u64)count * LANES;
   const u32 grid = (u32)((threads + 255) / 256);
   k_refine_small<LANES><<<grid, 256, 0, e->stream>>>(p, Lout, items, count, h);
}
template <int LANES>
static void launch_apply_small(Engine* e, const S2Params& p, const u32* items, u32 count)
{
   const u64 threads = (u64)count * LANES;
   const u32 grid = (u32)((threads + 255) / 256);
   k_apply_small<LANES><<<grid, 256, 0, e->stream>>>(p, items, count);
}

int stage2_run(Engine* e, u32 nb, u32 E)
{
   cudaStream_t st = e->stream;
   S2Params p;
   p.T = e->enc; p.X = e->bt.X; p.nb = nb;
   p.sa = e->sa; p.rank = e->rank; p.nrank = e->nrank;
   p.keyA = e->keyA; p.keyB = e->keyB; p.idxB = e->idxB;
   p.hist = e->hist; p.blockmap = e->blockmap;
   p.power_q = e->bt.power_q; p.inuse = e->bt.inuse; p.ninuse = e->bt.ninuse;

   const u32 nchunks = (E >> 12) + 1;
   k_blockmap<<<(nchunks + 255) / 256, 256, 0, st>>>(e->bt.X, nb, e->blockmap, nchunks);     BZ_KCHECK(e);
   BZ_CUDA(e, cudaMemsetAsync(e->hist, 0, (size_t)nb * 65536 * sizeof(u32), st));
   BZ_CUDA(e, cudaMemsetAsync(e->bt.power_q, 0, sizeof(u32) * nb, st));
   BZ_CUDA(e, cudaMemsetAsync(e->lists.counts[0], 0, sizeof(u32) * N_CLASSES, st));
   BZ_CUDA(e, cudaMemsetAsync(e->s1_scalars + 4, 0, sizeof(u32), st));
   const u32 max_n = e->nmax + 16;
   const dim3 gtiles((max_n + BG_TILE - 1) / BG_TILE, nb);
   k_bigram<BG_HIST><<<gtiles, BG_THREADS, 0, st>>>(p);                                      BZ_KCHECK(e);
   k_bigram_scan<<<nb, 1024, 0, st>>>(p);                                                    BZ_KCHECK(e);
   k_bigram<BG_RANK><<<gtiles, BG_THREADS, 0, st>>>(p);                                      BZ_KCHECK(e);
   k_bigram<BG_SCATTER><<<gtiles, BG_THREADS, 0, st>>>(p);                                   BZ_KCHECK(e);
   k_seg_init<<<dim3(256, nb), 256, 0, st>>>(p, lists_dev(e, 0));                            BZ_KCHECK(e);

   int cur = 0;
   for (u32 h = 2; ; h *= 2) {
      BZ_CUDA(e, cudaMemcpyAsync(e->h_counts, e->lists.counts[cur], sizeof(u32) * (N_CLASSES), cudaMemcpyDeviceToHost, st));
      BZ_CUDA(e, cudaMemcpyAsync(e->h_scalars + 4, e->s1_scalars + 4, sizeof(u32), cudaMemcpyDeviceToHost, st));
      BZ_CUDA(e, cudaStreamSynchronize(st));
      if (e->h_scalars[4]) { snprintf(e->err, sizeof e->err, "segment worklist overflow"); return -4; }
      u32 cnt[N_CLASSES]; u64 total = 0;
      for (int c = 0; c < N_CLASSES; c++) { cnt[c] = e->h_counts[c]; total += cnt[c]; }
      if (total == 0) break;
      if (h >= (1u << 21)) { snprintf(e->err, sizeof e->err, "prefix doubling did not terminate"); return -5; }
      const int nxt = cur ^ 1;
      BZ_CUDA(e, cudaMemsetAsync(e->lists.counts[nxt], 0, sizeof(u32) * N_CLASSES, st));
      ListsDev Lout = lists_dev(e, nxt);
      u32** si = e->lists.small_items[cur];
      u64** bi = e->lists.big_items[cur];
      if (cnt[CLS_LARGE]) k_refine_large<<<cnt[CLS_LARGE], LG_THREADS, 0, st>>>(p, Lout, bi[2], h);
      if (cnt[CLS_MED2])  k_refine_medium<4096, 512><<<cnt[CLS_MED2], 512, 0, st>>>(p, Lout, bi[1], h);
      if (cnt[CLS_MED1])  k_refine_medium<512, 128><<<cnt[CLS_MED1], 128, 0, st>>>(p, Lout, bi[0], h);
      if (cnt[4]) launch_small<32>(e, p, Lout, si[4], cnt[4], h);
      if (cnt[3]) launch_small<16>(e, p, Lout, si[3], cnt[3], h);
      if (cnt[2]) launch_small<8>(e, p, Lout, si[2], cnt[2], h);
      if (cnt[1]) launch_small<4>(e, p, Lout, si[1], cnt[1], h);
      if (cnt[0]) launch_small<2>(e, p, Lout, si[0], cnt[0], h);
      u32 nl = 0;
      for (int c = 0; c < N_CLASSES; c++) nl += cnt[c] ? 1u : 0u;
      e->launches += nl - 1;
      e->bwt_rounds++;
      BZ_KCHECK(e);
      if (cnt[CLS_LARGE]) k_apply_big<<<cnt[CLS_LARGE], 256, 0, st>>>(p, bi[2]);
      if (cnt[CLS_MED2])  k_apply_big<<<cnt[CLS_MED2], 256, 0, st>>>(p, bi[1]);
      if (cnt[CLS_MED1])  k_apply_big<<<cnt[CLS_MED1], 256, 0, st>>>(p, bi[0]);
      if (cnt[4]) launch_apply_small<32>(e, p, si[4], cnt[4]);
      if (cnt[3]) launch_apply_small<16>(e, p, si[3], cnt[3]);
      if (cnt[2]) launch_apply_small<8>(e, p, si[2], cnt[2]);
      if (cnt[1]) launch_apply_small<4>(e, p, si[1], cnt[1]);
      if (cnt[0]) launch_apply_small<2>(e, p, si[0], cnt[0]);
      e->launches += nl - 1;
      BZ_KCHECK(e);
      cur = nxt;
   }
   k_bwt_out<<<gtiles, BG_THREADS, 0, st>>>(p, e->bwt, e->bt.origptr);                       BZ_KCHECK(e);
   return 0;
}

} // namespace bz
