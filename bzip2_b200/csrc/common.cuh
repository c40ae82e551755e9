// common.cuh -- shared device helpers for the sm_100a bzip2 compression kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace bz {

typedef uint8_t  u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t  i32;
typedef int64_t  i64;

constexpr int BZ_G_SIZE      = 50;    // symbols per selector group   (bzlib_private.h:150)
constexpr int BZ_N_ITERS     = 4;     // table refinement passes      (bzlib_private.h:151)
constexpr int BZ_MAX_ALPHA   = 258;   // bzlib_private.h:139
constexpr int BZ_MAX_CODELEN = 17;    // bzlib_private.h:140
constexpr int BZ_MAX_SEL     = 18002; // bzlib_private.h:152

constexpr u32 FULL = 0xffffffffu;

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt() { u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// warp inclusive sum
__device__ __forceinline__ u32 warp_incl_sum(u32 v)
{
#pragma unroll
   for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(FULL, v, d); if (lane_id() >= (u32)d) v += t; }
   return v;
}
__device__ __forceinline__ u32 warp_sum(u32 v)
{
#pragma unroll
   for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
   return v;
}
__device__ __forceinline__ u32 warp_incl_max(u32 v)
{
#pragma unroll
   for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(FULL, v, d); if (lane_id() >= (u32)d) v = max(v, t); }
   return v;
}

// Block-wide exclusive sum over one value per thread.  `sm` needs 33 u32.
// Returns the exclusive prefix; *total (optional) receives the block total.
template <int THREADS>
__device__ __forceinline__ u32 block_excl_sum(u32 v, u32* sm, u32* total)
{
   constexpr int NW = THREADS / 32;
   u32 inc = warp_incl_sum(v);
   u32 w = threadIdx.x >> 5, l = lane_id();
   __syncthreads();
   if (l == 31) sm[w] = inc;
   __syncthreads();
   if (w == 0) {
      u32 x = (l < NW) ? sm[l] : 0;
      u32 xi = warp_incl_sum(x);
      if (l < NW) sm[l] = xi - x;
      if (l == 31) sm[32] = xi;
   }
   __syncthreads();
   u32 r = sm[w] + inc - v;
   if (total) *total = sm[32];
   return r;
}

// Block-wide inclusive max over one value per thread.  `sm` needs 33 u32.
template <int THREADS>
__device__ __forceinline__ u32 block_incl_max(u32 v, u32* sm)
{
   constexpr int NW = THREADS / 32;
   u32 inc = warp_incl_max(v);
   u32 w = threadIdx.x >> 5, l = lane_id();
   __syncthreads();
   if (l == 31) sm[w] = inc;
   __syncthreads();
   if (w == 0) {
      u32 x = (l < NW) ? sm[l] : 0;
      u32 xi = warp_incl_max(x);
      // exclusive: value of previous warps
      u32 prev = __shfl_up_sync(FULL, xi, 1);
      if (l == 0) prev = 0;
      if (l < NW) sm[l] = prev;
   }
   __syncthreads();
   return max(inc, sm[w]);
}

} // namespace bz
