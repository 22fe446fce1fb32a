// dist.cuh — device helpers shared by the multi-GPU stage kernels (dist.cu: records exchanged
// by the caller's collectives; dist_peer.cu: records written straight into peer memory).
#pragma once
#include "tree.h"

namespace stb {

constexpr int DP_THREADS = 256;
constexpr int DP_ITEMS = 4;
constexpr int DP_TILE = DP_THREADS * DP_ITEMS;
constexpr int DP_WARPS = DP_THREADS / 32;
constexpr int MAX_WORLD = 16;

__device__ __forceinline__ uint32_t owner_of(unsigned long long key, int world) {
  unsigned long long k = key * 0x9E3779B97F4A7C15ull;
  k ^= k >> 29;
  k *= 0xBF58476D1CE4E5B9ull;
  return __umulhi((uint32_t)(k >> 32), (uint32_t)world);
}

// (key, flags) of local position i.  KIND 0: packed leaf; KIND 1: pair of child pointers.
template <int KIND>
__device__ __forceinline__ void produce(const void* items, uint64_t n_items, int S, uint64_t i, unsigned long long& key,
                                        uint32_t& flags) {
  if (KIND == 0) {
    key = canonical_leaf(__ldg(reinterpret_cast<const unsigned long long*>(items) + i), S, flags);
  } else {
    const uint32_t* cur = reinterpret_cast<const uint32_t*>(items);
    uint32_t l, r;
    if (2 * i + 1 < n_items) {
      const uint2 pr = __ldg(reinterpret_cast<const uint2*>(cur) + i);
      l = pr.x;
      r = pr.y;
    } else {
      l = cur[2 * i];
      r = PTR_NULL;
    }
    uint32_t cl, cr;
    canonical_node(l, r, cl, cr, flags);
    key = ((unsigned long long)cl << 32) | cr;
  }
}

// one CTA per owner: exclusive scan of its row + row total
static __global__ void __launch_bounds__(1024) rowscan_kernel(uint32_t* __restrict__ hist, uint32_t nblocks, uint32_t* __restrict__ row_total) {
  __shared__ uint32_t warp_sum[32];
  __shared__ uint32_t carry_s;
  uint32_t* row = hist + (size_t)blockIdx.x * nblocks;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nblocks; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < nblocks ? row[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const uint32_t before = carry_s + (warp ? warp_sum[warp - 1] : 0u) + x - v;
    if (i < nblocks) row[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) row_total[blockIdx.x] = carry_s;
}

constexpr uint32_t OWNER_SINGLETON = 0xffffffffu;  // slot marker: certified singleton, never in the table

__device__ __forceinline__ void owner_filter_cell(unsigned long long key, uint32_t log2_bits, uint32_t& word, uint32_t& bit) {
  const uint32_t h = (uint32_t)mix64(key) >> (32 - log2_bits);
  word = h >> 5;
  bit = 1u << (h & 31);
}

__device__ __forceinline__ uint32_t rank_of(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ word_prefix,
                                            uint64_t n_bits, uint64_t total_word, uint64_t q) {
  if (q >= n_bits) return word_prefix[total_word];
  return __ldg(word_prefix + (q >> 5)) + __popc(__ldg(bitmap + (q >> 5)) & ((1u << (q & 31)) - 1u));
}

}  // namespace stb
