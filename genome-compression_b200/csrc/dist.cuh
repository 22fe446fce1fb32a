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

// Position order: first occurrences append their item and get their pointer.
template <int KIND>
__global__ void __launch_bounds__(256)
finish_first_kernel(const void* __restrict__ items, uint64_t n_items, uint64_t n_pos, int S, uint64_t gpos0,
                    const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ word_prefix, uint64_t n_bits,
                    uint64_t n_words, uint32_t* __restrict__ pointers, void* __restrict__ slice,
                    uint32_t* __restrict__ base_count) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  const uint32_t base = rank_of(bitmap, word_prefix, n_bits, n_words, gpos0);
  if (i == 0) {
    base_count[0] = base;
    base_count[1] = rank_of(bitmap, word_prefix, n_bits, n_words, gpos0 + n_pos) - base;
  }
  if (i >= n_pos) return;
  const uint64_t g = gpos0 + i;
  const uint32_t word = __ldg(bitmap + (g >> 5));
  if (!((word >> (g & 31)) & 1u)) return;
  const uint32_t id = __ldg(word_prefix + (g >> 5)) + __popc(word & ((1u << (g & 31)) - 1u));
  unsigned long long key;
  uint32_t f;
  produce<KIND>(items, n_items, S, i, key, f);
  if (KIND == 0) reinterpret_cast<unsigned long long*>(slice)[id - base] = key;
  else reinterpret_cast<uint2*>(slice)[id - base] = make_uint2((uint32_t)(key >> 32), (uint32_t)key);
  pointers[i] = finish_pointer(id, f);
}

// launches finish_first_kernel (shared by both exchanges)
inline int finish_first(Tree& t, int kind, const void* items_dev, uint64_t n_items, uint64_t gpos0, const uint32_t* bitmap_dev,
                        const uint32_t* word_prefix_dev, uint64_t n_level_positions, uint32_t* pointers_dev, void* layer_slice_dev,
                        uint32_t* base_count_dev) {
  const uint64_t n_pos = kind == 0 ? n_items : ceil_div(n_items, 2);
  const uint64_t n_words = ceil_div(n_level_positions, 32);
  const unsigned nb = (unsigned)(n_pos ? ceil_div(n_pos, 256) : 1);
  Launch l(t, "dist_finish_first");
  if (kind == 0)
    finish_first_kernel<0><<<nb, 256, 0, t.stream>>>(items_dev, n_items, n_pos, t.S, gpos0, bitmap_dev, word_prefix_dev, n_level_positions, n_words,
                                                     pointers_dev, layer_slice_dev, base_count_dev);
  else
    finish_first_kernel<1><<<nb, 256, 0, t.stream>>>(items_dev, n_items, n_pos, t.S, gpos0, bitmap_dev, word_prefix_dev, n_level_positions, n_words,
                                                     pointers_dev, layer_slice_dev, base_count_dev);
  return STB_OK;
}

}  // namespace stb
