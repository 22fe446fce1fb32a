// level.cuh — device code shared by the level pipelines of build.cu (one GPU) and shard.cu
// (sharded): tiles of a level, first-occurrence counts, id assignment.
#pragma once

#include "common.cuh"

namespace stb {

constexpr int LVL_THREADS = 256;
constexpr int LVL_ITERS = 4;
constexpr int LVL_TILE = LVL_THREADS * LVL_ITERS;  // positions per CTA in count/assign/resolve

struct LevelTable {
  Slot* slots;        // hash mode: cap + 1 slots
  uint32_t* dminpos;  // direct mode: 4^S entries each
  uint32_t* dids;
  uint32_t cap;
  uint32_t* first_bits = nullptr;  // the level's first-occurrence bitmap, kept current by the inserts
};

__device__ __forceinline__ void load_children(const uint32_t* __restrict__ cur, uint32_t n_cur, uint32_t p, uint32_t& l, uint32_t& r) {
  if (2 * (uint64_t)p + 1 < n_cur) {
    const uint2 pr = __ldg(reinterpret_cast<const uint2*>(cur) + p);
    l = pr.x;
    r = pr.y;
  } else {  // odd tail: node{last, nullptr} (utility.h:17-29)
    l = cur[2 * (uint64_t)p];
    r = PTR_NULL;
  }
}

enum { MODE_LEAF_DIRECT = 0, MODE_LEAF_HASH = 1, MODE_NODE = 2 };

// Loads the tile's 32 bitmap words (one per warp and iteration), leaves the exclusive prefix of
// their popcounts in word_pref[0..31] and the tile's total in word_pref[32].
__device__ __forceinline__ void tile_prefix(const uint32_t* __restrict__ bitmask, uint32_t block, uint32_t n, uint32_t (&words)[LVL_ITERS],
                                            uint32_t* word_pref) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int it = 0; it < LVL_ITERS; ++it) {
    const uint32_t p0 = block * LVL_TILE + it * LVL_THREADS + warp * 32;
    words[it] = p0 < n ? bitmask[p0 >> 5] : 0u;
    if (lane == 0) word_pref[it * (LVL_THREADS / 32) + warp] = __popc(words[it]);
  }
  __syncthreads();
  if (warp == 0) {
    const uint32_t v = word_pref[lane];
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    word_pref[lane] = x - v;
    if (lane == 31) word_pref[32] = x;
  }
  __syncthreads();
}

// First occurrences of one tile: id = rank in position order; append the item to its layer in id
// order and emit the finished pointer.  Node items are recomputed from the child pointers
// (coalesced) rather than fetched from a table (random).
// `id_add` (sharded build): the ids of this rank's slice start there; the slice itself is stored from 0.
template <int MODE>
__device__ __forceinline__ void assign_tile(uint32_t block, uint32_t base, const uint32_t (&words)[LVL_ITERS], const uint32_t* word_pref,
                                            uint32_t* __restrict__ tmp, uint32_t n, const LevelTable& tab, void* __restrict__ uniq, int S,
                                            const uint32_t* __restrict__ children, uint32_t n_children, uint32_t id_add = 0u) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int it = 0; it < LVL_ITERS; ++it) {
    const uint32_t p = block * LVL_TILE + it * LVL_THREADS + threadIdx.x;
    if (p < n && ((words[it] >> lane) & 1u)) {
      const uint32_t rank = base + word_pref[it * (LVL_THREADS / 32) + warp] + __popc(words[it] & ((1u << lane) - 1u));
      const uint32_t t = MODE == MODE_NODE ? 0u : tmp[p];
      const uint32_t s = t & IDX_MASK;
      uint32_t flags = t & ~IDX_MASK;
      if (MODE == MODE_LEAF_DIRECT && (s & LEAF_SIDE)) {
        reinterpret_cast<unsigned long long*>(uniq)[rank] = __ldcg(&tab.slots[s & (LEAF_SIDE - 1u)].key);
        flags &= ~LEAF_SIDE;
      } else if (MODE == MODE_LEAF_DIRECT) {
        tab.dids[s] = rank;
        reinterpret_cast<unsigned long long*>(uniq)[rank] = leaf_from_2bit(s, S);
      } else if (MODE == MODE_LEAF_HASH) {
        reinterpret_cast<unsigned long long*>(uniq)[rank] = (s == tab.cap) ? EMPTY_KEY : __ldcg(&tab.slots[s].key);
      } else {
        uint32_t l, r, cl, cr;
        load_children(children, n_children, p, l, r);
        canonical_node(l, r, cl, cr, flags);
        reinterpret_cast<uint2*>(uniq)[rank] = make_uint2(cl, cr);
      }
      tmp[p] = finish_pointer(id_add + rank, flags);
    }
  }
}

// ids = ranks of the first occurrences in position order.  Two levels of counts instead of a scan:
// count_kernel leaves the number of first occurrences of every tile in blockcnt[] and adds it to
// its chunk's total (CHUNK_TILES tiles per chunk); a tile's base is then the sum of the chunk
// totals before its chunk plus the tile counts before it inside the chunk - at most a thousand
// L2-resident words per CTA at 3.1 Gbp, and no CTA ever waits for another one.  (A single-pass
// scan with decoupled look-back was measured here first: with 1024-position tiles the look-back
// chains through every tile in flight and the kernel was 2-4x slower than this, profiles/README.md.)
constexpr uint32_t CHUNK_TILES = 256;

__device__ __forceinline__ void count_tile(const uint32_t* __restrict__ bitmask, uint32_t tile, uint32_t rel, uint32_t* __restrict__ blockcnt,
                                           uint32_t* __restrict__ chunkcnt) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t c = __popc(bitmask[(uint64_t)tile * (LVL_TILE / 32) + lane]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if (lane == 0) {
    blockcnt[rel] = c;
    if (c) atomicAdd(chunkcnt + rel / CHUNK_TILES, c);
  }
}

// one warp per tile; tiles [first_block, first_block + nb) of the level, counts indexed from 0
static __global__ void __launch_bounds__(256)
count_kernel(const uint32_t* __restrict__ bitmask, uint32_t first_block, uint32_t nb, uint32_t* __restrict__ blockcnt,
             uint32_t* __restrict__ chunkcnt) {
  const uint32_t rel = (blockIdx.x * 256 + threadIdx.x) >> 5;
  if (rel < nb) count_tile(bitmask, first_block + rel, rel, blockcnt, chunkcnt);
}

// Sum over the CTA of: the chunk totals before tile `rel`'s chunk + the tile counts before it in
// its chunk.  `red` is shared scratch of LVL_THREADS / 32 words; the caller synchronises before
// reading the result of the next call into the same scratch.
__device__ __forceinline__ uint32_t tile_base(uint32_t rel, const uint32_t* __restrict__ blockcnt, const uint32_t* __restrict__ chunkcnt,
                                              uint32_t* red) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t chunk = rel / CHUNK_TILES, in_chunk = chunk * CHUNK_TILES + threadIdx.x;
  uint32_t part = in_chunk < rel ? blockcnt[in_chunk] : 0u;
  for (uint32_t i = threadIdx.x; i < chunk; i += LVL_THREADS) part += chunkcnt[i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  uint32_t base = 0;
#pragma unroll
  for (int w = 0; w < LVL_THREADS / 32; ++w) base += red[w];
  return base;
}

// `carry_in` (streaming build): the level's running total before this launch's tiles; the last
// tile leaves the new total in *total_out (a different word: other CTAs still read carry_in).
template <int MODE>
__global__ void __launch_bounds__(LVL_THREADS)
assign_kernel(uint32_t* __restrict__ tmp, uint32_t n, LevelTable tab, const uint32_t* __restrict__ bitmask,
              const uint32_t* __restrict__ blockcnt, const uint32_t* __restrict__ chunkcnt, uint32_t first_block,
              const uint32_t* __restrict__ carry_in, uint32_t* __restrict__ total_out, void* __restrict__ uniq, int S,
              const uint32_t* __restrict__ children, uint32_t n_children, const uint32_t* __restrict__ id_offset = nullptr) {
  static_assert(CHUNK_TILES == LVL_THREADS, "tile_base reads one tile count per thread");
  __shared__ uint32_t word_pref[33];
  __shared__ uint32_t red[LVL_THREADS / 32];
  const uint32_t rel = blockIdx.x, block = first_block + rel;
  // a tile without first occurrences has nothing to assign (most tiles of a genome's leaf level: the 4^12
  // possible leaves have nearly all occurred after the first few per cent of the text)
  if (rel != gridDim.x - 1 && blockcnt[rel] == 0u) return;
  const uint32_t base = tile_base(rel, blockcnt, chunkcnt, red) + (carry_in ? *carry_in : 0u);
  uint32_t words[LVL_ITERS];
  tile_prefix(bitmask, block, n, words, word_pref);
  if (rel == gridDim.x - 1 && threadIdx.x == 0) *total_out = base + word_pref[32];
  assign_tile<MODE>(block, base, words, word_pref, tmp, n, tab, uniq, S, children, n_children, id_offset ? *id_offset : 0u);
}

}  // namespace stb
