// build.cu — level-by-level bottom-up construction of the shared tree (hash-consing).
//
// Replaces tree_constructor (reference include/shared_tree.h:245-316,
// src/shared_tree.cpp:621-763) and the phmap dedup behind it.  The reference
// streams 2^22-leaf segments through per-layer hash maps, one element at a time;
// because its segments are powers of two that is equivalent to ONE global pass per
// level, which is what runs here:
//
//   per level (leaves, then node layers bottom-up):
//     insert   every position canonicalises its item and lowers the item's
//              min-position in an open-addressing table (or a direct-addressed
//              table for ACGT-only leaves); whoever becomes a key's minimum
//              toggles its own bit in the level's bitmap and the bit of the
//              position it displaced (XOR commutes), so when the kernel ends the
//              bitmap marks exactly the first occurrences      [random HBM/L2]
//     scan     per-CTA popcounts of the bitmap, exclusive scan [tiny]
//     assign   first occurrences get id = rank, append the item to the layer in id
//              order and emit their pointer                    [coalesced]
//     resolve  later occurrences: slot -> min-position -> the first occurrence's
//              finished pointer -> id                          [random reads]
//
// IDs are first-occurrence ranks in position order, independent of which thread won
// which atomic, so the result equals the reference's sequential emplace order.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "pack.cuh"
#include "tree.h"

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

struct BuildFlags {
  unsigned long long bad_symbol;  // (byte offset << 8) | upper-cased byte; ~0 = none
  uint32_t non_acgt;              // direct leaf mode met a non-ACGT leaf
  uint32_t bad_leaf;              // packed input leaf with bits >= 4S
};

template <bool DIRECT>
__device__ __forceinline__ void insert_leaf(unsigned long long v, int S, uint32_t pos, const LevelTable& tab,
                                            uint32_t* __restrict__ tmp_at_pos, BuildFlags* flags) {
  uint32_t f;
  const unsigned long long canon = canonical_leaf(v, S, f);
  uint32_t s;
  if (DIRECT) {
    if (!leaf_is_acgt(v, S)) {
      flags->non_acgt = 1u;
      *tmp_at_pos = 0u;
      return;
    }
    s = leaf_to_2bit(canon);
    if (__ldcg(tab.dminpos + s) > pos) {
      const uint32_t old = atomicMin(tab.dminpos + s, pos);
      if (tab.first_bits && old > pos) {  // see toggle_bit (common.cuh)
        toggle_bit(tab.first_bits, pos);
        if (old != 0xffffffffu) toggle_bit(tab.first_bits, old);
      }
    }
  } else {
    s = table_insert<true>(tab.slots, tab.cap, canon, pos, tab.first_bits);
  }
  *tmp_at_pos = s | f;
}

// Leaves straight from the ASCII body: pack (dna.cpp:79-84) + canonical (dna.cpp:135)
// + emplace_leaf (shared_tree.cpp:630) fused; the packed leaves never touch HBM.
template <int S_T, bool DIRECT>
__global__ void __launch_bounds__(PACK_THREADS)
leaf_insert_text_kernel(const char* __restrict__ body, uint64_t n_leaves, int S_rt, LevelTable tab,
                        uint32_t* __restrict__ tmp, BuildFlags* flags, uint32_t pos0) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* lut = smem;
  uint8_t* tile = smem + 256;
  const int S = S_T > 0 ? S_T : S_rt;
  const uint64_t tile_first = (uint64_t)blockIdx.x * PACK_TILE_LEAVES;
  const uint32_t here = (uint32_t)min((uint64_t)PACK_TILE_LEAVES, n_leaves - tile_first);
  stage_text_tile(lut, tile, body + tile_first * S, here * (uint32_t)S);
#pragma unroll
  for (int it = 0; it < PACK_LEAVES_PER_THREAD; ++it) {
    const uint32_t j = it * PACK_THREADS + threadIdx.x;
    if (j < here) {
      uint32_t bad = 0xFFFFFFFFu;
      const unsigned long long v = pack_leaf<S_T>(lut, tile, j, S, bad);
      if (bad != 0xFFFFFFFFu) {
        uint32_t c = tile[bad];
        if (c >= 'a' && c <= 'z') c -= 32;
        atomicMin(&flags->bad_symbol, ((tile_first * S + bad) << 8) | c);
      }
      insert_leaf<DIRECT>(v, S, pos0 + (uint32_t)(tile_first + j), tab, tmp + tile_first + j, flags);
    }
  }
}

// ACGT fast path for the default leaf length (12): the whole leaf level in 32-bit registers.
// Four text bytes -> four 2-bit codes by SWAR (A0 C1 G2 T3; any other byte, in either case,
// fails the check), twelve of them = one 24-bit word whose numeric order equals the order of
// the reference's 4-bit words (A1<C2<G4<T8, same nucleotide significance), so the canonical
// variant (dna.cpp:135-143) is the minimum of {c, ~c, reverse(c), reverse(~c)} taken on that
// word, and the word itself is the direct-table index.  Anything that is not pure ACGT raises
// the non_acgt flag and the host re-runs the level through the general kernel.
__device__ __forceinline__ bool acgt_word_to_2bit(uint32_t w, uint32_t& out8) {
  const uint32_t u = w & 0xDFDFDFDFu;                       // fold case
  const uint32_t x = (u >> 1) & 0x03030303u;                // A0 C1 G3 T2
  const uint32_t idx = x ^ ((x >> 1) & 0x01010101u);        // A0 C1 G2 T3
  const uint32_t b0 = idx & 0x01010101u, b1 = (idx >> 1) & 0x01010101u, bb = b0 & b1;
  // the byte each code must have come from: 'A' + {0, 2, 6, 0x13}
  const uint32_t expect = 0x41414141u + (b0 << 1) + (b1 << 1) + (b1 << 2) + bb + (bb << 1) + (bb << 3);
  out8 = (idx * 0x01041040u) >> 24;                          // 4 codes -> 8 bits, first char lowest
  return expect == u;
}

__device__ __forceinline__ uint32_t reverse_2bit_24(uint32_t c) {
  const uint32_t r = __brev(c) >> 8;  // reverses the pairs and the bits inside each pair
  return ((r & 0x555555u) << 1) | ((r >> 1) & 0x555555u);
}

__global__ void __launch_bounds__(PACK_THREADS)
leaf_insert_acgt12_kernel(const char* __restrict__ body, uint64_t n_leaves, LevelTable tab, uint32_t* __restrict__ tmp,
                          BuildFlags* flags, uint32_t pos0) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* tile = smem + 256;  // same layout as the general kernel (the table area is unused here)
  const uint64_t tile_first = (uint64_t)blockIdx.x * PACK_TILE_LEAVES;
  const uint32_t here = (uint32_t)min((uint64_t)PACK_TILE_LEAVES, n_leaves - tile_first);
  stage_text_tile(smem, tile, body + tile_first * 12, here * 12u);
  const uint32_t* words = reinterpret_cast<const uint32_t*>(tile);
#pragma unroll
  for (int it = 0; it < PACK_LEAVES_PER_THREAD; ++it) {
    const uint32_t j = it * PACK_THREADS + threadIdx.x;
    if (j >= here) continue;
    uint32_t c0, c1, c2;
    const bool ok = acgt_word_to_2bit(words[3 * j], c0) & acgt_word_to_2bit(words[3 * j + 1], c1) & acgt_word_to_2bit(words[3 * j + 2], c2);
    if (!ok) {
      flags->non_acgt = 1u;
      tmp[tile_first + j] = 0u;
      continue;
    }
    const uint32_t c = c0 | (c1 << 8) | (c2 << 16);
    const uint32_t t = ~c & 0xFFFFFFu, m = reverse_2bit_24(c), i = ~m & 0xFFFFFFu;
    uint32_t best = c, f = 0;
    if (t < best) { best = t; f = TRANSPOSE; }
    if (m < best) { best = m; f = MIRROR; }
    if (i < best) { best = i; f = MIRROR | TRANSPOSE; }
    if (c == m) f |= INVARIANT;
    const uint32_t pos = pos0 + (uint32_t)(tile_first + j);
    if (__ldcg(tab.dminpos + best) > pos) {
      const uint32_t old = atomicMin(tab.dminpos + best, pos);
      if (tab.first_bits && old > pos) {
        toggle_bit(tab.first_bits, pos);
        if (old != 0xffffffffu) toggle_bit(tab.first_bits, old);
      }
    }
    tmp[tile_first + j] = best | f;
  }
}

// Leaves from a packed array (the std::vector<dna> constructor, shared_tree.cpp:212).
template <bool DIRECT>
__global__ void __launch_bounds__(LVL_THREADS)
leaf_insert_u64_kernel(const unsigned long long* __restrict__ leaves, uint32_t n, int S, LevelTable tab,
                       uint32_t* __restrict__ tmp, BuildFlags* flags) {
  const uint32_t p = blockIdx.x * LVL_THREADS + threadIdx.x;
  if (p >= n) return;
  const unsigned long long v = __ldg(leaves + p);
  if (v & ~leaf_mask(S)) flags->bad_leaf = 1u;
  insert_leaf<DIRECT>(v, S, p, tab, tmp + p, flags);
}

// Singleton filter for the first node layer (its keys are pairs of leaf ids: no locality to
// exploit, so every table access is a random HBM line).  Two bit planes that fit in L2: plane A
// = "some position hashed here", plane B = "at least two did".  A position whose B bit stays
// clear shares its filter cell with nobody, so its key occurs exactly once in the level: it is
// a first occurrence that nobody will ever look up, and it skips the table altogether.
struct SingletonFilter {
  uint32_t* plane_a = nullptr;
  uint32_t* plane_b = nullptr;
  uint32_t log2_bits = 0;
};

__device__ __forceinline__ void filter_cell(const SingletonFilter& flt, unsigned long long key, uint32_t& word, uint32_t& bit) {
  const uint32_t h = (uint32_t)mix64(key) >> (32 - flt.log2_bits);  // low half of the mix: the table uses the high half
  word = h >> 5;
  bit = 1u << (h & 31);
}

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

__global__ void __launch_bounds__(LVL_THREADS)
node_filter_kernel(const uint32_t* __restrict__ cur, uint32_t n_cur, uint32_t n_next, SingletonFilter flt) {
#pragma unroll
  for (int it = 0; it < LVL_ITERS; ++it) {
    const uint32_t p = blockIdx.x * LVL_TILE + it * LVL_THREADS + threadIdx.x;
    if (p >= n_next) return;
    uint32_t l, r, cl, cr, f, word, bit;
    load_children(cur, n_cur, p, l, r);
    canonical_node(l, r, cl, cr, f);
    filter_cell(flt, ((unsigned long long)cl << 32) | cr, word, bit);
    if (atomicOr(flt.plane_a + word, bit) & bit) atomicOr(flt.plane_b + word, bit);
  }
}

// One node level: reduce_nodes + emplace_node (shared_tree.cpp:697-712, :662-672).
__global__ void __launch_bounds__(LVL_THREADS)
node_insert_kernel(const uint32_t* __restrict__ cur, uint32_t n_cur, uint32_t n_next, LevelTable tab, uint32_t* __restrict__ tmp,
                   const uint32_t* __restrict__ child_unique, uint32_t serial, SingletonFilter flt, uint32_t first_block) {
  // several positions per thread: CTAs that live for one probe each are dispatch-bound
  for (int it = 0; it < LVL_ITERS; ++it) {
  const uint32_t p = (first_block + blockIdx.x) * LVL_TILE + it * LVL_THREADS + threadIdx.x;
  if (p >= n_next) return;
  uint32_t l, r;
  load_children(cur, n_cur, p, l, r);
  uint32_t cl, cr, f;
  canonical_node(l, r, cl, cr, f);
  const unsigned long long key = ((unsigned long long)cl << 32) | cr;
  if (flt.plane_b) {
    uint32_t word, bit;
    filter_cell(flt, key, word, bit);
    if (!(__ldcg(flt.plane_b + word) & bit)) {  // the only position with this key
      atomicOr(tab.first_bits + (p >> 5), 1u << (p & 31));
      continue;  // assign_kernel recomputes the node; nobody resolves through tmp[p]
    }
  }
  const uint32_t hashed = __umulhi(hash64(key), tab.cap);
  uint32_t start = hashed, limit = 0xffffffffu;
  if (child_unique) {
    // Locality placement (node layers above the first): child ids are first-occurrence ranks, so
    // they grow with the position; a slot proportional to a child id makes neighbouring positions
    // probe neighbouring slots (one 128-byte line serves several positions instead of one line
    // per position).  Crowded neighbourhoods (one child with many partners) fall back to the hash.
    const uint32_t child = ptr_is_null(cl) ? (cr & IDX_MASK) : (cl & IDX_MASK);
    // (any deterministic function of the key will do: single-precision scaling, no 64-bit divide)
    const float ratio = __fdividef((float)tab.cap, (float)max(1u, __ldg(child_unique)));
    start = (uint32_t)min((float)(tab.cap - 9u), (float)child * ratio) + (hashed & 7u);
    limit = 24u;
  }
  tmp[p] = tagged_insert(tab.slots, tab.cap, key, p, serial, start, hashed, limit, tab.first_bits) | f;
  }
}

// per-CTA first-occurrence counts (LVL_TILE positions = 32 bitmask words) for the scan
__global__ void __launch_bounds__(256)
bitmask_blockcnt_kernel(const uint32_t* __restrict__ bitmask, uint32_t n_blocks, uint32_t* __restrict__ blockcnt) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t blk = (blockIdx.x * 256 + threadIdx.x) >> 5;  // one warp per block of 1024 positions
  if (blk >= n_blocks) return;
  uint32_t c = __popc(bitmask[blk * (LVL_TILE / 32) + lane]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if (lane == 0) blockcnt[blk] = c;
}

// In-place exclusive scan of the per-CTA counts (single CTA, 16 entries per thread and
// round); total -> *total_out.
constexpr int SCAN_PER_THREAD = 16;
__global__ void __launch_bounds__(1024) scan_blocks_kernel(uint32_t* __restrict__ cnt, uint32_t nb,
                                                           uint32_t* total_out, const uint32_t* carry_in) {
  __shared__ uint32_t warp_sum[32];
  __shared__ uint32_t carry_s;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = carry_in ? *carry_in : 0u;  // streaming build: ids continue where the last chunk stopped
  __syncthreads();
  for (uint32_t base = 0; base < nb; base += 1024 * SCAN_PER_THREAD) {
    const uint32_t i0 = base + threadIdx.x * SCAN_PER_THREAD;
    uint32_t v[SCAN_PER_THREAD];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; ++j) {
      v[j] = i0 + j < nb ? cnt[i0 + j] : 0u;
      sum += v[j];
    }
    uint32_t x = sum;
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
      warp_sum[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    uint32_t run = carry_s + (warp ? warp_sum[warp - 1] : 0u) + x - sum;
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; ++j) {
      if (i0 + j < nb) cnt[i0 + j] = run;
      run += v[j];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry_s;
}

// Large levels: one CTA per chunk of 1024 * SCAN_PER_THREAD counts scans its chunk in place
// (exclusive, from 0) and leaves the chunk total in chunk_sum[]; scan_blocks_kernel then scans
// the few chunk totals, and assign_kernel adds its chunk's base.  The single-CTA loop over a
// 3.1 Gbp leaf level is 16 dependent rounds; this is one.
constexpr uint32_t SCAN_CHUNK = 1024 * SCAN_PER_THREAD;
constexpr uint32_t SCAN_MAX_CHUNKS = 1024;
__global__ void __launch_bounds__(1024) scan_chunks_kernel(uint32_t* __restrict__ cnt, uint32_t nb, uint32_t* __restrict__ chunk_sum) {
  __shared__ uint32_t warp_sum[32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t i0 = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_PER_THREAD;
  uint32_t v[SCAN_PER_THREAD];
  uint32_t sum = 0;
#pragma unroll
  for (int j = 0; j < SCAN_PER_THREAD; ++j) {
    v[j] = i0 + j < nb ? cnt[i0 + j] : 0u;
    sum += v[j];
  }
  uint32_t x = sum;
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
    warp_sum[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  uint32_t run = (warp ? warp_sum[warp - 1] : 0u) + x - sum;
#pragma unroll
  for (int j = 0; j < SCAN_PER_THREAD; ++j) {
    if (i0 + j < nb) cnt[i0 + j] = run;
    run += v[j];
  }
  if (threadIdx.x == 1023) chunk_sum[blockIdx.x] = run;
}

enum { MODE_LEAF_DIRECT = 0, MODE_LEAF_HASH = 1, MODE_NODE = 2 };

// First occurrences: id = rank in position order; append the item to its layer in id order
// and emit the finished pointer.  Node items are recomputed from the child pointers
// (coalesced) rather than fetched from the table (random).
template <int MODE>
__global__ void __launch_bounds__(LVL_THREADS)
assign_kernel(uint32_t* __restrict__ tmp, uint32_t n, LevelTable tab, const uint32_t* __restrict__ bitmask,
              const uint32_t* __restrict__ blockbase, void* __restrict__ uniq, int S,
              const uint32_t* __restrict__ children, uint32_t n_children, uint32_t first_block,
              const uint32_t* __restrict__ chunk_base = nullptr) {
  __shared__ uint32_t word_pref[32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t block = first_block + blockIdx.x;
  uint32_t words[LVL_ITERS];
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
  }
  __syncthreads();
  const uint32_t base = blockbase[block] + (chunk_base ? chunk_base[block / SCAN_CHUNK] : 0u);
#pragma unroll
  for (int it = 0; it < LVL_ITERS; ++it) {
    const uint32_t p = block * LVL_TILE + it * LVL_THREADS + threadIdx.x;
    if (p < n && ((words[it] >> lane) & 1u)) {
      const uint32_t rank = base + word_pref[it * (LVL_THREADS / 32) + warp] + __popc(words[it] & ((1u << lane) - 1u));
      const uint32_t t = MODE == MODE_NODE ? 0u : tmp[p];
      const uint32_t s = t & IDX_MASK;
      uint32_t flags = t & ~IDX_MASK;
      if (MODE == MODE_LEAF_DIRECT) {
        tab.dids[s] = rank;
        reinterpret_cast<unsigned long long*>(uniq)[rank] = leaf_from_2bit(s, S);
      } else if (MODE == MODE_LEAF_HASH) {
        reinterpret_cast<unsigned long long*>(uniq)[rank] = (s == tab.cap) ? EMPTY_KEY : __ldcg(&tab.slots[s].key);
      } else {
        uint32_t l, r;
        if (2 * (uint64_t)p + 1 < n_children) {
          const uint2 pr = __ldg(reinterpret_cast<const uint2*>(children) + p);
          l = pr.x;
          r = pr.y;
        } else {
          l = children[2 * (uint64_t)p];
          r = PTR_NULL;
        }
        uint32_t cl, cr;
        canonical_node(l, r, cl, cr, flags);
        reinterpret_cast<uint2*>(uniq)[rank] = make_uint2(cl, cr);
      }
      tmp[p] = finish_pointer(rank, flags);
    }
  }
}

// Later occurrences: direct mode looks the id up in the (L2-sized) id table; hash levels
// read it from the finished pointer of the first occurrence, whose position count_first
// left in tmp[p].
template <bool DIRECT>
__global__ void __launch_bounds__(LVL_THREADS)
resolve_kernel(uint32_t* __restrict__ tmp, uint32_t n, LevelTable tab, const uint32_t* __restrict__ bitmask, uint32_t first_block) {
  const uint32_t lane = threadIdx.x & 31;
#pragma unroll
  for (int it = 0; it < LVL_ITERS; ++it) {
    const uint32_t p = (first_block + blockIdx.x) * LVL_TILE + it * LVL_THREADS + threadIdx.x;
    if (p < n) {
      const uint32_t word = bitmask[p >> 5];
      if (!((word >> lane) & 1u)) {
        const uint32_t t = tmp[p];
        const uint32_t s = t & IDX_MASK;
        // hash levels: slot -> position of the first occurrence -> its finished pointer -> id
        const uint32_t id = DIRECT ? __ldcg(tab.dids + s) : (__ldcg(tmp + __ldcg(&tab.slots[s].minpos)) & IDX_MASK);
        tmp[p] = finish_pointer(id, t & ~IDX_MASK);
      }
    }
  }
}

// ---- host orchestration -----------------------------------------------------------

namespace {

struct LeafInput {
  const char* body = nullptr;                 // device, 16-byte aligned
  const unsigned long long* leaves = nullptr; // device
};

struct Scratch {
  DevBuf<uint32_t> ptr_a, ptr_b, bitmask, blockcnt, chunk_sum, counts, dminpos, dids;
  DevBuf<Slot> slots;
  DevBuf<BuildFlags> flags;
  DevBuf<uint32_t> root;
  // Node tables: a slot's last word is an epoch tag; every node level of every build on this
  // handle takes a fresh serial, so the table is cleared only when its memory is new.
  bool tags_cleared = false;
  uint32_t serial = 0;
  DevBuf<uint32_t> filter;    // singleton filter of the first node layer (two bit planes)

  // streaming build (host input): every chunked level keeps its own pointer array, bitmap,
  // per-CTA counts and table for the whole build
  DevBuf<uint32_t> ptr_arena, bit_arena, cnt_arena, level_sizes;
  DevBuf<Slot> stream_slots;
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> chunk_events;
  ~Scratch() {
    for (auto e : chunk_events) cudaEventDestroy(e);
    if (copy_stream) cudaStreamDestroy(copy_stream);
  }
};

Scratch& workspace_of(Tree& t) {
  if (!t.workspace) t.workspace = std::make_shared<Scratch>();
  Scratch& sc = *static_cast<Scratch*>(t.workspace.get());
  if (sc.serial > 0xfff00000u) {  // far from wrapping into a tag that is still in the table
    sc.serial = 0;
    sc.tags_cleared = false;
  }
  return sc;
}

uint32_t table_cap(uint64_t n) { return (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1024, 2 * n), 0x1ffffffeull); }

template <int S_T, bool DIRECT>
void launch_leaf_text(Ctx& ctx, const char* body, uint64_t n, LevelTable tab, uint32_t* tmp, BuildFlags* flags, uint32_t pos0 = 0) {
  const size_t smem = pack_smem_bytes(ctx.S);
  Launch l(ctx, "leaf_insert");
  if (S_T == 12 && DIRECT) {
    leaf_insert_acgt12_kernel<<<(unsigned)ceil_div(n, PACK_TILE_LEAVES), PACK_THREADS, smem, ctx.stream>>>(body, n, tab, tmp, flags, pos0);
    return;
  }
  leaf_insert_text_kernel<S_T, DIRECT><<<(unsigned)ceil_div(n, PACK_TILE_LEAVES), PACK_THREADS, smem, ctx.stream>>>(
      body, n, ctx.S, tab, tmp, flags, pos0);
}

// count -> scan -> assign -> resolve for one level whose inserts are already queued.
template <int MODE>
void finish_level(Ctx& ctx, uint32_t* tmp, uint32_t n, LevelTable tab, Scratch& sc, uint32_t* total_out, void* uniq,
                  const uint32_t* children = nullptr, uint32_t n_children = 0) {
  constexpr bool DIRECT = (MODE == MODE_LEAF_DIRECT);
  const unsigned nb = (unsigned)ceil_div(n, LVL_TILE);
  {  // the inserts kept the first-occurrence bitmap current: count it per CTA tile for the scan
    Launch l(ctx, "bitmask_blockcnt");
    bitmask_blockcnt_kernel<<<(unsigned)ceil_div((uint64_t)nb * 32, 256), 256, 0, ctx.stream>>>(sc.bitmask.ptr, nb, sc.blockcnt.ptr);
  }
  const uint32_t* chunk_base = nullptr;
  if (nb > SCAN_CHUNK && sc.chunk_sum.ptr) {  // two levels: chunks in parallel, then their (few) totals
    const unsigned nchunks = (unsigned)ceil_div(nb, SCAN_CHUNK);
    {
      Launch l(ctx, "scan_blocks");
      scan_chunks_kernel<<<nchunks, 1024, 0, ctx.stream>>>(sc.blockcnt.ptr, nb, sc.chunk_sum.ptr);
    }
    Launch l(ctx, "scan_blocks");
    scan_blocks_kernel<<<1, 1024, 0, ctx.stream>>>(sc.chunk_sum.ptr, nchunks, total_out, nullptr);
    chunk_base = sc.chunk_sum.ptr;
  } else {
    Launch l(ctx, "scan_blocks");
    scan_blocks_kernel<<<1, 1024, 0, ctx.stream>>>(sc.blockcnt.ptr, nb, total_out, nullptr);
  }
  {
    Launch l(ctx, "assign_ids");
    assign_kernel<MODE><<<nb, LVL_THREADS, 0, ctx.stream>>>(tmp, n, tab, sc.bitmask.ptr, sc.blockcnt.ptr, uniq, ctx.S, children, n_children, 0u,
                                                            chunk_base);
  }
  {
    Launch l(ctx, "resolve_ids");
    resolve_kernel<DIRECT><<<nb, LVL_THREADS, 0, ctx.stream>>>(tmp, n, tab, sc.bitmask.ptr, 0u);
  }
}

// ---- the small top of the tree in ONE launch ----------------------------------------------
// Once a level has at most SMALL_MAX pointers, every remaining level (about a dozen) runs inside
// a single CTA: pointers ping-pong in shared memory, the hash table lives in shared memory, and
// __syncthreads() replaces the ~5 launches per level of the general path.
constexpr uint32_t SMALL_MAX = 2048;        // pointers entering the kernel (<= 1024 positions per level)
constexpr uint32_t SMALL_SLOTS = 2048;      // shared-memory table slots (load <= 0.5)
constexpr int SMALL_MAX_LEVELS = 16;

struct SmallOut {
  uint2* nodes[SMALL_MAX_LEVELS];  // layer buffers of the levels this launch will build
};

__global__ void __launch_bounds__(1024)
small_levels_kernel(const uint32_t* __restrict__ cur_in, uint32_t n_cur, SmallOut out, uint32_t* __restrict__ counts,
                    uint32_t* __restrict__ root_out) {
  __shared__ uint32_t ptr_a[SMALL_MAX], ptr_b[SMALL_MAX / 2];
  __shared__ unsigned long long skey[SMALL_SLOTS];
  __shared__ uint32_t smin[SMALL_SLOTS];
  __shared__ uint32_t warp_cnt[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (uint32_t i = tid; i < n_cur; i += 1024) ptr_a[i] = cur_in[i];
  uint32_t* cur = ptr_a;
  uint32_t* nxt = ptr_b;
  int level = 0;
  do {
    const uint32_t n_next = (n_cur + 1) / 2;
    for (uint32_t i = tid; i < SMALL_SLOTS; i += 1024) {
      skey[i] = EMPTY_KEY;
      smin[i] = 0xffffffffu;
    }
    __syncthreads();
    uint32_t cl = 0, cr = 0, f = 0, slot = 0;
    const bool active = tid < n_next;
    if (active) {
      const uint32_t l = cur[2 * tid];
      const uint32_t r = 2 * tid + 1 < n_cur ? cur[2 * tid + 1] : PTR_NULL;
      canonical_node(l, r, cl, cr, f);
      const unsigned long long key = ((unsigned long long)cl << 32) | cr;
      slot = hash64(key) & (SMALL_SLOTS - 1);
      for (;;) {
        const unsigned long long old = atomicCAS(&skey[slot], EMPTY_KEY, key);
        if (old == EMPTY_KEY || old == key) break;
        slot = (slot + 1) & (SMALL_SLOTS - 1);
      }
      atomicMin(&smin[slot], tid);
    }
    __syncthreads();
    const bool first = active && smin[slot] == tid;
    const uint32_t word = __ballot_sync(0xffffffffu, first);
    if (lane == 0) warp_cnt[warp] = __popc(word);
    __syncthreads();
    if (warp == 0) {
      const uint32_t v = warp_cnt[lane];
      uint32_t x = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
      }
      warp_cnt[lane] = x - v;
      if (lane == 31) counts[level] = x;
    }
    __syncthreads();
    if (first) {
      const uint32_t id = warp_cnt[warp] + __popc(word & ((1u << lane) - 1u));
      out.nodes[level][id] = make_uint2(cl, cr);
      nxt[tid] = finish_pointer(id, f);
    }
    __syncthreads();
    if (active && !first) nxt[tid] = finish_pointer(nxt[smin[slot]] & IDX_MASK, f);
    __syncthreads();
    // the next level reads what this one wrote
    if (cur == ptr_a) { cur = ptr_b; nxt = ptr_a; } else { cur = ptr_a; nxt = ptr_b; }
    n_cur = n_next;
    ++level;
  } while (n_cur > 1);
  if (tid == 0) *root_out = cur[0];
}

// Tunables of the partitioned path (environment overrides are for experiments only).
static uint64_t env_u64(const char* name, uint64_t fallback) {
  const char* v = getenv(name);
  return v ? strtoull(v, nullptr, 0) : fallback;
}

// words of the two filter planes for a first node layer of n positions (0 = filter not used)
static uint64_t filter_words(uint64_t n) {
  if (n < env_u64("STB_FILTER_MIN", 1ull << 22)) return 0;
  uint32_t log2_bits = 22;
  while (log2_bits < 28 && (1ull << log2_bits) < 2 * n) ++log2_bits;
  return 2 * ((1ull << log2_bits) / 32);
}

// Node levels from a pointer array down to a single root pointer.  Appends one layer per
// level to t.layers; counts_dev[level] receives each layer's size; returns the buffer that
// holds the root pointer in *root_buf.
int run_node_levels(Tree& t, Scratch& sc, uint32_t* cur, uint32_t* nxt, uint64_t n_cur, uint32_t* counts_dev, int* levels_out,
                    uint32_t** root_buf) {
  static const bool locality = env_u64("STB_LOCALITY", 1) != 0;
  cudaStream_t st = t.stream;
  int level = 0;
  do {
    if (n_cur <= SMALL_MAX) {  // the rest of the tree in one launch
      SmallOut out{};
      int extra = 0;
      for (uint64_t n = n_cur;;) {
        n = ceil_div(n, 2);
        t.layers.emplace_back();
        STB_CUDA(t, t.layers.back().nodes.alloc(n, st));
        out.nodes[extra++] = t.layers.back().nodes.ptr;
        if (n <= 1) break;
      }
      STB_CUDA(t, sc.root.ensure(1, st));
      Launch l(t, "small_levels");
      small_levels_kernel<<<1, 1024, 0, st>>>(cur, (uint32_t)n_cur, out, counts_dev + level, sc.root.ptr);
      level += extra;
      cur = sc.root.ptr;
      break;
    }
    const uint64_t n_next = ceil_div(n_cur, 2);
    t.layers.emplace_back();
    Layer& layer = t.layers.back();
    STB_CUDA(t, layer.nodes.alloc(n_next, st));
    LevelTable nt{sc.slots.ptr, nullptr, nullptr, table_cap(n_next)};
    nt.first_bits = sc.bitmask.ptr;
    if (!sc.tags_cleared) {
      // Node tables are never cleared between levels: a slot's last word is an epoch tag and a
      // slot whose tag is not this level's serial counts as empty.  One clear per build.
      Launch l(t, "table_clear", false);
      STB_CUDA(t, cudaMemsetAsync(sc.slots.ptr, 0xff, sc.slots.bytes(), st));
      sc.tags_cleared = true;
      sc.serial = 0;
    }
    STB_CUDA(t, cudaMemsetAsync(sc.bitmask.ptr, 0, ceil_div(n_next, LVL_TILE) * (LVL_TILE / 8), st));
    SingletonFilter flt;
    static const uint64_t filter_min = env_u64("STB_FILTER_MIN", 1ull << 22);
    if (level == 0 && n_next >= filter_min) {
      // two planes of 2^k bits, k as large as keeps both in L2 (<= 2 x 32 MiB)
      static const uint32_t max_log2 = (uint32_t)env_u64("STB_FILTER_LOG2", 28);
      flt.log2_bits = 22;
      while (flt.log2_bits < max_log2 && (1ull << flt.log2_bits) < 2 * n_next) ++flt.log2_bits;
      const uint64_t words = (1ull << flt.log2_bits) / 32;
      STB_CUDA(t, sc.filter.ensure(2 * words, st));  // already there: sized with the rest of the workspace
      STB_CUDA(t, cudaMemsetAsync(sc.filter.ptr, 0, 2 * words * 4, st));
      flt.plane_a = sc.filter.ptr;
      flt.plane_b = sc.filter.ptr + words;
      Launch l(t, "node_filter");
      node_filter_kernel<<<(unsigned)ceil_div(n_next, LVL_TILE), LVL_THREADS, 0, st>>>(cur, (uint32_t)n_cur, (uint32_t)n_next, flt);
    }
    {
      // Probe-then-claim measured faster than claim-first on B200 (12.0 vs 13.2 ms per 3.1 Gbp),
      // and chunking the level to keep table lines in L2 did not pay (profiles/README.md).
      Launch l(t, "node_insert");
      // children of level 0 are leaf ids (or an imported array): not position-ordered
      const uint32_t* child_unique = (locality && level > 0) ? counts_dev + level - 1 : nullptr;
      node_insert_kernel<<<(unsigned)ceil_div(n_next, LVL_TILE), LVL_THREADS, 0, st>>>(cur, (uint32_t)n_cur, (uint32_t)n_next, nt, nxt, child_unique,
                                                                                          ++sc.serial, flt, 0u);
    }
    finish_level<MODE_NODE>(t, nxt, (uint32_t)n_next, nt, sc, counts_dev + level, layer.nodes.ptr, cur, (uint32_t)n_cur);
    std::swap(cur, nxt);
    n_cur = n_next;
    ++level;
  } while (n_cur > 1);
  *levels_out = level;
  *root_buf = cur;
  return STB_OK;
}

// Reads back the per-layer counts, the flags and the root; fills in the tree; trims storage.
int finish_build(Tree& t, Scratch& sc, uint64_t n0, int level, const uint32_t* cur, bool direct) {
  cudaStream_t st = t.stream;
  std::vector<uint32_t> counts(level + 1);
  BuildFlags flags{};
  uint32_t root = PTR_NULL;
  STB_CUDA(t, cudaMemcpyAsync(counts.data(), sc.counts.ptr, counts.size() * 4, cudaMemcpyDeviceToHost, st));
  STB_CUDA(t, cudaMemcpyAsync(&flags, sc.flags.ptr, sizeof(flags), cudaMemcpyDeviceToHost, st));
  STB_CUDA(t, cudaMemcpyAsync(&root, cur, 4, cudaMemcpyDeviceToHost, st));
  STB_CUDA(t, cudaStreamSynchronize(st));
  STB_CUDA(t, cudaGetLastError());

  if (flags.bad_symbol != ~0ull) {
    t.clear();
    return t.fail(STB_ERR_UNKNOWN_SYMBOL, unknown_symbol_message((int)(flags.bad_symbol & 0xff)));
  }
  if (flags.bad_leaf) {
    t.clear();
    return t.fail(STB_ERR_BAD_LEAF, "packed leaf has bits set at or above 4*dna_size");
  }
  if (direct && flags.non_acgt) {
    t.clear();
    return -1;  // caller retries with the hash-table leaf level
  }
  for (uint32_t c : counts)
    if (c >= IDX_MASK) {
      t.clear();
      return t.fail(STB_ERR_INDEX_CEILING, "a layer has 2^29-1 or more unique items; the pointer format cannot index it");
    }

  t.n_leaves = counts[0];
  for (int k = 0; k < level; ++k) t.layers[k].count = counts[k + 1];
  t.root = root;
  t.width = n0;
  t.built = true;
  t.plan_valid = false;

  // Give back over-provisioned storage (worst case was one item per position).
  if (t.n_leaves * 2 < t.leaves.count) {
    DevBuf<unsigned long long> exact;
    STB_CUDA(t, exact.alloc(t.n_leaves, st));
    STB_CUDA(t, cudaMemcpyAsync(exact.ptr, t.leaves.ptr, t.n_leaves * 8, cudaMemcpyDeviceToDevice, st));
    t.leaves = std::move(exact);
  }
  for (auto& layer : t.layers) {
    if (layer.count * 2 < layer.nodes.count) {
      DevBuf<uint2> exact;
      STB_CUDA(t, exact.alloc(layer.count, st));
      STB_CUDA(t, cudaMemcpyAsync(exact.ptr, layer.nodes.ptr, layer.count * 8, cudaMemcpyDeviceToDevice, st));
      layer.nodes = std::move(exact);
    }
  }
  return STB_OK;
}

int build_impl(Tree& t, const LeafInput& in, uint64_t n0, bool direct) {
  const int S = t.S;
  cudaStream_t st = t.stream;
  Scratch& sc = workspace_of(t);
  const uint64_t n1 = ceil_div(n0, 2);
  STB_CUDA(t, sc.ptr_a.ensure(n0, st));
  STB_CUDA(t, sc.ptr_b.ensure(n1, st));
  STB_CUDA(t, sc.bitmask.ensure(ceil_div(n0, LVL_TILE) * (LVL_TILE / 32), st));
  STB_CUDA(t, sc.blockcnt.ensure(ceil_div(n0, LVL_TILE) + 1, st));
  STB_CUDA(t, sc.chunk_sum.ensure(SCAN_MAX_CHUNKS, st));
  STB_CUDA(t, sc.counts.ensure(80, st));
  STB_CUDA(t, sc.flags.ensure(1, st));
  STB_CUDA(t, sc.root.ensure(1, st));
  STB_CUDA(t, sc.filter.ensure(filter_words(n1), st));
  {
    BuildFlags init{~0ull, 0u, 0u};
    STB_CUDA(t, cudaMemcpyAsync(sc.flags.ptr, &init, sizeof(init), cudaMemcpyHostToDevice, st));
  }

  // ---- leaf level ----
  const uint64_t direct_entries = direct ? (1ull << (2 * S)) : 0;
  const uint32_t leaf_cap = direct ? 0u : table_cap(n0);
  const uint32_t node_cap_max = table_cap(n1);
  {
    bool grew = false;
    STB_CUDA(t, sc.slots.ensure((uint64_t)std::max(leaf_cap, node_cap_max) + 1, st, &grew));
    if (grew) sc.tags_cleared = false;
  }
  LevelTable tab{sc.slots.ptr, nullptr, nullptr, leaf_cap};
  tab.first_bits = sc.bitmask.ptr;
  STB_CUDA(t, cudaMemsetAsync(sc.bitmask.ptr, 0, ceil_div(n0, LVL_TILE) * (LVL_TILE / 8), st));
  if (direct) {
    STB_CUDA(t, sc.dminpos.ensure(direct_entries, st));
    STB_CUDA(t, sc.dids.ensure(direct_entries, st));
    tab.dminpos = sc.dminpos.ptr;
    tab.dids = sc.dids.ptr;
    Launch l(t, "table_clear", false);
    STB_CUDA(t, cudaMemsetAsync(sc.dminpos.ptr, 0xff, direct_entries * 4, st));
  } else {
    Launch l(t, "table_clear", false);
    STB_CUDA(t, cudaMemsetAsync(sc.slots.ptr, 0xff, ((uint64_t)leaf_cap + 1) * sizeof(Slot), st));
  }
  const uint64_t leaf_store = direct ? std::min<uint64_t>(n0, direct_entries) : n0;
  STB_CUDA(t, t.leaves.alloc(leaf_store, st));

  if (in.body) {
    if (direct) {
      if (S == 12) launch_leaf_text<12, true>(t, in.body, n0, tab, sc.ptr_a.ptr, sc.flags.ptr);
      else launch_leaf_text<0, true>(t, in.body, n0, tab, sc.ptr_a.ptr, sc.flags.ptr);
    } else {
      if (S == 12) launch_leaf_text<12, false>(t, in.body, n0, tab, sc.ptr_a.ptr, sc.flags.ptr);
      else launch_leaf_text<0, false>(t, in.body, n0, tab, sc.ptr_a.ptr, sc.flags.ptr);
    }
  } else {
    Launch l(t, "leaf_insert");
    const unsigned nb = (unsigned)ceil_div(n0, LVL_THREADS);
    if (direct) leaf_insert_u64_kernel<true><<<nb, LVL_THREADS, 0, st>>>(in.leaves, (uint32_t)n0, S, tab, sc.ptr_a.ptr, sc.flags.ptr);
    else leaf_insert_u64_kernel<false><<<nb, LVL_THREADS, 0, st>>>(in.leaves, (uint32_t)n0, S, tab, sc.ptr_a.ptr, sc.flags.ptr);
  }
  if (direct) finish_level<MODE_LEAF_DIRECT>(t, sc.ptr_a.ptr, (uint32_t)n0, tab, sc, sc.counts.ptr, t.leaves.ptr);
  else finish_level<MODE_LEAF_HASH>(t, sc.ptr_a.ptr, (uint32_t)n0, tab, sc, sc.counts.ptr, t.leaves.ptr);

  // ---- node levels ----
  t.layers.clear();
  int level = 0;
  uint32_t* cur = nullptr;
  STB_TRY(run_node_levels(t, sc, sc.ptr_a.ptr, sc.ptr_b.ptr, n0, sc.counts.ptr + 1, &level, &cur));
  // `cur` now holds the single root pointer.

  return finish_build(t, sc, n0, level, cur, direct);
}

// ---- streaming build from host memory --------------------------------------------------------
// Ids are first-occurrence ranks in position order, and a later position can never displace an
// earlier one, so once a chunk of positions has been inserted its first-occurrence bits and ids
// are final.  The text is therefore copied in power-of-two chunks (the reference's own streaming
// unit, include/shared_tree.h:305-316) and every chunk runs through all the levels it spans while
// the next one is still on the PCIe bus; the scan of each level simply continues from the running
// total.  Only the last chunk's work and the small top of the tree remain after the copy ends.
int build_streaming(Tree& t, const char* h_body, uint64_t body_len) {
  const int S = t.S;
  const uint64_t n0 = body_len / (uint64_t)S;
  static const int chunk_log2 = (int)env_u64("STB_STREAM_CHUNK_LOG2", 24);
  static const uint64_t min_chunks = env_u64("STB_STREAM_MIN_CHUNKS", 4);
  const uint64_t C = 1ull << chunk_log2;
  if (S > 12 || chunk_log2 < 12 || n0 < min_chunks * C || n0 >= 0x7f000000ull) return -1;
  const int Lc = chunk_log2 - 11;  // chunked node levels: a chunk still holds 2048 positions at the last one
  cudaStream_t st = t.stream;
  t.clear();
  Scratch& sc = workspace_of(t);

  std::vector<uint64_t> n(Lc + 1), ptr_off(Lc + 2, 0), bit_off(Lc + 2, 0), cnt_off(Lc + 2, 0), slot_off(Lc + 2, 0);
  n[0] = n0;
  for (int j = 1; j <= Lc; ++j) n[j] = ceil_div(n[j - 1], 2);
  for (int j = 0; j <= Lc; ++j) {
    const uint64_t blocks = ceil_div(n[j], LVL_TILE);
    ptr_off[j + 1] = ptr_off[j] + blocks * LVL_TILE;
    bit_off[j + 1] = bit_off[j] + blocks * (LVL_TILE / 32);
    cnt_off[j + 1] = cnt_off[j] + blocks + 1;
    slot_off[j + 1] = slot_off[j] + (j == 0 ? 0 : (uint64_t)table_cap(n[j]) + 1);
  }
  const uint64_t direct_entries = 1ull << (2 * S);
  STB_CUDA(t, t.staging.ensure(n0 * (uint64_t)S + 16, st));
  STB_CUDA(t, sc.ptr_arena.ensure(ptr_off[Lc + 1], st));
  STB_CUDA(t, sc.bit_arena.ensure(bit_off[Lc + 1], st));
  STB_CUDA(t, sc.cnt_arena.ensure(cnt_off[Lc + 1], st));
  STB_CUDA(t, sc.level_sizes.ensure(Lc + 2, st));
  STB_CUDA(t, sc.counts.ensure(80, st));
  STB_CUDA(t, sc.flags.ensure(1, st));
  STB_CUDA(t, sc.root.ensure(1, st));
  STB_CUDA(t, sc.dminpos.ensure(direct_entries, st));
  STB_CUDA(t, sc.dids.ensure(direct_entries, st));
  {
    bool grew = false;
    STB_CUDA(t, sc.stream_slots.ensure(slot_off[Lc + 1], st, &grew));
    if (grew) {
      Launch l(t, "table_clear", false);
      STB_CUDA(t, cudaMemsetAsync(sc.stream_slots.ptr, 0xff, sc.stream_slots.bytes(), st));
    }
  }
  STB_CUDA(t, cudaMemsetAsync(sc.bit_arena.ptr, 0, bit_off[Lc + 1] * 4, st));
  STB_CUDA(t, cudaMemsetAsync(sc.counts.ptr, 0, 80 * 4, st));
  STB_CUDA(t, cudaMemsetAsync(sc.dminpos.ptr, 0xff, direct_entries * 4, st));
  {
    BuildFlags init{~0ull, 0u, 0u};
    STB_CUDA(t, cudaMemcpyAsync(sc.flags.ptr, &init, sizeof(init), cudaMemcpyHostToDevice, st));
    std::vector<uint32_t> sizes(Lc + 2, 0);
    for (int j = 0; j <= Lc; ++j) sizes[j] = (uint32_t)n[j];
    STB_CUDA(t, cudaMemcpyAsync(sc.level_sizes.ptr, sizes.data(), sizes.size() * 4, cudaMemcpyHostToDevice, st));
    STB_CUDA(t, cudaStreamSynchronize(st));  // `sizes` and `init` are stack memory
  }
  STB_CUDA(t, t.leaves.alloc(std::min<uint64_t>(n0, direct_entries), st));
  t.layers.clear();
  for (int j = 1; j <= Lc; ++j) {
    t.layers.emplace_back();
    STB_CUDA(t, t.layers.back().nodes.alloc(n[j], st));
  }
  std::vector<uint32_t> serial(Lc + 1, 0);
  for (int j = 1; j <= Lc; ++j) serial[j] = ++sc.serial;

  // all chunk copies are queued at once on their own stream; compute waits chunk by chunk
  if (!sc.copy_stream) STB_CUDA(t, cudaStreamCreateWithFlags(&sc.copy_stream, cudaStreamNonBlocking));
  const uint64_t chunks = ceil_div(n0, C);
  while (sc.chunk_events.size() < chunks) {
    cudaEvent_t e;
    STB_CUDA(t, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    sc.chunk_events.push_back(e);
  }
  char* text = t.staging.ptr;
  for (uint64_t c = 0; c < chunks; ++c) {
    const uint64_t first = c * C, cnt = std::min<uint64_t>(C, n0 - first);
    STB_CUDA(t, cudaMemcpyAsync(text + first * S, h_body + first * S, cnt * S, cudaMemcpyHostToDevice, sc.copy_stream));
    STB_CUDA(t, cudaEventRecord(sc.chunk_events[c], sc.copy_stream));
  }

  uint32_t* const ptrs = sc.ptr_arena.ptr;
  uint32_t* const bits = sc.bit_arena.ptr;
  uint32_t* const cnts = sc.cnt_arena.ptr;
  LevelTable leaf_tab{nullptr, sc.dminpos.ptr, sc.dids.ptr, 0u};
  leaf_tab.first_bits = bits;
  for (uint64_t c = 0; c < chunks; ++c) {
    const uint64_t first = c * C, cnt = std::min<uint64_t>(C, n0 - first);
    STB_CUDA(t, cudaStreamWaitEvent(st, sc.chunk_events[c], 0));
    // leaf level of this chunk
    if (S == 12) launch_leaf_text<12, true>(t, text + first * S, cnt, leaf_tab, ptrs + first, sc.flags.ptr, (uint32_t)first);
    else launch_leaf_text<0, true>(t, text + first * S, cnt, leaf_tab, ptrs + first, sc.flags.ptr, (uint32_t)first);
    for (int j = 0; j <= Lc; ++j) {
      const uint64_t begin = first >> j, end = ceil_div(first + cnt, 1ull << j);
      const uint32_t fb = (uint32_t)(begin / LVL_TILE), nbk = (uint32_t)ceil_div(end - begin, LVL_TILE);
      uint32_t* lvl_ptr = ptrs + ptr_off[j];
      uint32_t* lvl_bits = bits + bit_off[j];
      uint32_t* lvl_cnt = cnts + cnt_off[j];
      LevelTable tab = leaf_tab;
      if (j > 0) {
        tab = LevelTable{sc.stream_slots.ptr + slot_off[j], nullptr, nullptr, table_cap(n[j])};
        tab.first_bits = lvl_bits;
        Launch l(t, "node_insert");
        // placement by child id above the first node layer; child ids are bounded by the child level's size
        const uint32_t* child_unique = j > 1 ? sc.level_sizes.ptr + (j - 1) : nullptr;
        node_insert_kernel<<<nbk, LVL_THREADS, 0, st>>>(ptrs + ptr_off[j - 1], (uint32_t)n[j - 1], (uint32_t)end, tab, lvl_ptr, child_unique,
                                                         serial[j], SingletonFilter{}, fb);
      }
      {
        Launch l(t, "bitmask_blockcnt");
        bitmask_blockcnt_kernel<<<(unsigned)ceil_div((uint64_t)nbk * 32, 256), 256, 0, st>>>(lvl_bits + (uint64_t)fb * (LVL_TILE / 32), nbk, lvl_cnt + fb);
      }
      {
        Launch l(t, "scan_blocks");
        scan_blocks_kernel<<<1, 1024, 0, st>>>(lvl_cnt + fb, nbk, sc.counts.ptr + j, sc.counts.ptr + j);
      }
      {
        Launch l(t, "assign_ids");
        if (j == 0)
          assign_kernel<MODE_LEAF_DIRECT><<<nbk, LVL_THREADS, 0, st>>>(lvl_ptr, (uint32_t)n[0], tab, lvl_bits, lvl_cnt, t.leaves.ptr, S, nullptr, 0u, fb);
        else
          assign_kernel<MODE_NODE><<<nbk, LVL_THREADS, 0, st>>>(lvl_ptr, (uint32_t)end, tab, lvl_bits, lvl_cnt, t.layers[j - 1].nodes.ptr, S,
                                                                ptrs + ptr_off[j - 1], (uint32_t)n[j - 1], fb);
      }
      {
        Launch l(t, "resolve_ids");
        if (j == 0) resolve_kernel<true><<<nbk, LVL_THREADS, 0, st>>>(lvl_ptr, (uint32_t)std::min<uint64_t>(n[0], first + cnt), tab, lvl_bits, fb);
        else resolve_kernel<false><<<nbk, LVL_THREADS, 0, st>>>(lvl_ptr, (uint32_t)end, tab, lvl_bits, fb);
      }
    }
  }
  // the top of the tree: everything above the last chunked level, as in the one-shot build
  const uint64_t n_top = n[Lc];
  STB_CUDA(t, sc.ptr_a.ensure(n_top, st));
  STB_CUDA(t, sc.ptr_b.ensure(ceil_div(n_top, 2), st));
  STB_CUDA(t, sc.bitmask.ensure(ceil_div(n_top, LVL_TILE) * (LVL_TILE / 32), st));
  STB_CUDA(t, sc.blockcnt.ensure(ceil_div(n_top, LVL_TILE) + 1, st));
  STB_CUDA(t, sc.chunk_sum.ensure(SCAN_MAX_CHUNKS, st));
  STB_CUDA(t, sc.filter.ensure(filter_words(ceil_div(n_top, 2)), st));
  {
    bool grew = false;
    STB_CUDA(t, sc.slots.ensure((uint64_t)table_cap(ceil_div(n_top, 2)) + 1, st, &grew));
    if (grew) sc.tags_cleared = false;
  }
  STB_CUDA(t, cudaMemcpyAsync(sc.ptr_a.ptr, ptrs + ptr_off[Lc], n_top * 4, cudaMemcpyDeviceToDevice, st));
  int more = 0;
  uint32_t* cur = nullptr;
  STB_TRY(run_node_levels(t, sc, sc.ptr_a.ptr, sc.ptr_b.ptr, n_top, sc.counts.ptr + 1 + Lc, &more, &cur));
  return finish_build(t, sc, n0, Lc + more, cur, true);
}

int build_dispatch(Tree& t, const LeafInput& in, uint64_t n0) {
  if (n0 == 0) return t.fail(STB_ERR_EMPTY, "input holds fewer than dna_size bases");
  if (n0 >= 0xffffffffull) return t.fail(STB_ERR_TOO_LARGE, "more than 2^32-2 leaf positions");
  t.clear();
  const bool try_direct = t.S <= 12;
  if (try_direct) {
    const int s = build_impl(t, in, n0, true);
    if (s != -1) return s;
  }
  if (n0 > 480000000ull)
    return t.fail(STB_ERR_TOO_LARGE, "hash-table leaf level supports at most 480M leaf positions in this version");
  return build_impl(t, in, n0, false);
}

}  // namespace

// Small top of a sharded build: node layers only, from a gathered pointer array.
int build_upper_levels(Tree& t, const uint32_t* d_ptrs, uint64_t n, bool at_least_one) {
  if (n == 0) return t.fail(STB_ERR_EMPTY, "no pointers");
  if (n > 400000000ull) return t.fail(STB_ERR_TOO_LARGE, "upper levels: too many pointers");
  t.clear();
  cudaStream_t st = t.stream;
  if (n == 1 && !at_least_one) {  // already the root
    uint32_t root = PTR_NULL;
    STB_CUDA(t, cudaMemcpyAsync(&root, d_ptrs, 4, cudaMemcpyDeviceToHost, st));
    STB_CUDA(t, cudaStreamSynchronize(st));
    t.root = root;
    t.built = true;
    return STB_OK;
  }
  Scratch& sc = workspace_of(t);
  const uint64_t n1 = ceil_div(n, 2);
  STB_CUDA(t, sc.ptr_a.ensure(n, st));
  STB_CUDA(t, sc.ptr_b.ensure(n1, st));
  STB_CUDA(t, sc.bitmask.ensure(ceil_div(n1, LVL_TILE) * (LVL_TILE / 32), st));
  STB_CUDA(t, sc.blockcnt.ensure(ceil_div(n1, LVL_TILE) + 1, st));
  STB_CUDA(t, sc.chunk_sum.ensure(SCAN_MAX_CHUNKS, st));
  STB_CUDA(t, sc.counts.ensure(80, st));
  STB_CUDA(t, sc.filter.ensure(filter_words(n1), st));
  {
    bool grew = false;
    STB_CUDA(t, sc.slots.ensure((uint64_t)table_cap(n1) + 1, st, &grew));
    if (grew) sc.tags_cleared = false;
  }
  STB_CUDA(t, cudaMemcpyAsync(sc.ptr_a.ptr, d_ptrs, n * 4, cudaMemcpyDeviceToDevice, st));
  int level = 0;
  uint32_t* cur = nullptr;
  STB_TRY(run_node_levels(t, sc, sc.ptr_a.ptr, sc.ptr_b.ptr, n, sc.counts.ptr, &level, &cur));
  std::vector<uint32_t> counts(level);
  uint32_t root = PTR_NULL;
  STB_CUDA(t, cudaMemcpyAsync(counts.data(), sc.counts.ptr, counts.size() * 4, cudaMemcpyDeviceToHost, st));
  STB_CUDA(t, cudaMemcpyAsync(&root, cur, 4, cudaMemcpyDeviceToHost, st));
  STB_CUDA(t, cudaStreamSynchronize(st));
  STB_CUDA(t, cudaGetLastError());
  for (int k = 0; k < level; ++k) t.layers[k].count = counts[k];
  t.root = root;
  t.built = true;
  return STB_OK;
}

// ---- multi-GPU leaf level on the direct-addressed table (dist.cu drives it) ---------------
// Step 1: this rank's leaves lower the replicated table with GLOBAL positions.
int dist_leaf_direct_minpos(Ctx& ctx, const char* d_body, uint64_t n_local, uint64_t gpos0, uint32_t* dminpos, uint32_t* tmp,
                            int* non_acgt) {
  *non_acgt = 0;
  if (n_local == 0) return STB_OK;
  DevBuf<BuildFlags> flags;
  STB_CUDA(ctx, flags.alloc(1, ctx.stream));
  BuildFlags init{~0ull, 0u, 0u};
  STB_CUDA(ctx, cudaMemcpyAsync(flags.ptr, &init, sizeof(init), cudaMemcpyHostToDevice, ctx.stream));
  LevelTable tab{nullptr, dminpos, nullptr, 0u};
  if (ctx.S == 12) launch_leaf_text<12, true>(ctx, d_body, n_local, tab, tmp, flags.ptr, (uint32_t)gpos0);
  else launch_leaf_text<0, true>(ctx, d_body, n_local, tab, tmp, flags.ptr, (uint32_t)gpos0);
  BuildFlags h{};
  STB_CUDA(ctx, cudaMemcpyAsync(&h, flags.ptr, sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
  STB_CUDA(ctx, cudaStreamSynchronize(ctx.stream));
  STB_CUDA(ctx, cudaGetLastError());
  if (h.bad_symbol != ~0ull) return ctx.fail(STB_ERR_UNKNOWN_SYMBOL, unknown_symbol_message((int)(h.bad_symbol & 0xff)));
  *non_acgt = (int)h.non_acgt;
  return STB_OK;
}

int build_from_body(Tree& t, const char* d_body, uint64_t body_len) {
  LeafInput in;
  in.body = d_body;
  return build_dispatch(t, in, body_len / (uint64_t)t.S);
}

// Host text: overlap the copy with the build when the input is large enough; -1 = not applicable
// (small input, dna_size > 12, or non-ACGT symbols), the caller copies and builds in one shot.
int build_from_host_body(Tree& t, const char* h_body, uint64_t body_len) { return build_streaming(t, h_body, body_len); }

int build_from_leaves(Tree& t, const unsigned long long* d_leaves, uint64_t n) {
  LeafInput in;
  in.leaves = d_leaves;
  return build_dispatch(t, in, n);
}

}  // namespace stb
