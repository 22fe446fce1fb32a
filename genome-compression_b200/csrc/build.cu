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

#include <cooperative_groups.h>

#include "pack.cuh"
#include "tree.h"

namespace cg = cooperative_groups;

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

// One node level: reduce_nodes + emplace_node (shared_tree.cpp:697-712, :662-672), for the
// positions [p_begin, p_end) of the level.
template <bool PROBE_FIRST>
__global__ void __launch_bounds__(LVL_THREADS)
node_insert_kernel(const uint32_t* __restrict__ cur, uint32_t n_cur, uint32_t p_begin, uint32_t p_end, LevelTable tab,
                   uint32_t* __restrict__ tmp, const uint32_t* __restrict__ child_unique) {
  const uint32_t p = p_begin + blockIdx.x * LVL_THREADS + threadIdx.x;
  if (p >= p_end) return;
  uint32_t l, r;
  if (2 * (uint64_t)p + 1 < n_cur) {
    const uint2 pr = __ldg(reinterpret_cast<const uint2*>(cur) + p);
    l = pr.x;
    r = pr.y;
  } else {  // odd tail: node{last, nullptr} (utility.h:17-29)
    l = cur[2 * (uint64_t)p];
    r = PTR_NULL;
  }
  uint32_t cl, cr, f;
  canonical_node(l, r, cl, cr, f);
  const unsigned long long key = ((unsigned long long)cl << 32) | cr;
  const uint32_t hashed = __umulhi(hash64(key), tab.cap);
  uint32_t s;
  if (child_unique) {
    // Locality placement (node layers above the first): child ids are first-occurrence ranks, so
    // they grow with the position; a slot proportional to a child id makes neighbouring positions
    // probe neighbouring slots (one 128-byte line serves several positions instead of one line
    // per position).  Crowded neighbourhoods (one child with many partners) fall back to the hash.
    const uint32_t child = ptr_is_null(cl) ? (cr & IDX_MASK) : (cl & IDX_MASK);
    const uint32_t unique = max(1u, __ldg(child_unique));
    uint32_t near = (uint32_t)(((unsigned long long)child * tab.cap) / unique) + (hashed & 7u);
    if (near >= tab.cap) near = tab.cap - 1;
    s = table_insert_from<PROBE_FIRST>(tab.slots, tab.cap, key, p, near, tab.first_bits, hashed, 24u);
  } else {
    s = table_insert_from<PROBE_FIRST>(tab.slots, tab.cap, key, p, hashed, tab.first_bits);
  }
  tmp[p] = s | f;
}

// ---- hash-partitioned node level ----------------------------------------------------------
// A random slot access costs a whole 128-byte line of HBM traffic (profiles/README.md), so a
// large level is first split by key hash into buckets small enough that a bucket's table stays
// in L2; the (key, position) records then stream through HBM once, 12 bytes each, and the
// table itself never leaves the cache.  Buckets are processed in batches of PART_BATCH tables.
constexpr int PART_THREADS = 256;
constexpr int PART_WARPS = PART_THREADS / 32;
constexpr int PART_GROUPS = 16;                          // 32-record groups per warp
constexpr int PART_TILE = PART_THREADS * PART_GROUPS;    // 4096 positions per CTA
constexpr int PART_WARP_ITEMS = 32 * PART_GROUPS;
constexpr int PART_MAX_BUCKETS = 512;

__device__ __forceinline__ void node_key_at(const uint32_t* __restrict__ cur, uint32_t n_cur, uint32_t p, unsigned long long& key,
                                            uint32_t& flags) {
  uint32_t l, r;
  if (2 * (uint64_t)p + 1 < n_cur) {
    const uint2 pr = __ldg(reinterpret_cast<const uint2*>(cur) + p);
    l = pr.x;
    r = pr.y;
  } else {
    l = cur[2 * (uint64_t)p];
    r = PTR_NULL;
  }
  uint32_t cl, cr;
  canonical_node(l, r, cl, cr, flags);
  key = ((unsigned long long)cl << 32) | cr;
}

// bucket = top bits of one hash, slot inside the bucket's table = another hash
__device__ __forceinline__ uint32_t bucket_of(unsigned long long key, int log2_buckets) {
  return log2_buckets ? hash64(key) >> (32 - log2_buckets) : 0u;
}

__global__ void __launch_bounds__(PART_THREADS)
part_hist_kernel(const uint32_t* __restrict__ cur, uint32_t n_cur, uint32_t n_pos, int log2_buckets, uint32_t nblocks,
                 uint32_t* __restrict__ hist) {
  extern __shared__ uint32_t bins[];
  const uint32_t buckets = 1u << log2_buckets;
  for (uint32_t b = threadIdx.x; b < buckets; b += PART_THREADS) bins[b] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t warp_first = blockIdx.x * PART_TILE + warp * PART_WARP_ITEMS;
#pragma unroll 4
  for (int g = 0; g < PART_GROUPS; ++g) {
    const uint32_t p = warp_first + g * 32 + lane;
    uint32_t b = 0xffffffffu;
    if (p < n_pos) {
      unsigned long long key;
      uint32_t f;
      node_key_at(cur, n_cur, p, key, f);
      b = bucket_of(key, log2_buckets);
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, b);
    if (p < n_pos && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&bins[b], (uint32_t)__popc(peers));
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < buckets; b += PART_THREADS) hist[b * nblocks + blockIdx.x] = bins[b];
}

// one CTA per bucket: exclusive scan of its row of per-CTA counts + bucket total
__global__ void __launch_bounds__(1024) part_rowscan_kernel(uint32_t* __restrict__ hist, uint32_t nblocks, uint32_t* __restrict__ row_total) {
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

// bucket_off[b] = records before bucket b (single CTA, <= 1024 buckets); bucket_off[buckets] = total
__global__ void __launch_bounds__(1024) part_offsets_kernel(const uint32_t* __restrict__ row_total, uint32_t buckets,
                                                            uint32_t* __restrict__ bucket_off) {
  __shared__ uint32_t warp_sum[32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t v = threadIdx.x < buckets ? row_total[threadIdx.x] : 0u;
  uint32_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) warp_sum[warp] = x;
  __syncthreads();
  uint32_t before = 0;
  for (uint32_t w = 0; w < warp; ++w) before += warp_sum[w];
  if (threadIdx.x < buckets) bucket_off[threadIdx.x] = before + x - v;
  if (threadIdx.x == buckets - 1) bucket_off[buckets] = before + x;
}

// Scatter pass.  The tile's records are first ordered by bucket in shared memory, then written
// out: consecutive threads write consecutive records of one bucket, so the stores leave the SM as
// runs instead of one 32-byte sector per lane (the un-staged version was L2-request bound: 8.3 ms
// instead of ~1.5 ms per 3.1 Gbp).
struct PartSmem {
  unsigned long long key[PART_TILE];
  uint32_t meta[PART_TILE];
  uint32_t tile_off[PART_MAX_BUCKETS + 1];   // tile-local exclusive prefix of bucket counts
  uint32_t gbase[PART_MAX_BUCKETS];          // global index of this tile's first record of bucket b
  uint32_t warp_cnt[PART_WARPS * PART_MAX_BUCKETS];
};

__global__ void __launch_bounds__(PART_THREADS)
part_scatter_kernel(const uint32_t* __restrict__ cur, uint32_t n_cur, uint32_t n_pos, int log2_buckets, uint32_t nblocks,
                    const uint32_t* __restrict__ hist, const uint32_t* __restrict__ bucket_off,
                    unsigned long long* __restrict__ rec_key, uint32_t* __restrict__ rec_meta) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  PartSmem& sm = *reinterpret_cast<PartSmem*>(smem_raw);
  const uint32_t buckets = 1u << log2_buckets;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < PART_WARPS * buckets; i += PART_THREADS) sm.warp_cnt[i] = 0;
  __syncthreads();
  const uint32_t tile_first = blockIdx.x * PART_TILE;
  const uint32_t warp_first = tile_first + warp * PART_WARP_ITEMS;
  unsigned long long key[PART_GROUPS];
  uint32_t meta[PART_GROUPS], bkt[PART_GROUPS];
  uint32_t* mine = sm.warp_cnt + warp * buckets;
#pragma unroll
  for (int g = 0; g < PART_GROUPS; ++g) {
    const uint32_t p = warp_first + g * 32 + lane;
    bkt[g] = 0xffffffffu;
    key[g] = 0;
    meta[g] = 0;
    if (p < n_pos) {
      uint32_t f;
      node_key_at(cur, n_cur, p, key[g], f);
      meta[g] = p | f;
      bkt[g] = bucket_of(key[g], log2_buckets);
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, bkt[g]);
    if (p < n_pos && lane == (uint32_t)(__ffs(peers) - 1)) mine[bkt[g]] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // per bucket: tile count, exclusive scan over the warps (tile-local offsets), global base
  for (uint32_t b = threadIdx.x; b < buckets; b += PART_THREADS) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < PART_WARPS; ++w) {
      const uint32_t c = sm.warp_cnt[w * buckets + b];
      sm.warp_cnt[w * buckets + b] = run;
      run += c;
    }
    sm.tile_off[b] = run;  // count for now
    sm.gbase[b] = bucket_off[b] + hist[b * nblocks + blockIdx.x];
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of the bucket counts (<= 1024 entries, 32 per lane)
    const uint32_t per = (buckets + 31) / 32;
    uint32_t sum = 0;
    for (uint32_t j = 0; j < per; ++j) {
      const uint32_t b = lane * per + j;
      if (b < buckets) sum += sm.tile_off[b];
    }
    uint32_t x = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    uint32_t run = x - sum;
    for (uint32_t j = 0; j < per; ++j) {
      const uint32_t b = lane * per + j;
      if (b < buckets) {
        const uint32_t c = sm.tile_off[b];
        sm.tile_off[b] = run;
        run += c;
      }
    }
    if (lane == 31) sm.tile_off[buckets] = run;
  }
  __syncthreads();
  // stage: record -> tile-local slot (bucket order, stable inside a bucket)
#pragma unroll
  for (int g = 0; g < PART_GROUPS; ++g) {
    const uint32_t p = warp_first + g * 32 + lane;
    const bool ok = p < n_pos;
    const uint32_t peers = __match_any_sync(0xffffffffu, bkt[g]);
    uint32_t li = 0;
    if (ok) li = sm.tile_off[bkt[g]] + mine[bkt[g]] + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
    if (ok && lane == (uint32_t)(__ffs(peers) - 1)) mine[bkt[g]] += __popc(peers);
    __syncwarp();
    if (ok) {
      sm.key[li] = key[g];
      sm.meta[li] = meta[g];
    }
  }
  __syncthreads();
  // write out in staged order: runs of one bucket are contiguous in HBM
  const uint32_t here = sm.tile_off[buckets];
  for (uint32_t j = threadIdx.x; j < here; j += PART_THREADS) {
    const unsigned long long k = sm.key[j];
    const uint32_t b = bucket_of(k, log2_buckets);
    const uint32_t dst = sm.gbase[b] + (j - sm.tile_off[b]);
    rec_key[dst] = k;
    rec_meta[dst] = sm.meta[j];
  }
}

// 128-bit CAS with an arbitrary expected value; returns the previous contents.
__device__ __forceinline__ void cas_slot(Slot* s, unsigned long long exp_lo, unsigned long long exp_hi, unsigned long long new_lo,
                                         unsigned long long new_hi, unsigned long long& old_lo, unsigned long long& old_hi) {
  asm volatile(
      "{\n\t"
      ".reg .b128 cmp, val, old;\n\t"
      "mov.b128 cmp, {%2, %3};\n\t"
      "mov.b128 val, {%4, %5};\n\t"
      "atom.global.cas.b128 old, [%6], cmp, val;\n\t"
      "mov.b128 {%0, %1}, old;\n\t"
      "}"
      : "=l"(old_lo), "=l"(old_hi)
      : "l"(exp_lo), "l"(exp_hi), "l"(new_lo), "l"(new_hi), "l"(s)
      : "memory");
}

// Insert into a table whose slots carry an epoch tag in their last word: a slot whose tag is not
// `serial` is stale, i.e. empty, so tables are reused batch after batch without being cleared.
__device__ __forceinline__ uint32_t tagged_insert(Slot* tab, uint32_t cap, unsigned long long key, uint32_t pos, uint32_t serial) {
  uint32_t s = __umulhi((uint32_t)mix64(key), cap);
  const unsigned long long fresh_hi = ((unsigned long long)serial << 32) | pos;
  for (;;) {
    unsigned long long k, w;
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(k), "=l"(w) : "l"(tab + s));
    if ((uint32_t)(w >> 32) != serial) {
      unsigned long long ok, ow;
      cas_slot(tab + s, k, w, key, fresh_hi, ok, ow);
      if (ok == k && ow == w) return s;  // claimed with key and min-position in one atomic
      k = ok;
      w = ow;  // somebody else claimed it in this epoch
    }
    if (k == key) {
      if ((uint32_t)w > pos) atomicMin(&tab[s].minpos, pos);
      return s;
    }
    if (++s == cap) s = 0;
  }
}

constexpr int BK_PER_THREAD = 2;                 // independent probe chains in flight per thread
constexpr int BK_TILE = 256 * BK_PER_THREAD;     // records per CTA

// Bucket b owns the table region [2*off[b], 2*off[b+1]) (twice its record count), so the whole
// level's tables are exactly as large as the un-partitioned table and no bucket can overflow,
// whatever the skew.  Records are stored bucket after bucket and CTAs are dispatched in index
// order, so at any moment only a few buckets' regions are being touched; each CTA first prefetches
// its share of its bucket's region into L2 (a streaming read) so the random probes that follow
// hit L2 instead of each missing to HBM.
__device__ __forceinline__ void prefetch_table_share(const Slot* tables, const uint32_t* __restrict__ bucket_off, uint32_t buckets,
                                                     uint32_t b, uint32_t first_record) {
  // The CTAs of bucket b run together; what they stream in is the region of bucket b + 1, whose
  // CTAs come next, so that bucket's probes find their lines already in L2.
  const uint32_t begin = bucket_off[b], count = bucket_off[b + 1] - begin;
  const uint32_t blocks_in_bucket = (count + BK_TILE - 1) / BK_TILE;
  const uint32_t rank = (first_record - begin) / BK_TILE;
  if (b + 1 >= buckets) return;
  const uint32_t next_begin = bucket_off[b + 1], next_count = bucket_off[b + 2] - next_begin;
  const uint32_t lines = (next_count * 2 * (uint32_t)sizeof(Slot) + 127) / 128;  // 128-byte lines of the region
  const uint32_t per_block = (lines + blocks_in_bucket - 1) / blocks_in_bucket;
  const char* region = reinterpret_cast<const char*>(tables + 2 * (size_t)next_begin);
  for (uint32_t j = threadIdx.x; j < per_block; j += 256) {
    const uint32_t line = rank * per_block + j;
    if (line < lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(region + (size_t)line * 128));
  }
}

__global__ void __launch_bounds__(256)
bucket_insert_kernel(const unsigned long long* __restrict__ rec_key, const uint32_t* __restrict__ rec_meta, uint32_t n,
                     const uint32_t* __restrict__ bucket_off, int log2_buckets, Slot* tables, uint32_t serial,
                     uint32_t* __restrict__ rec_slot) {
  const uint32_t i0 = blockIdx.x * BK_TILE;
  prefetch_table_share(tables, bucket_off, 1u << log2_buckets, bucket_of(rec_key[i0], log2_buckets), i0);
  unsigned long long key[BK_PER_THREAD];
  uint32_t pos[BK_PER_THREAD];
#pragma unroll
  for (int j = 0; j < BK_PER_THREAD; ++j) {
    const uint32_t i = i0 + j * 256 + threadIdx.x;
    key[j] = i < n ? rec_key[i] : 0ull;
    pos[j] = i < n ? rec_meta[i] & IDX_MASK : 0u;
  }
  Slot* tab[BK_PER_THREAD];
  uint32_t cap[BK_PER_THREAD];
#pragma unroll
  for (int j = 0; j < BK_PER_THREAD; ++j) {
    const uint32_t b = bucket_of(key[j], log2_buckets);
    const uint32_t begin = __ldg(bucket_off + b);
    cap[j] = 2 * (__ldg(bucket_off + b + 1) - begin);
    tab[j] = tables + 2 * (size_t)begin;
  }
#pragma unroll
  for (int j = 0; j < BK_PER_THREAD; ++j) {
    const uint32_t i = i0 + j * 256 + threadIdx.x;
    if (i < n) rec_slot[i] = tagged_insert(tab[j], cap[j], key[j], pos[j], serial);
  }
}

__global__ void __launch_bounds__(256)
bucket_answer_kernel(const unsigned long long* __restrict__ rec_key, const uint32_t* __restrict__ rec_meta, uint32_t n,
                     const uint32_t* __restrict__ rec_slot, const uint32_t* __restrict__ bucket_off, int log2_buckets,
                     const Slot* tables, uint32_t* __restrict__ bitmask, uint32_t* __restrict__ tmp) {
  const uint32_t i0 = blockIdx.x * BK_TILE;
  prefetch_table_share(tables, bucket_off, 1u << log2_buckets, bucket_of(rec_key[i0], log2_buckets), i0);
  uint32_t meta[BK_PER_THREAD], q[BK_PER_THREAD];
  const uint32_t* where[BK_PER_THREAD];
#pragma unroll
  for (int j = 0; j < BK_PER_THREAD; ++j) {
    const uint32_t i = i0 + j * 256 + threadIdx.x;
    meta[j] = 0;
    where[j] = nullptr;
    if (i < n) {
      meta[j] = rec_meta[i];
      const uint32_t begin = __ldg(bucket_off + bucket_of(rec_key[i], log2_buckets));
      where[j] = &tables[2 * (size_t)begin + rec_slot[i]].minpos;
    }
  }
#pragma unroll
  for (int j = 0; j < BK_PER_THREAD; ++j) q[j] = where[j] ? __ldcg(where[j]) : 0u;
#pragma unroll
  for (int j = 0; j < BK_PER_THREAD; ++j) {
    if (!where[j]) continue;
    const uint32_t pos = meta[j] & IDX_MASK;
    if (q[j] == pos) atomicOr(bitmask + (pos >> 5), 1u << (pos & 31));
    else tmp[pos] = (meta[j] & ~IDX_MASK) | q[j];
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

// In-place exclusive scan of the per-CTA counts (single CTA); total -> *total_out.
__global__ void __launch_bounds__(1024) scan_blocks_kernel(uint32_t* __restrict__ cnt, uint32_t nb,
                                                           uint32_t* __restrict__ total_out) {
  __shared__ uint32_t warp_sum[32];
  __shared__ uint32_t carry_s;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nb; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < nb ? cnt[i] : 0u;
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
      warp_sum[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    const uint32_t before = carry + (warp ? warp_sum[warp - 1] : 0u) + x - v;
    if (i < nb) cnt[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry_s;
}

enum { MODE_LEAF_DIRECT = 0, MODE_LEAF_HASH = 1, MODE_NODE = 2 };

// First occurrences: id = rank in position order; append the item to its layer in id order
// and emit the finished pointer.  Node items are recomputed from the child pointers
// (coalesced) rather than fetched from the table (random).
template <int MODE>
__global__ void __launch_bounds__(LVL_THREADS)
assign_kernel(uint32_t* __restrict__ tmp, uint32_t n, LevelTable tab, const uint32_t* __restrict__ bitmask,
              const uint32_t* __restrict__ blockbase, void* __restrict__ uniq, int S,
              const uint32_t* __restrict__ children, uint32_t n_children) {
  __shared__ uint32_t word_pref[32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t words[LVL_ITERS];
#pragma unroll
  for (int it = 0; it < LVL_ITERS; ++it) {
    const uint32_t p0 = blockIdx.x * LVL_TILE + it * LVL_THREADS + warp * 32;
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
  const uint32_t base = blockbase[blockIdx.x];
#pragma unroll
  for (int it = 0; it < LVL_ITERS; ++it) {
    const uint32_t p = blockIdx.x * LVL_TILE + it * LVL_THREADS + threadIdx.x;
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
resolve_kernel(uint32_t* __restrict__ tmp, uint32_t n, LevelTable tab, const uint32_t* __restrict__ bitmask, bool via_position) {
  const uint32_t lane = threadIdx.x & 31;
#pragma unroll
  for (int it = 0; it < LVL_ITERS; ++it) {
    const uint32_t p = blockIdx.x * LVL_TILE + it * LVL_THREADS + threadIdx.x;
    if (p < n) {
      const uint32_t word = bitmask[p >> 5];
      if (!((word >> lane) & 1u)) {
        const uint32_t t = tmp[p];
        const uint32_t s = t & IDX_MASK;
        // hash levels: slot -> position of the first occurrence -> its finished pointer -> id
        // (the partitioned path already left the position instead of the slot: via_position)
        const uint32_t id = DIRECT ? __ldcg(tab.dids + s)
                                   : (__ldcg(tmp + (via_position ? s : __ldcg(&tab.slots[s].minpos))) & IDX_MASK);
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
  DevBuf<uint32_t> ptr_a, ptr_b, bitmask, blockcnt, counts, dminpos, dids;
  DevBuf<Slot> slots;
  DevBuf<BuildFlags> flags;
  DevBuf<uint32_t> root;
  bool tags_cleared = false;  // partitioned levels: slots carry an epoch tag, cleared once per build
  uint32_t serial = 0;
};

uint32_t table_cap(uint64_t n) { return (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1024, 2 * n), 0x1ffffffeull); }

template <int S_T, bool DIRECT>
void launch_leaf_text(Ctx& ctx, const char* body, uint64_t n, LevelTable tab, uint32_t* tmp, BuildFlags* flags, uint32_t pos0 = 0) {
  const size_t smem = pack_smem_bytes(ctx.S);
  Launch l(ctx, "leaf_insert");
  leaf_insert_text_kernel<S_T, DIRECT><<<(unsigned)ceil_div(n, PACK_TILE_LEAVES), PACK_THREADS, smem, ctx.stream>>>(
      body, n, ctx.S, tab, tmp, flags, pos0);
}

// count -> scan -> assign -> resolve for one level whose inserts are already queued.
template <int MODE>
void finish_level(Ctx& ctx, uint32_t* tmp, uint32_t n, LevelTable tab, Scratch& sc, uint32_t* total_out, void* uniq,
                  const uint32_t* children = nullptr, uint32_t n_children = 0, bool via_position = false) {
  constexpr bool DIRECT = (MODE == MODE_LEAF_DIRECT);
  const unsigned nb = (unsigned)ceil_div(n, LVL_TILE);
  {  // the inserts kept the first-occurrence bitmap current: count it per CTA tile for the scan
    Launch l(ctx, "bitmask_blockcnt");
    bitmask_blockcnt_kernel<<<(unsigned)ceil_div((uint64_t)nb * 32, 256), 256, 0, ctx.stream>>>(sc.bitmask.ptr, nb, sc.blockcnt.ptr);
  }
  {
    Launch l(ctx, "scan_blocks");
    scan_blocks_kernel<<<1, 1024, 0, ctx.stream>>>(sc.blockcnt.ptr, nb, total_out);
  }
  {
    Launch l(ctx, "assign_ids");
    assign_kernel<MODE><<<nb, LVL_THREADS, 0, ctx.stream>>>(tmp, n, tab, sc.bitmask.ptr, sc.blockcnt.ptr, uniq, ctx.S, children, n_children);
  }
  {
    Launch l(ctx, "resolve_ids");
    resolve_kernel<DIRECT><<<nb, LVL_THREADS, 0, ctx.stream>>>(tmp, n, tab, sc.bitmask.ptr, via_position);
  }
}

// Tunables of the partitioned path (environment overrides are for experiments only).
static uint64_t env_u64(const char* name, uint64_t fallback) {
  const char* v = getenv(name);
  return v ? strtoull(v, nullptr, 0) : fallback;
}

// insert + classify of one large node level through hash buckets with L2-resident tables.
// Leaves the level's first-occurrence bitmask set and, for later occurrences, the position of
// the first one in tmp[] (finish_level is told via_position).
int partitioned_insert_count(Tree& t, Scratch& sc, const uint32_t* cur, uint64_t n_cur, uint32_t* nxt, uint64_t n_next) {
  cudaStream_t st = t.stream;
  static const uint64_t bucket_target = env_u64("STB_PART_BUCKET", 1ull << 18);
  int log2p = 0;
  while (log2p < 9 && (n_next >> log2p) > bucket_target) ++log2p;
  const uint32_t buckets = 1u << log2p;
  const uint32_t nblocks = (uint32_t)ceil_div(n_next, PART_TILE);
  DevBuf<uint32_t> hist, row_total, bucket_off, rec_meta, rec_slot;
  DevBuf<unsigned long long> rec_key;
  STB_CUDA(t, hist.alloc((uint64_t)buckets * nblocks, st));
  STB_CUDA(t, row_total.alloc(buckets, st));
  STB_CUDA(t, bucket_off.alloc(buckets + 1, st));
  STB_CUDA(t, rec_key.alloc(n_next, st));
  STB_CUDA(t, rec_meta.alloc(n_next, st));
  STB_CUDA(t, rec_slot.alloc(n_next, st));
  static bool attr_set = false;
  if (!attr_set) {
    STB_CUDA(t, cudaFuncSetAttribute(part_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PartSmem)));
    attr_set = true;
  }
  {
    Launch l(t, "part_hist");
    part_hist_kernel<<<nblocks, PART_THREADS, buckets * 4, st>>>(cur, (uint32_t)n_cur, (uint32_t)n_next, log2p, nblocks, hist.ptr);
  }
  {
    Launch l(t, "part_scan");
    part_rowscan_kernel<<<buckets, 1024, 0, st>>>(hist.ptr, nblocks, row_total.ptr);
    part_offsets_kernel<<<1, 1024, 0, st>>>(row_total.ptr, buckets, bucket_off.ptr);
  }
  {
    Launch l(t, "part_scatter");
    part_scatter_kernel<<<nblocks, PART_THREADS, sizeof(PartSmem), st>>>(cur, (uint32_t)n_cur, (uint32_t)n_next, log2p, nblocks, hist.ptr,
                                                                        bucket_off.ptr, rec_key.ptr, rec_meta.ptr);
  }
  const uint64_t mask_words = ceil_div(n_next, LVL_TILE) * (LVL_TILE / 32);
  STB_CUDA(t, cudaMemsetAsync(sc.bitmask.ptr, 0, mask_words * 4, st));
  if (!sc.tags_cleared) {  // once per build: every later level uses a fresh epoch tag instead of clearing
    Launch l(t, "table_clear", false);
    STB_CUDA(t, cudaMemsetAsync(sc.slots.ptr, 0xff, sc.slots.bytes(), st));
    sc.tags_cleared = true;
  }
  const uint32_t serial = ++sc.serial;
  const unsigned nb = (unsigned)ceil_div(n_next, BK_TILE);
  {
    Launch l(t, "bucket_insert");
    bucket_insert_kernel<<<nb, 256, 0, st>>>(rec_key.ptr, rec_meta.ptr, (uint32_t)n_next, bucket_off.ptr, log2p, sc.slots.ptr, serial, rec_slot.ptr);
  }
  {
    Launch l(t, "bucket_answer");
    bucket_answer_kernel<<<nb, 256, 0, st>>>(rec_key.ptr, rec_meta.ptr, (uint32_t)n_next, rec_slot.ptr, bucket_off.ptr, log2p, sc.slots.ptr,
                                             sc.bitmask.ptr, nxt);
  }
  STB_CUDA(t, cudaGetLastError());
  return STB_OK;
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
    const uint64_t n_next = ceil_div(n_cur, 2);
    t.layers.emplace_back();
    Layer& layer = t.layers.back();
    STB_CUDA(t, layer.nodes.alloc(n_next, st));
    LevelTable nt{sc.slots.ptr, nullptr, nullptr, table_cap(n_next)};
    nt.first_bits = sc.bitmask.ptr;
    bool via_position = false;
    static const uint64_t part_min = env_u64("STB_PART_MIN", ~0ull);  // experimental, off (profiles/README.md)
    if (n_next >= part_min) {
      STB_TRY(partitioned_insert_count(t, sc, cur, n_cur, nxt, n_next));
      via_position = true;
    } else {
    {
      Launch l(t, "table_clear", false);
      STB_CUDA(t, cudaMemsetAsync(nt.slots, 0xff, ((uint64_t)nt.cap + 1) * sizeof(Slot), st));
      STB_CUDA(t, cudaMemsetAsync(sc.bitmask.ptr, 0, ceil_div(n_next, LVL_TILE) * (LVL_TILE / 8), st));
    }
    {
      // Probe-then-claim measured faster than claim-first on B200 (12.0 vs 13.2 ms per 3.1 Gbp),
      // and chunking insert+count to keep table lines in L2 did not pay (profiles/README.md).
      Launch l(t, "node_insert");
      // children of level 0 are leaf ids (or an imported array): not position-ordered
      const uint32_t* child_unique = (locality && level > 0) ? counts_dev + level - 1 : nullptr;
      node_insert_kernel<true><<<(unsigned)ceil_div(n_next, LVL_THREADS), LVL_THREADS, 0, st>>>(cur, (uint32_t)n_cur, 0u, (uint32_t)n_next, nt, nxt, child_unique);
    }
    }
    finish_level<MODE_NODE>(t, nxt, (uint32_t)n_next, nt, sc, counts_dev + level, layer.nodes.ptr, cur, (uint32_t)n_cur, via_position);
    std::swap(cur, nxt);
    n_cur = n_next;
    ++level;
  } while (n_cur > 1);
  *levels_out = level;
  *root_buf = cur;
  return STB_OK;
}

int build_impl(Tree& t, const LeafInput& in, uint64_t n0, bool direct) {
  const int S = t.S;
  cudaStream_t st = t.stream;
  Scratch sc;
  const uint64_t n1 = ceil_div(n0, 2);
  STB_CUDA(t, sc.ptr_a.alloc(n0, st));
  STB_CUDA(t, sc.ptr_b.alloc(n1, st));
  STB_CUDA(t, sc.bitmask.alloc(ceil_div(n0, LVL_TILE) * (LVL_TILE / 32), st));
  STB_CUDA(t, sc.blockcnt.alloc(ceil_div(n0, LVL_TILE) + 1, st));
  STB_CUDA(t, sc.counts.alloc(80, st));
  STB_CUDA(t, sc.flags.alloc(1, st));
  STB_CUDA(t, sc.root.alloc(1, st));
  {
    BuildFlags init{~0ull, 0u, 0u};
    STB_CUDA(t, cudaMemcpyAsync(sc.flags.ptr, &init, sizeof(init), cudaMemcpyHostToDevice, st));
  }

  // ---- leaf level ----
  const uint64_t direct_entries = direct ? (1ull << (2 * S)) : 0;
  const uint32_t leaf_cap = direct ? 0u : table_cap(n0);
  const uint32_t node_cap_max = table_cap(n1);
  STB_CUDA(t, sc.slots.alloc((uint64_t)std::max(leaf_cap, node_cap_max) + 1, st));
  LevelTable tab{sc.slots.ptr, nullptr, nullptr, leaf_cap};
  tab.first_bits = sc.bitmask.ptr;
  STB_CUDA(t, cudaMemsetAsync(sc.bitmask.ptr, 0, sc.bitmask.bytes(), st));
  if (direct) {
    STB_CUDA(t, sc.dminpos.alloc(direct_entries, st));
    STB_CUDA(t, sc.dids.alloc(direct_entries, st));
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
  Scratch sc;
  const uint64_t n1 = ceil_div(n, 2);
  STB_CUDA(t, sc.ptr_a.alloc(n, st));
  STB_CUDA(t, sc.ptr_b.alloc(n1, st));
  STB_CUDA(t, sc.bitmask.alloc(ceil_div(n1, LVL_TILE) * (LVL_TILE / 32), st));
  STB_CUDA(t, sc.blockcnt.alloc(ceil_div(n1, LVL_TILE) + 1, st));
  STB_CUDA(t, sc.counts.alloc(80, st));
  STB_CUDA(t, sc.slots.alloc((uint64_t)table_cap(n1) + 1, st));
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

int build_from_leaves(Tree& t, const unsigned long long* d_leaves, uint64_t n) {
  LeafInput in;
  in.leaves = d_leaves;
  return build_dispatch(t, in, n);
}

}  // namespace stb
