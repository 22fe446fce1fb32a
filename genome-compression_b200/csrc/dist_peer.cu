// dist_peer.cu — the sharded level with the exchange fused into the kernels: records go
// straight into the owner's memory over NVLink (peer-mapped arenas), answers come back the same
// way.  No all-to-all, no count matrix on the host: per level the caller only runs one tiny
// barrier collective (records landed) and the bitmap all-reduce (which is also the second
// barrier).  Protocol in include/shared_tree_b200_dist.h ("peer exchange").
//
// Arena of one rank (identical layout on every rank, P = region capacity in records):
//   hdr    [world]      count: how many records source `src` sent me
//   cursor [world]      (private) how many records I have sent to owner o so far
//   ans    [world][P]   region o: answers from owner o to MY records, in the order I sent them
//   gpos   [world][P]   region `src`: global positions of the records source `src` sent me
//   keys   [world][P]   region `src`: their keys
// A region holds everything one source could ever send (all its positions), so no counts are
// needed before writing and nothing can overflow.  The order of the records inside a region is
// whatever the CTAs' reservations made it; every result is independent of it.
#include <algorithm>

#include "dist.cuh"

namespace stb {

struct PeerHdr {
  uint32_t count, pad0, pad1, pad2;
};

struct PeerBases {
  char* base[MAX_WORLD];
};

struct ArenaLayout {
  uint64_t hdr, cursor, ans, gpos, keys, bytes;
};

__host__ __device__ inline uint64_t up256(uint64_t x) { return (x + 255) & ~255ull; }

__host__ __device__ inline ArenaLayout arena_layout(int world, uint64_t P) {
  ArenaLayout L;
  L.hdr = 0;
  L.cursor = up256(MAX_WORLD * sizeof(PeerHdr));  // MAX_WORLD cursors + the CTA completion count
  L.ans = L.cursor + 256;
  L.gpos = L.ans + up256((uint64_t)world * P * 4);
  L.keys = L.gpos + up256((uint64_t)world * P * 4);
  L.bytes = L.keys + up256((uint64_t)world * P * 8);
  return L;
}

// ---- source side: canonicalise, split by owner in shared memory, write the runs to the owners ----
// One pass: a CTA reserves its run in every owner's region with one atomicAdd per owner on the
// rank's cursors; the last CTA to finish publishes the final counts to the owners.
#ifndef PS_ITEMS
#define PS_ITEMS 4
#endif
constexpr int PS_TILE = DP_THREADS * PS_ITEMS;  // records staged per CTA: runs of PS_TILE / world records per owner

template <int KIND>
__global__ void __launch_bounds__(DP_THREADS)
peer_scatter_kernel(const void* __restrict__ items, uint64_t n_items, uint64_t n_pos, int S, int world, int rank, uint64_t gpos0,
                    PeerBases peers, uint64_t P, ArenaLayout L, uint32_t* __restrict__ meta) {
  __shared__ uint32_t cnt[PS_ITEMS * DP_WARPS][MAX_WORLD];  // per (row, warp) counts -> offsets inside the owner's run
  __shared__ uint32_t cta_off[MAX_WORLD + 1];               // owner's run inside this CTA's staged tile
  __shared__ uint32_t seg_off[MAX_WORLD];                   // where that run goes inside my region at the owner
  __shared__ unsigned long long skey[PS_TILE];
  __shared__ uint32_t sgpos[PS_TILE], smeta[PS_TILE];
  __shared__ char* sbase[MAX_WORLD];
  __shared__ bool last_cta;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < MAX_WORLD; ++w)
    if (threadIdx.x == w) sbase[w] = peers.base[w];
  unsigned long long key[PS_ITEMS];
  uint32_t flg[PS_ITEMS], own[PS_ITEMS], rank_in_warp[PS_ITEMS];
#pragma unroll
  for (int it = 0; it < PS_ITEMS; ++it) {
    const uint64_t i = (uint64_t)blockIdx.x * PS_TILE + it * DP_THREADS + threadIdx.x;
    own[it] = 0xffffffffu;
    key[it] = 0;
    flg[it] = 0;
    rank_in_warp[it] = 0;
    if (i < n_pos) {
      produce<KIND>(items, n_items, S, i, key[it], flg[it]);
      own[it] = owner_of(key[it], world);
    }
    for (int w = 0; w < world; ++w) {
      const uint32_t m = __ballot_sync(0xffffffffu, own[it] == (uint32_t)w);
      if (own[it] == (uint32_t)w) rank_in_warp[it] = __popc(m & ((1u << lane) - 1u));
      if (lane == 0) cnt[it * DP_WARPS + warp][w] = __popc(m);
    }
  }
  __syncthreads();
  uint32_t* cursor = reinterpret_cast<uint32_t*>(sbase[rank] + L.cursor);
  if (threadIdx.x < world) {  // exclusive scan over the (row, warp) sequence, per owner; then reserve the run
    uint32_t run = 0;
    for (int j = 0; j < PS_ITEMS * DP_WARPS; ++j) {
      const uint32_t c = cnt[j][threadIdx.x];
      cnt[j][threadIdx.x] = run;
      run += c;
    }
    cta_off[threadIdx.x + 1] = run;  // run length for now
    seg_off[threadIdx.x] = run ? atomicAdd(cursor + threadIdx.x, run) : 0u;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    cta_off[0] = 0;
    for (int w = 0; w < world; ++w) cta_off[w + 1] += cta_off[w];
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < PS_ITEMS; ++it) {
    const uint64_t i = (uint64_t)blockIdx.x * PS_TILE + it * DP_THREADS + threadIdx.x;
    if (i < n_pos) {
      const uint32_t slot = cta_off[own[it]] + cnt[it * DP_WARPS + warp][own[it]] + rank_in_warp[it];
      skey[slot] = key[it];
      sgpos[slot] = (uint32_t)(gpos0 + i);
      smeta[slot] = (uint32_t)i | flg[it];
    }
  }
  __syncthreads();
  // the staged tile is now grouped by owner: consecutive threads write consecutive records of
  // one owner's run (full lines over NVLink instead of one record per owner per warp)
  const uint32_t total = cta_off[world];
  for (uint32_t t = threadIdx.x; t < total; t += DP_THREADS) {
    uint32_t o = 0;
    while (t >= cta_off[o + 1]) ++o;
    const uint32_t k = seg_off[o] + (t - cta_off[o]);
    char* base = sbase[o];
    reinterpret_cast<unsigned long long*>(base + L.keys)[(uint64_t)rank * P + k] = skey[t];
    reinterpret_cast<uint32_t*>(base + L.gpos)[(uint64_t)rank * P + k] = sgpos[t];
    meta[(uint64_t)o * P + k] = smeta[t];
  }
  // last CTA out publishes the counts (every reservation happened before its CTA's ticket)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last_cta = atomicAdd(cursor + MAX_WORLD, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last_cta && threadIdx.x < world) {
    __threadfence();
    PeerHdr h{__ldcg(cursor + threadIdx.x), 0u, 0u, 0u};
    reinterpret_cast<PeerHdr*>(sbase[threadIdx.x] + L.hdr)[rank] = h;
  }
}

// ---- owner side: the records lie in `world` regions; one virtual index space over them ----
struct OwnerView {
  uint32_t start[MAX_WORLD + 1];
};

__device__ __forceinline__ void load_view(OwnerView& v, const PeerHdr* __restrict__ hdr, int world) {
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int src = 0; src < world; ++src) {
      v.start[src] = run;
      run += __ldcg(&hdr[src].count);
    }
    v.start[world] = run;
  }
  __syncthreads();
}

__device__ __forceinline__ uint32_t owner_cap(uint32_t m) { return max(1024u, 2u * m); }

__device__ __forceinline__ void locate(const OwnerView& v, int world, uint32_t j, uint32_t& src, uint32_t& k) {
  src = 0;
  while ((int)src + 1 < world && j >= v.start[src + 1]) ++src;
  k = j - v.start[src];
}

__global__ void __launch_bounds__(256)
peer_owner_filter_kernel(const char* __restrict__ arena, ArenaLayout L, uint64_t P, int world, uint32_t* plane_a, uint32_t* plane_b,
                         uint32_t log2_bits) {
  __shared__ OwnerView v;
  load_view(v, reinterpret_cast<const PeerHdr*>(arena + L.hdr), world);
  const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(arena + L.keys);
  const uint32_t m = v.start[world];
  for (uint64_t tile = blockIdx.x; tile * 1024 < m; tile += gridDim.x) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const uint64_t j = tile * 1024 + it * 256 + threadIdx.x;
      if (j >= m) break;
      uint32_t src, k, word, bit;
      locate(v, world, (uint32_t)j, src, k);
      owner_filter_cell(__ldg(keys + (uint64_t)src * P + k), log2_bits, word, bit);
      if (atomicOr(plane_a + word, bit) & bit) atomicOr(plane_b + word, bit);
    }
  }
}

__global__ void __launch_bounds__(256)
peer_owner_insert_kernel(const char* __restrict__ arena, ArenaLayout L, uint64_t P, int world, Slot* tab, uint32_t serial,
                         uint32_t* __restrict__ slot_of, uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ plane_b,
                         uint32_t log2_bits) {
  __shared__ OwnerView v;
  load_view(v, reinterpret_cast<const PeerHdr*>(arena + L.hdr), world);
  const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(arena + L.keys);
  const uint32_t* gpos = reinterpret_cast<const uint32_t*>(arena + L.gpos);
  const uint32_t m = v.start[world], cap = owner_cap(m);
  for (uint64_t tile = blockIdx.x; tile * 1024 < m; tile += gridDim.x) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const uint64_t j = tile * 1024 + it * 256 + threadIdx.x;
      if (j >= m) break;
      uint32_t src, k;
      locate(v, world, (uint32_t)j, src, k);
      const unsigned long long key = __ldg(keys + (uint64_t)src * P + k);
      const uint32_t pos = __ldg(gpos + (uint64_t)src * P + k);
      if (plane_b) {
        uint32_t word, bit;
        owner_filter_cell(key, log2_bits, word, bit);
        if (!(__ldcg(plane_b + word) & bit)) {  // certified singleton: first occurrence, never in the table
          atomicOr(bitmap + (pos >> 5), 1u << (pos & 31));
          slot_of[j] = OWNER_SINGLETON;
          continue;
        }
      }
      slot_of[j] = tagged_insert(tab, cap, key, pos, serial, __umulhi(hash64(key), cap), 0u, 0xffffffffu, bitmap);
    }
  }
}

// answers only where they carry information: a later occurrence learns where its key came first
// (the source recognises its own first occurrences from the all-reduced bitmap)
__global__ void __launch_bounds__(256)
peer_owner_answer_kernel(const char* __restrict__ arena, ArenaLayout L, uint64_t P, int world, const Slot* __restrict__ tab,
                         const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ bitmap, PeerBases peers, int rank) {
  __shared__ OwnerView v;
  __shared__ char* sbase[MAX_WORLD];
#pragma unroll
  for (int w = 0; w < MAX_WORLD; ++w)
    if (threadIdx.x == w) sbase[w] = peers.base[w];
  load_view(v, reinterpret_cast<const PeerHdr*>(arena + L.hdr), world);
  const uint32_t* gpos = reinterpret_cast<const uint32_t*>(arena + L.gpos);
  const uint32_t m = v.start[world];
  for (uint64_t tile = blockIdx.x; tile * 1024 < m; tile += gridDim.x) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const uint64_t j = tile * 1024 + it * 256 + threadIdx.x;
      if (j >= m) break;
      uint32_t src, k;
      locate(v, world, (uint32_t)j, src, k);
      const uint32_t pos = __ldg(gpos + (uint64_t)src * P + k);
      if ((__ldcg(bitmap + (pos >> 5)) >> (pos & 31)) & 1u) continue;
      reinterpret_cast<uint32_t*>(sbase[src] + L.ans)[(uint64_t)rank * P + k] = __ldcg(&tab[slot_of[j]].minpos);
    }
  }
}

// Back at the source, send order (regions by owner, cursor[o] records each): every later
// occurrence gets the id of its key's first position; a first occurrence shows in the bitmap.
__global__ void __launch_bounds__(256)
peer_finish_rest_kernel(const char* __restrict__ arena, ArenaLayout L, uint64_t P, int world, uint64_t gpos0,
                        const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ word_prefix, uint64_t n_bits, uint64_t n_words,
                        const uint32_t* __restrict__ meta, uint32_t* __restrict__ pointers) {
  __shared__ OwnerView v;
  if (threadIdx.x == 0) {
    const uint32_t* cursor = reinterpret_cast<const uint32_t*>(arena + L.cursor);
    uint32_t run = 0;
    for (int o = 0; o < world; ++o) {
      v.start[o] = run;
      run += __ldcg(cursor + o);
    }
    v.start[world] = run;
  }
  __syncthreads();
  const uint32_t* ans = reinterpret_cast<const uint32_t*>(arena + L.ans);
  const uint32_t m = v.start[world];
  for (uint64_t tile = blockIdx.x; tile * 1024 < m; tile += gridDim.x) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const uint64_t j = tile * 1024 + it * 256 + threadIdx.x;
      if (j >= m) break;
      uint32_t o, k;
      locate(v, world, (uint32_t)j, o, k);
      const uint32_t mt = __ldg(meta + (uint64_t)o * P + k);
      const uint32_t pos = mt & IDX_MASK;
      const uint64_t g = gpos0 + pos;
      if ((__ldg(bitmap + (g >> 5)) >> (g & 31)) & 1u) continue;
      const uint32_t q = __ldcg(ans + (uint64_t)o * P + k);
      pointers[pos] = finish_pointer(rank_of(bitmap, word_prefix, n_bits, n_words, q), mt & ~IDX_MASK);
    }
  }
}

static int peer_bases(Ctx& t, int world, void* const* arenas, PeerBases& out) {
  for (int r = 0; r < MAX_WORLD; ++r) out.base[r] = nullptr;
  for (int r = 0; r < world; ++r) {
    if (!arenas[r]) return t.fail(STB_ERR_INVALID_ARG, "null arena pointer");
    out.base[r] = static_cast<char*>(arenas[r]);
  }
  return STB_OK;
}

}  // namespace stb

using namespace stb;

extern "C" {

uint64_t stb_dist_peer_arena_bytes(int world, uint64_t region_cap) {
  if (world < 1 || world > MAX_WORLD) return 0;
  return arena_layout(world, region_cap).bytes;
}

int stb_dist_peer_alloc(stb_tree* ctx, uint64_t bytes, void** ptr_out, unsigned char* handle_out) {
  if (!ctx || !ptr_out || !handle_out || bytes == 0) return STB_ERR_INVALID_ARG;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return STB_ERR_CUDA;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
  void* p = nullptr;
  STB_CUDA(*ctx, cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return ctx->fail_cuda(e, "cudaIpcGetMemHandle", __FILE__, __LINE__);
  }
  memcpy(handle_out, &h, sizeof h);
  *ptr_out = p;
  return STB_OK;
}

int stb_dist_peer_open(stb_tree* ctx, const unsigned char* handle, void** ptr_out) {
  if (!ctx || !handle || !ptr_out) return STB_ERR_INVALID_ARG;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return STB_ERR_CUDA;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  STB_CUDA(*ctx, cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return STB_OK;
}

int stb_dist_peer_close(stb_tree* ctx, void* ptr) {
  if (!ctx || !ptr) return STB_ERR_INVALID_ARG;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return STB_ERR_CUDA;
  STB_CUDA(*ctx, cudaIpcCloseMemHandle(ptr));
  return STB_OK;
}

int stb_dist_peer_free(stb_tree* ctx, void* ptr) {
  if (!ctx || !ptr) return STB_ERR_INVALID_ARG;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return STB_ERR_CUDA;
  STB_CUDA(*ctx, cudaFree(ptr));
  return STB_OK;
}

int stb_dist_peer_put(stb_tree* ctx, void* dst_dev, const void* src_dev, uint64_t bytes) {
  if (!ctx || (bytes && (!dst_dev || !src_dev))) return STB_ERR_INVALID_ARG;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return STB_ERR_CUDA;
  if (bytes) {
    Launch l(*ctx, "dist_peer_put", false);
    STB_CUDA(*ctx, cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDefault, ctx->stream));
  }
  return STB_OK;
}

int stb_dist_peer_scatter(stb_tree* ctx, int kind, const void* items_dev, uint64_t n_items, uint64_t gpos0, int world, int rank,
                          void* const* arenas, uint64_t region_cap, uint32_t* meta_dev) {
  if (!ctx || world < 1 || world > MAX_WORLD || rank < 0 || rank >= world || (kind != 0 && kind != 1) || !arenas)
    return STB_ERR_INVALID_ARG;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return STB_ERR_CUDA;
  Tree& t = *ctx;
  cudaStream_t st = t.stream;
  const uint64_t n_pos = kind == 0 ? n_items : ceil_div(n_items, 2);
  if (n_pos > region_cap) return t.fail(STB_ERR_INVALID_ARG, "more positions than the arena's region capacity");
  if (n_pos >= (1ull << 29)) return t.fail(STB_ERR_TOO_LARGE, "a rank holds 2^29 or more positions of one level");
  if (n_pos && (!items_dev || !meta_dev)) return STB_ERR_INVALID_ARG;
  PeerBases peers;
  STB_TRY(peer_bases(t, world, arenas, peers));
  const ArenaLayout L = arena_layout(world, region_cap);
  const uint32_t nblocks = (uint32_t)std::max<uint64_t>(1, ceil_div(n_pos, PS_TILE));
  STB_CUDA(t, cudaMemsetAsync(peers.base[rank] + L.cursor, 0, (MAX_WORLD + 1) * 4, st));
  {
    Launch l(t, "peer_scatter");
    if (kind == 0)
      peer_scatter_kernel<0><<<nblocks, DP_THREADS, 0, st>>>(items_dev, n_items, n_pos, t.S, world, rank, gpos0, peers, region_cap, L, meta_dev);
    else
      peer_scatter_kernel<1><<<nblocks, DP_THREADS, 0, st>>>(items_dev, n_items, n_pos, t.S, world, rank, gpos0, peers, region_cap, L, meta_dev);
  }
  STB_CUDA(t, cudaGetLastError());
  return STB_OK;
}

int stb_dist_peer_owner(stb_tree* ctx, int world, int rank, void* const* arenas, uint64_t region_cap, uint64_t expected_records,
                        void* table_dev, uint64_t table_slots, uint32_t serial, uint32_t* slot_scratch_dev, uint32_t* planes_dev,
                        uint64_t planes_words, uint32_t* bitmap_dev) {
  if (!ctx || world < 1 || world > MAX_WORLD || rank < 0 || rank >= world || !arenas || !table_dev || !slot_scratch_dev || !bitmap_dev ||
      serial == 0xffffffffu)
    return STB_ERR_INVALID_ARG;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return STB_ERR_CUDA;
  Tree& t = *ctx;
  cudaStream_t st = t.stream;
  const uint64_t worst = (uint64_t)world * region_cap;  // every record of the level lands here
  if (2 * worst >= 0xffffffffull) return t.fail(STB_ERR_TOO_LARGE, "level too large for one owner table");  // 32-bit slot numbers
  if (table_slots < std::max<uint64_t>(1024, 2 * worst) + 1) return t.fail(STB_ERR_BUFFER_TOO_SMALL, "owner table smaller than 2 * world * region_cap + 1 slots");
  PeerBases peers;
  STB_TRY(peer_bases(t, world, arenas, peers));
  const ArenaLayout L = arena_layout(world, region_cap);
  const char* mine = peers.base[rank];
  // grids: sized for the expected share, grid-stride for whatever actually arrived
  const uint64_t expect = std::max<uint64_t>(expected_records, 1024);
  const unsigned nb = (unsigned)std::min<uint64_t>(ceil_div(expect, 1024) + 64, 1u << 20);
  uint32_t log2_bits = 0;
  if (planes_dev && expected_records >= (1u << 16)) {
    log2_bits = 22;
    while (log2_bits < 28 && (1ull << log2_bits) < 2 * expected_records) ++log2_bits;
    while (log2_bits > 5 && 2 * ((1ull << log2_bits) / 32) > planes_words) --log2_bits;
    if (log2_bits < 16) log2_bits = 0;
  }
  uint32_t* plane_a = nullptr;
  uint32_t* plane_b = nullptr;
  if (log2_bits) {
    const uint64_t words = (1ull << log2_bits) / 32;
    plane_a = planes_dev;
    plane_b = planes_dev + words;
    STB_CUDA(t, cudaMemsetAsync(planes_dev, 0, 2 * words * 4, st));
    Launch l(t, "peer_owner_filter");
    peer_owner_filter_kernel<<<nb, 256, 0, st>>>(mine, L, region_cap, world, plane_a, plane_b, log2_bits);
  }
  {
    Launch l(t, "peer_owner_insert");
    peer_owner_insert_kernel<<<nb, 256, 0, st>>>(mine, L, region_cap, world, reinterpret_cast<Slot*>(table_dev), serial, slot_scratch_dev,
                                                 bitmap_dev, plane_b, log2_bits);
  }
  {
    Launch l(t, "peer_owner_answer");
    peer_owner_answer_kernel<<<nb, 256, 0, st>>>(mine, L, region_cap, world, reinterpret_cast<const Slot*>(table_dev), slot_scratch_dev, bitmap_dev,
                                                 peers, rank);
  }
  STB_CUDA(t, cudaGetLastError());
  return STB_OK;
}

int stb_dist_peer_finish(stb_tree* ctx, int kind, const void* items_dev, uint64_t n_items, uint64_t gpos0, const uint32_t* bitmap_dev,
                         const uint32_t* word_prefix_dev, uint64_t n_level_positions, const uint32_t* meta_dev, const void* arena,
                         int world, uint64_t region_cap, uint32_t* pointers_dev, void* layer_slice_dev, uint32_t* base_count_dev) {
  if (!ctx || (kind != 0 && kind != 1) || !bitmap_dev || !word_prefix_dev || !base_count_dev || !arena || world < 1 || world > MAX_WORLD)
    return STB_ERR_INVALID_ARG;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return STB_ERR_CUDA;
  Tree& t = *ctx;
  const uint64_t n_pos = kind == 0 ? n_items : ceil_div(n_items, 2);
  STB_TRY(finish_first(t, kind, items_dev, n_items, gpos0, bitmap_dev, word_prefix_dev, n_level_positions, pointers_dev, layer_slice_dev,
                       base_count_dev));
  if (n_pos) {
    Launch l(t, "peer_finish_rest");
    peer_finish_rest_kernel<<<(unsigned)ceil_div(n_pos, 1024), 256, 0, t.stream>>>(static_cast<const char*>(arena), arena_layout(world, region_cap),
                                                                                   region_cap, world, gpos0, bitmap_dev, word_prefix_dev,
                                                                                   n_level_positions, ceil_div(n_level_positions, 32), meta_dev,
                                                                                   pointers_dev);
  }
  STB_CUDA(t, cudaGetLastError());
  return STB_OK;
}

void* stb_dist_peer_payload(void* arena, int world, uint64_t region_cap) {
  if (!arena || world < 1 || world > MAX_WORLD) return nullptr;
  return static_cast<char*>(arena) + arena_layout(world, region_cap).keys;
}

}  // extern "C"
