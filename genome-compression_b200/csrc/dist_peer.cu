// dist_peer.cu — the sharded level with the exchange fused into the kernels: records go
// straight into the owner's memory over NVLink (peer-mapped arenas), answers come back the same
// way.  No all-to-all, no count matrix on the host: per level the caller only runs one tiny
// barrier collective (records landed) and the bitmap all-reduce (which is also the second
// barrier).  Protocol in include/shared_tree_b200_dist.h ("peer exchange").
//
// Arena of one rank (identical layout on every rank, P = region capacity in records):
//   hdr   [world]      {count, loc_off}: what source `src` sent me, and where that segment
//                      starts in the source's own send order (for the answers)
//   ans   [P]          answers to MY records, in my send order, written by their owners
//   gpos  [world][P]   region `src`: global positions of the records source `src` sent me
//   keys  [world][P]   region `src`: their keys
// A region holds everything one source could ever send (all its positions), so no counts are
// needed before writing and nothing can overflow.
#include <algorithm>

#include "dist.cuh"

namespace stb {

struct PeerHdr {
  uint32_t count, loc_off, pad0, pad1;
};

struct PeerBases {
  char* base[MAX_WORLD];
};

struct ArenaLayout {
  uint64_t hdr, ans, gpos, keys, bytes;
};

__host__ __device__ inline uint64_t up256(uint64_t x) { return (x + 255) & ~255ull; }

__host__ __device__ inline ArenaLayout arena_layout(int world, uint64_t P) {
  ArenaLayout L;
  L.hdr = 0;
  L.ans = up256(MAX_WORLD * sizeof(PeerHdr));
  L.gpos = L.ans + up256(P * 4);
  L.keys = L.gpos + up256((uint64_t)world * P * 4);
  L.bytes = L.keys + up256((uint64_t)world * P * 8);
  return L;
}

// ---- source side: canonicalise, split by owner in shared memory, write the runs to the owners ----
template <int KIND>
__global__ void __launch_bounds__(DP_THREADS)
peer_hist_kernel(const void* __restrict__ items, uint64_t n_items, uint64_t n_pos, int S, int world, uint32_t nblocks,
                 uint32_t* __restrict__ hist) {
  __shared__ uint32_t cnt[MAX_WORLD];
  if (threadIdx.x < MAX_WORLD) cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
#pragma unroll
  for (int it = 0; it < DP_ITEMS; ++it) {
    const uint64_t i = (uint64_t)blockIdx.x * DP_TILE + it * DP_THREADS + threadIdx.x;
    uint32_t o = 0xffffffffu;
    if (i < n_pos) {
      unsigned long long key;
      uint32_t f;
      produce<KIND>(items, n_items, S, i, key, f);
      o = owner_of(key, world);
    }
    for (int w = 0; w < world; ++w) {
      const uint32_t m = __ballot_sync(0xffffffffu, o == (uint32_t)w);
      if (lane == 0 && m) atomicAdd(&cnt[w], (uint32_t)__popc(m));
    }
  }
  __syncthreads();
  if (threadIdx.x < world) hist[threadIdx.x * nblocks + blockIdx.x] = cnt[threadIdx.x];
}

template <int KIND>
__global__ void __launch_bounds__(DP_THREADS)
peer_scatter_kernel(const void* __restrict__ items, uint64_t n_items, uint64_t n_pos, int S, int world, int rank, uint64_t gpos0,
                    uint32_t nblocks, const uint32_t* __restrict__ hist, const uint32_t* __restrict__ row_total, PeerBases peers,
                    uint64_t P, ArenaLayout L, uint32_t* __restrict__ meta) {
  __shared__ uint32_t cnt[DP_ITEMS * DP_WARPS][MAX_WORLD];  // per (row, warp) counts -> offsets inside the owner's run
  __shared__ uint32_t cta_off[MAX_WORLD + 1];               // owner's run inside this CTA's staged tile
  __shared__ uint32_t seg_off[MAX_WORLD];                   // where that run goes inside my region at the owner
  __shared__ uint32_t loc_off[MAX_WORLD];                   // where the owner's segment starts in my send order
  __shared__ unsigned long long skey[DP_TILE];
  __shared__ uint32_t sgpos[DP_TILE], smeta[DP_TILE];
  __shared__ char* sbase[MAX_WORLD];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < MAX_WORLD; ++w)
    if (threadIdx.x == w) sbase[w] = peers.base[w];
  if (threadIdx.x < MAX_WORLD) {
    uint32_t b = 0;
    for (int w = 0; w < (int)threadIdx.x && w < world; ++w) b += row_total[w];
    loc_off[threadIdx.x] = b;
    seg_off[threadIdx.x] = threadIdx.x < world ? hist[threadIdx.x * nblocks + blockIdx.x] : 0u;
  }
  unsigned long long key[DP_ITEMS];
  uint32_t flg[DP_ITEMS], own[DP_ITEMS], rank_in_warp[DP_ITEMS];
#pragma unroll
  for (int it = 0; it < DP_ITEMS; ++it) {
    const uint64_t i = (uint64_t)blockIdx.x * DP_TILE + it * DP_THREADS + threadIdx.x;
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
  if (threadIdx.x < world) {  // exclusive scan over the (row, warp) sequence, per owner
    uint32_t run = 0;
    for (int j = 0; j < DP_ITEMS * DP_WARPS; ++j) {
      const uint32_t c = cnt[j][threadIdx.x];
      cnt[j][threadIdx.x] = run;
      run += c;
    }
    cta_off[threadIdx.x + 1] = run;  // run length for now
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    cta_off[0] = 0;
    for (int w = 0; w < world; ++w) cta_off[w + 1] += cta_off[w];
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < DP_ITEMS; ++it) {
    const uint64_t i = (uint64_t)blockIdx.x * DP_TILE + it * DP_THREADS + threadIdx.x;
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
    meta[loc_off[o] + k] = smeta[t];
  }
  if (blockIdx.x == 0 && threadIdx.x < world) {
    PeerHdr h{row_total[threadIdx.x], loc_off[threadIdx.x], 0u, 0u};
    reinterpret_cast<PeerHdr*>(sbase[threadIdx.x] + L.hdr)[rank] = h;
  }
}

// ---- owner side: the records lie in `world` regions; one virtual index space over them ----
struct OwnerView {
  uint32_t start[MAX_WORLD + 1];
  uint32_t loc_off[MAX_WORLD];
};

__device__ __forceinline__ void load_view(OwnerView& v, const PeerHdr* __restrict__ hdr, int world) {
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int src = 0; src < world; ++src) {
      v.start[src] = run;
      run += __ldcg(&hdr[src].count);
      v.loc_off[src] = __ldcg(&hdr[src].loc_off);
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

__global__ void __launch_bounds__(256) peer_table_clear_kernel(const PeerHdr* __restrict__ hdr, int world, uint4* __restrict__ tab) {
  __shared__ OwnerView v;
  load_view(v, hdr, world);
  const uint64_t slots = (uint64_t)owner_cap(v.start[world]) + 1;
  const uint4 ones = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  for (uint64_t s = (uint64_t)blockIdx.x * 256 + threadIdx.x; s < slots; s += (uint64_t)gridDim.x * 256) tab[s] = ones;
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
peer_owner_insert_kernel(const char* __restrict__ arena, ArenaLayout L, uint64_t P, int world, Slot* tab, uint32_t* __restrict__ slot_of,
                         uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ plane_b, uint32_t log2_bits) {
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
      slot_of[j] = table_insert<true>(tab, cap, key, pos, bitmap);
    }
  }
}

// answers only where they carry information: a later occurrence learns where its key came first
// (the source recognises its own first occurrences from the all-reduced bitmap)
__global__ void __launch_bounds__(256)
peer_owner_answer_kernel(const char* __restrict__ arena, ArenaLayout L, uint64_t P, int world, const Slot* __restrict__ tab,
                         const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ bitmap, PeerBases peers) {
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
      reinterpret_cast<uint32_t*>(sbase[src] + L.ans)[v.loc_off[src] + k] = __ldcg(&tab[slot_of[j]].minpos);
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
  const uint32_t nblocks = (uint32_t)std::max<uint64_t>(1, ceil_div(n_pos, DP_TILE));
  DevBuf<uint32_t> hist;
  STB_CUDA(t, hist.alloc((uint64_t)world * nblocks + MAX_WORLD, st));
  uint32_t* row_total = hist.ptr + (uint64_t)world * nblocks;
  {
    Launch l(t, "peer_hist");
    if (kind == 0) peer_hist_kernel<0><<<nblocks, DP_THREADS, 0, st>>>(items_dev, n_items, n_pos, t.S, world, nblocks, hist.ptr);
    else peer_hist_kernel<1><<<nblocks, DP_THREADS, 0, st>>>(items_dev, n_items, n_pos, t.S, world, nblocks, hist.ptr);
  }
  {
    Launch l(t, "dist_rowscan");
    rowscan_kernel<<<world, 1024, 0, st>>>(hist.ptr, nblocks, row_total);
  }
  {
    Launch l(t, "peer_scatter");
    if (kind == 0)
      peer_scatter_kernel<0><<<nblocks, DP_THREADS, 0, st>>>(items_dev, n_items, n_pos, t.S, world, rank, gpos0, nblocks, hist.ptr, row_total,
                                                             peers, region_cap, L, meta_dev);
    else
      peer_scatter_kernel<1><<<nblocks, DP_THREADS, 0, st>>>(items_dev, n_items, n_pos, t.S, world, rank, gpos0, nblocks, hist.ptr, row_total,
                                                             peers, region_cap, L, meta_dev);
  }
  STB_CUDA(t, cudaGetLastError());
  return STB_OK;
}

int stb_dist_peer_owner(stb_tree* ctx, int world, int rank, void* const* arenas, uint64_t region_cap, uint64_t expected_records,
                        void* table_dev, uint64_t table_slots, uint32_t* slot_scratch_dev, uint32_t* planes_dev,
                        uint64_t planes_words, uint32_t* bitmap_dev) {
  if (!ctx || world < 1 || world > MAX_WORLD || rank < 0 || rank >= world || !arenas || !table_dev || !slot_scratch_dev || !bitmap_dev)
    return STB_ERR_INVALID_ARG;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return STB_ERR_CUDA;
  Tree& t = *ctx;
  cudaStream_t st = t.stream;
  const uint64_t worst = (uint64_t)world * region_cap;  // every record of the level lands here
  if (2 * worst > 0x1ffffffeull) return t.fail(STB_ERR_TOO_LARGE, "level too large for one owner table");
  if (table_slots < std::max<uint64_t>(1024, 2 * worst) + 1) return t.fail(STB_ERR_BUFFER_TOO_SMALL, "owner table smaller than 2 * world * region_cap + 1 slots");
  PeerBases peers;
  STB_TRY(peer_bases(t, world, arenas, peers));
  const ArenaLayout L = arena_layout(world, region_cap);
  const char* mine = peers.base[rank];
  const PeerHdr* hdr = reinterpret_cast<const PeerHdr*>(mine + L.hdr);
  // grids: sized for the expected share, grid-stride for whatever actually arrived
  const uint64_t expect = std::max<uint64_t>(expected_records, 1024);
  const unsigned nb = (unsigned)std::min<uint64_t>(ceil_div(expect, 1024) + 64, 1u << 20);
  {
    Launch l(t, "peer_table_clear");
    const unsigned cb = (unsigned)std::min<uint64_t>(ceil_div(2 * expect + 1, 256 * 8), 148 * 16);
    peer_table_clear_kernel<<<std::max(cb, 1u), 256, 0, st>>>(hdr, world, reinterpret_cast<uint4*>(table_dev));
  }
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
    peer_owner_insert_kernel<<<nb, 256, 0, st>>>(mine, L, region_cap, world, reinterpret_cast<Slot*>(table_dev), slot_scratch_dev, bitmap_dev,
                                                 plane_b, log2_bits);
  }
  {
    Launch l(t, "peer_owner_answer");
    peer_owner_answer_kernel<<<nb, 256, 0, st>>>(mine, L, region_cap, world, reinterpret_cast<const Slot*>(table_dev), slot_scratch_dev, bitmap_dev,
                                                 peers);
  }
  STB_CUDA(t, cudaGetLastError());
  return STB_OK;
}

const void* stb_dist_peer_answers(void* arena, int world, uint64_t region_cap) {
  if (!arena || world < 1 || world > MAX_WORLD) return nullptr;
  return static_cast<const char*>(arena) + arena_layout(world, region_cap).ans;
}

void* stb_dist_peer_payload(void* arena, int world, uint64_t region_cap) {
  if (!arena || world < 1 || world > MAX_WORLD) return nullptr;
  return static_cast<char*>(arena) + arena_layout(world, region_cap).keys;
}

}  // extern "C"
