// ingest.cu — FASTA text -> bare body -> packed leaves, on the device.
//
// Reproduces the observable semantics of fasta_reader::load_buffer (reference
// src/fasta_reader.cpp:40-68), which is a sequential getline loop:
//   * per round: if the next byte is '>' or '\n', ONE line is discarded, then ONE line
//     is read as data whatever it starts with;
//   * newlines are never data; lines are concatenated across records;
//   * the trailing (body_len mod dna_size) bases are dropped (at packing).
// Parallel restatement: a line that starts with '>' or is blank ("hb" line) is skipped
// iff the number of hb lines immediately before it is even; every other line is data.
// That is a segmented count over line starts, then a segmented broadcast of the
// line's skip flag over its bytes, then a compaction — three streaming passes with
// two tiny single-CTA scans in between.
#include "pack.cuh"
#include "tree.h"

namespace stb {

std::string unknown_symbol_message(int upper_byte) {
  // reference src/dna.cpp:45 prints the (int-typed) upper-cased code twice
  return "Encountered unknown symbol: " + std::to_string(upper_byte) + " (ASCII code " + std::to_string(upper_byte) + ")";
}

// ---- body -> packed leaves (read_genome's result, src/fasta_reader.cpp:108) ---------
template <int S_T>
__global__ void __launch_bounds__(PACK_THREADS)
pack_body_kernel(const char* __restrict__ body, uint64_t n_leaves, int S_rt, unsigned long long* __restrict__ out,
                 unsigned long long* bad_symbol) {
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
      out[tile_first + j] = pack_leaf<S_T>(lut, tile, j, S, bad);
      if (bad != 0xFFFFFFFFu) {
        uint32_t c = tile[bad];
        if (c >= 'a' && c <= 'z') c -= 32;
        atomicMin(bad_symbol, ((tile_first * S + bad) << 8) | c);
      }
    }
  }
}

int pack_body(Ctx& ctx, const char* d_body, uint64_t n_leaves, unsigned long long* d_leaves) {
  if (n_leaves == 0) return STB_OK;
  DevBuf<unsigned long long> bad;
  STB_CUDA(ctx, bad.alloc(1, ctx.stream));
  STB_CUDA(ctx, cudaMemsetAsync(bad.ptr, 0xff, 8, ctx.stream));
  {
    Launch l(ctx, "pack_body");
    const unsigned nb = (unsigned)ceil_div(n_leaves, PACK_TILE_LEAVES);
    const size_t smem = pack_smem_bytes(ctx.S);
    if (ctx.S == 12) pack_body_kernel<12><<<nb, PACK_THREADS, smem, ctx.stream>>>(d_body, n_leaves, ctx.S, d_leaves, bad.ptr);
    else pack_body_kernel<0><<<nb, PACK_THREADS, smem, ctx.stream>>>(d_body, n_leaves, ctx.S, d_leaves, bad.ptr);
  }
  unsigned long long h = 0;
  STB_CUDA(ctx, cudaMemcpyAsync(&h, bad.ptr, 8, cudaMemcpyDeviceToHost, ctx.stream));
  STB_CUDA(ctx, cudaStreamSynchronize(ctx.stream));
  STB_CUDA(ctx, cudaGetLastError());
  if (h != ~0ull) return ctx.fail(STB_ERR_UNKNOWN_SYMBOL, unknown_symbol_message((int)(h & 0xff)));
  return STB_OK;
}

// ---- FASTA text -> body ------------------------------------------------------------
constexpr int FX_THREADS = 256;
constexpr int FX_CHUNK = 16;                       // bytes per thread per sub-tile (one uint4)
constexpr int FX_SUB = FX_THREADS * FX_CHUNK;      // 4096 bytes per sub-tile
constexpr int FX_SUBS_PER_TILE = 16;
constexpr int FX_TILE = FX_SUB * FX_SUBS_PER_TILE; // 64 KiB per CTA

// run summary over line starts: composition  a ; b
struct RunSum {
  uint32_t all_hb;  // every line start in the span is an hb line (vacuously 1 when none)
  uint32_t trail;   // number of hb line starts at the end of the span
};
__device__ __forceinline__ RunSum run_compose(RunSum a, RunSum b) {
  RunSum r;
  r.all_hb = a.all_hb & b.all_hb;
  r.trail = b.all_hb ? a.trail + b.trail : b.trail;
  return r;
}
// skip-state summary: does the span contain a line start, and is its last line skipped
struct SkipSum {
  uint32_t has_ls;
  uint32_t last_skip;
};
__device__ __forceinline__ SkipSum skip_compose(SkipSum a, SkipSum b) { return b.has_ls ? b : a; }

// per-tile summary for the second scan: kept-byte counts for either incoming skip state
struct TileSum {
  uint32_t has_ls, last_skip;
  uint32_t cnt[2];  // kept bytes if the tile is entered with skip state 0 / 1
};

struct FxCarry {     // what a tile needs to start
  RunSum run;        // hb-run state before the tile's first line start
  uint32_t in_skip;  // is the line that straddles the tile start a skipped one
  uint64_t out_off;  // body offset of the tile's first kept byte
};

// Block-wide exclusive scan with a generic operator (256 threads), identity supplied.
template <typename T, typename Op>
__device__ __forceinline__ T block_exclusive(T v, T identity, Op op, T* smem_warp, T& total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T y;
    // shuffle the struct word by word
    uint32_t* xs = reinterpret_cast<uint32_t*>(&x);
    uint32_t* ys = reinterpret_cast<uint32_t*>(&y);
#pragma unroll
    for (int w = 0; w < (int)(sizeof(T) / 4); ++w) ys[w] = __shfl_up_sync(0xffffffffu, xs[w], d);
    if (lane >= d) x = op(y, x);
  }
  // x = inclusive within warp
  if (lane == 31) smem_warp[warp] = x;
  __syncthreads();
  T prefix = identity;
  for (uint32_t w = 0; w < warp; ++w) prefix = op(prefix, smem_warp[w]);
  T tot = identity;
  for (uint32_t w = 0; w < FX_THREADS / 32; ++w) tot = op(tot, smem_warp[w]);
  total = tot;
  // exclusive = prefix ; (inclusive of previous lane)
  T prev;
  {
    uint32_t* xs = reinterpret_cast<uint32_t*>(&x);
    uint32_t* ps = reinterpret_cast<uint32_t*>(&prev);
#pragma unroll
    for (int w = 0; w < (int)(sizeof(T) / 4); ++w) ps[w] = __shfl_up_sync(0xffffffffu, xs[w], 1);
  }
  T excl = lane == 0 ? prefix : op(prefix, prev);
  __syncthreads();
  return excl;
}

struct U32Sum {
  uint32_t v;
};

// PHASE 0: tile run summary.  PHASE 1: tile skip/count summary.  PHASE 2: emit body.
template <int PHASE>
__global__ void __launch_bounds__(FX_THREADS)
fasta_tile_kernel(const char* __restrict__ text, uint64_t len, RunSum* __restrict__ run_sums,
                  const RunSum* __restrict__ run_carry, TileSum* __restrict__ tile_sums,
                  const FxCarry* __restrict__ carry, char* __restrict__ body, const char* text_begin) {
  __shared__ RunSum sm_run[FX_THREADS / 32];
  __shared__ SkipSum sm_skip[FX_THREADS / 32];
  __shared__ U32Sum sm_cnt[FX_THREADS / 32];
  __shared__ uint8_t sm_last[FX_THREADS / 32];
  __shared__ __align__(16) uint8_t sm_out[FX_SUB + 16];

  const uint64_t tile_base = (uint64_t)blockIdx.x * FX_TILE;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  RunSum run_state = PHASE == 0 ? RunSum{1u, 0u} : (PHASE == 1 ? run_carry[blockIdx.x] : carry[blockIdx.x].run);
  // PHASE 1 tracks both possible incoming skip states; PHASE 2 knows the real one.
  uint32_t in_skip0 = 0, in_skip1 = 1;
  if (PHASE == 2) in_skip0 = in_skip1 = carry[blockIdx.x].in_skip;
  uint32_t tile_has_ls = 0, tile_last_skip = 0, tile_cnt0 = 0, tile_cnt1 = 0;
  uint64_t out_off = PHASE == 2 ? carry[blockIdx.x].out_off : 0;
  // `text` may be a later chunk of a text that begins at text_begin: the byte before it is then real
  uint8_t prev_sub_last = (text + tile_base == text_begin) ? (uint8_t)'\n' : (uint8_t)text[(int64_t)tile_base - 1];

  for (int sub = 0; sub < FX_SUBS_PER_TILE; ++sub) {
    const uint64_t sub_base = tile_base + (uint64_t)sub * FX_SUB;
    if (sub_base >= len) break;
    const uint64_t my_base = sub_base + (uint64_t)threadIdx.x * FX_CHUNK;
    // ---- load 16 bytes ----
    uint8_t c[FX_CHUNK];
    int valid = 0;
    if (my_base + FX_CHUNK <= len) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + my_base));
      *reinterpret_cast<uint4*>(c) = v;
      valid = FX_CHUNK;
    } else if (my_base < len) {
      valid = (int)(len - my_base);
      for (int i = 0; i < FX_CHUNK; ++i) c[i] = i < valid ? (uint8_t)text[my_base + i] : (uint8_t)0;
    }
    // ---- previous byte ----
    uint8_t my_last = valid ? c[valid - 1] : (uint8_t)0;
    // a thread with no valid bytes forwards nothing meaningful; threads after the end
    // of text never have line starts, so their prev byte is irrelevant.
    uint32_t pb = __shfl_up_sync(0xffffffffu, (uint32_t)my_last, 1);
    if (lane == 31) sm_last[warp] = my_last;
    __syncthreads();
    if (lane == 0) pb = warp == 0 ? prev_sub_last : sm_last[warp - 1];
    const uint8_t sub_last = sm_last[FX_THREADS / 32 - 1];

    // ---- line starts in my chunk ----
    uint32_t ls_mask = 0, hb_mask = 0;
#pragma unroll
    for (int i = 0; i < FX_CHUNK; ++i) {
      const uint8_t before = i == 0 ? (uint8_t)pb : c[i - 1];
      if (i < valid && before == '\n') {
        ls_mask |= 1u << i;
        if (c[i] == '>' || c[i] == '\n') hb_mask |= 1u << i;
      }
    }
    // ---- run summary of my chunk ----
    RunSum mine{1u, 0u};
    {
      uint32_t m = ls_mask;
      while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        if ((hb_mask >> i) & 1u) mine.trail += 1;
        else { mine.all_hb = 0; mine.trail = 0; }
      }
    }
    RunSum run_total;
    RunSum run_before = block_exclusive(mine, RunSum{1u, 0u}, run_compose, sm_run, run_total);
    run_before = run_compose(run_state, run_before);
    run_state = run_compose(run_state, run_total);
    if (PHASE == 0) {
      prev_sub_last = sub_last;
      __syncthreads();
      continue;
    }
    // ---- skip decision per line start ----
    uint32_t skip_mask = 0;
    {
      uint32_t run = run_before.trail;
      uint32_t m = ls_mask;
      while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        if ((hb_mask >> i) & 1u) {
          if ((run & 1u) == 0) skip_mask |= 1u << i;
          run += 1;
        } else {
          run = 0;
        }
      }
    }
    SkipSum sk{ls_mask != 0, 0u};
    if (ls_mask) sk.last_skip = (skip_mask >> (31 - __clz(ls_mask))) & 1u;
    SkipSum sk_total;
    const SkipSum sk_before = block_exclusive(sk, SkipSum{0u, 0u}, skip_compose, sm_skip, sk_total);
    // ---- kept bytes for incoming state x ----
    auto kept_mask_for = [&](uint32_t in_skip) {
      uint32_t cur = sk_before.has_ls ? sk_before.last_skip : in_skip;
      uint32_t keep = 0;
#pragma unroll
      for (int i = 0; i < FX_CHUNK; ++i) {
        if ((ls_mask >> i) & 1u) cur = (skip_mask >> i) & 1u;
        if (i < valid && c[i] != '\n' && !cur) keep |= 1u << i;
      }
      return keep;
    };
    const uint32_t keep0 = kept_mask_for(in_skip0);
    if (PHASE == 1) {
      const uint32_t keep1 = kept_mask_for(in_skip1);
      U32Sum t0, t1;
      auto add = [](U32Sum a, U32Sum b) { return U32Sum{a.v + b.v}; };
      block_exclusive(U32Sum{(uint32_t)__popc(keep0)}, U32Sum{0u}, add, sm_cnt, t0);
      block_exclusive(U32Sum{(uint32_t)__popc(keep1)}, U32Sum{0u}, add, sm_cnt, t1);
      tile_cnt0 += t0.v;
      tile_cnt1 += t1.v;
      if (sk_total.has_ls) {
        tile_has_ls = 1;
        tile_last_skip = sk_total.last_skip;
        in_skip0 = in_skip1 = sk_total.last_skip;
      }
    } else {
      U32Sum tot;
      auto add = [](U32Sum a, U32Sum b) { return U32Sum{a.v + b.v}; };
      const uint32_t my_off = block_exclusive(U32Sum{(uint32_t)__popc(keep0)}, U32Sum{0u}, add, sm_cnt, tot).v;
      // stage kept bytes so that sm_out[shift + k] <-> body[out_off + k]
      const uint32_t shift = (uint32_t)(out_off & 3u);
      uint32_t o = shift + my_off;
      uint32_t m = keep0;
      while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        sm_out[o++] = c[i];
      }
      __syncthreads();
      const uint32_t total = tot.v;
      if (total) {
        char* dst = body + (out_off - shift);  // 4-byte aligned
        const uint32_t end = shift + total;
        const uint32_t first_word = shift ? 1u : 0u;
        const uint32_t last_word = end >> 2;  // words [first_word, last_word) are complete
        for (uint32_t w = first_word + threadIdx.x; w < last_word; w += FX_THREADS)
          reinterpret_cast<uint32_t*>(dst)[w] = reinterpret_cast<const uint32_t*>(sm_out)[w];
        if (threadIdx.x == 0) {
          if (shift) {
            const uint32_t stop = end < 4u ? end : 4u;
            for (uint32_t k = shift; k < stop; ++k) dst[k] = (char)sm_out[k];
          }
          if (last_word >= first_word)
            for (uint32_t k = max(last_word << 2, shift); k < end; ++k) dst[k] = (char)sm_out[k];
        }
      }
      out_off += total;
      if (sk_total.has_ls) in_skip0 = in_skip1 = sk_total.last_skip;
    }
    prev_sub_last = sub_last;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (PHASE == 0) run_sums[blockIdx.x] = run_state;
    if (PHASE == 1) tile_sums[blockIdx.x] = TileSum{tile_has_ls, tile_last_skip, {tile_cnt0, tile_cnt1}};
  }
}

// Single-thread-per-element would be enough (<= ~50K tiles), but keep it parallel:
// exclusive scans over tiles with the two operators, done by one CTA sequentially in
// chunks of 256.
// What the automaton carries from one chunk of a text to the next (streaming ingest).
struct FxStream {
  RunSum run;        // hb-run state at the end of the text seen so far
  uint32_t in_skip;  // is the line the text ends in a skipped one
  uint32_t pad;
  unsigned long long out_off;  // body bytes produced so far
};

__global__ void __launch_bounds__(FX_THREADS) fasta_scan_run_kernel(const RunSum* __restrict__ sums, uint32_t n,
                                                                    RunSum* __restrict__ carry, FxStream* __restrict__ stream) {
  __shared__ RunSum sm[FX_THREADS / 32];
  RunSum state = stream ? stream->run : RunSum{1u, 0u};
  for (uint32_t base = 0; base < n; base += FX_THREADS) {
    const uint32_t i = base + threadIdx.x;
    const RunSum v = i < n ? sums[i] : RunSum{1u, 0u};
    RunSum total;
    const RunSum before = block_exclusive(v, RunSum{1u, 0u}, run_compose, sm, total);
    if (i < n) carry[i] = run_compose(state, before);
    state = run_compose(state, total);
  }
  if (stream && threadIdx.x == 0) stream->run = state;  // the skip scan below reads run_carry[], not this
}

struct ScanB {
  uint32_t has_ls, last_skip;
  uint32_t cnt0_lo, cnt0_hi, cnt1_lo, cnt1_hi;  // 64-bit counts split for word shuffles
};
__device__ __forceinline__ uint64_t u64_of(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
__device__ __forceinline__ ScanB scanb_compose(ScanB a, ScanB b) {
  // entering a with state x leaves it in state (a.has_ls ? a.last_skip : x)
  const uint64_t b0 = u64_of(b.cnt0_lo, b.cnt0_hi), b1 = u64_of(b.cnt1_lo, b.cnt1_hi);
  const uint32_t out0 = a.has_ls ? a.last_skip : 0u, out1 = a.has_ls ? a.last_skip : 1u;
  const uint64_t c0 = u64_of(a.cnt0_lo, a.cnt0_hi) + (out0 ? b1 : b0);
  const uint64_t c1 = u64_of(a.cnt1_lo, a.cnt1_hi) + (out1 ? b1 : b0);
  ScanB r;
  r.has_ls = a.has_ls | b.has_ls;
  r.last_skip = b.has_ls ? b.last_skip : a.last_skip;
  r.cnt0_lo = (uint32_t)c0; r.cnt0_hi = (uint32_t)(c0 >> 32);
  r.cnt1_lo = (uint32_t)c1; r.cnt1_hi = (uint32_t)(c1 >> 32);
  return r;
}

__global__ void __launch_bounds__(FX_THREADS)
fasta_scan_skip_kernel(const TileSum* __restrict__ sums, const RunSum* __restrict__ run_carry, uint32_t n,
                       FxCarry* __restrict__ carry, unsigned long long* __restrict__ body_len, FxStream* __restrict__ stream) {
  __shared__ ScanB sm[FX_THREADS / 32];
  const ScanB ident{0u, 0u, 0u, 0u, 0u, 0u};
  ScanB state = ident;  // the file is entered in skip state 0
  if (stream) {  // a later chunk is entered in the state, and at the body offset, the text so far ended with
    const unsigned long long o = stream->out_off;
    state = ScanB{1u, stream->in_skip, (uint32_t)o, (uint32_t)(o >> 32), (uint32_t)o, (uint32_t)(o >> 32)};
  }
  for (uint32_t base = 0; base < n; base += FX_THREADS) {
    const uint32_t i = base + threadIdx.x;
    ScanB v = ident;
    if (i < n) {
      const TileSum t = sums[i];
      v = ScanB{t.has_ls, t.last_skip, t.cnt[0], 0u, t.cnt[1], 0u};
    }
    ScanB total;
    const ScanB before = block_exclusive(v, ident, scanb_compose, sm, total);
    const ScanB upto = scanb_compose(state, before);
    if (i < n) {
      FxCarry c;
      c.run = run_carry[i];
      c.in_skip = upto.has_ls ? upto.last_skip : 0u;
      c.out_off = u64_of(upto.cnt0_lo, upto.cnt0_hi);
      carry[i] = c;
    }
    state = scanb_compose(state, total);
  }
  if (threadIdx.x == 0) {
    *body_len = u64_of(state.cnt0_lo, state.cnt0_hi);
    if (stream) {
      stream->in_skip = state.has_ls ? state.last_skip : 0u;
      stream->out_off = u64_of(state.cnt0_lo, state.cnt0_hi);
    }
  }
}

int fasta_extract_body(Ctx& ctx, const char* d_text, uint64_t len, DevBuf<char>& body, uint64_t* body_len) {
  *body_len = 0;
  STB_CUDA(ctx, body.alloc(len + 16, ctx.stream));
  if (len == 0) return STB_OK;
  const uint32_t tiles = (uint32_t)ceil_div(len, FX_TILE);
  DevBuf<RunSum> run_sums, run_carry;
  DevBuf<TileSum> tile_sums;
  DevBuf<FxCarry> carry;
  DevBuf<unsigned long long> d_len;
  STB_CUDA(ctx, run_sums.alloc(tiles, ctx.stream));
  STB_CUDA(ctx, run_carry.alloc(tiles, ctx.stream));
  STB_CUDA(ctx, tile_sums.alloc(tiles, ctx.stream));
  STB_CUDA(ctx, carry.alloc(tiles, ctx.stream));
  STB_CUDA(ctx, d_len.alloc(1, ctx.stream));
  {
    Launch l(ctx, "fasta_lines");
    fasta_tile_kernel<0><<<tiles, FX_THREADS, 0, ctx.stream>>>(d_text, len, run_sums.ptr, nullptr, nullptr, nullptr, nullptr, d_text);
  }
  {
    Launch l(ctx, "fasta_scan");
    fasta_scan_run_kernel<<<1, FX_THREADS, 0, ctx.stream>>>(run_sums.ptr, tiles, run_carry.ptr, nullptr);
  }
  {
    Launch l(ctx, "fasta_count");
    fasta_tile_kernel<1><<<tiles, FX_THREADS, 0, ctx.stream>>>(d_text, len, nullptr, run_carry.ptr, tile_sums.ptr, nullptr, nullptr, d_text);
  }
  {
    Launch l(ctx, "fasta_scan");
    fasta_scan_skip_kernel<<<1, FX_THREADS, 0, ctx.stream>>>(tile_sums.ptr, run_carry.ptr, tiles, carry.ptr, d_len.ptr, nullptr);
  }
  {
    Launch l(ctx, "fasta_emit");
    fasta_tile_kernel<2><<<tiles, FX_THREADS, 0, ctx.stream>>>(d_text, len, nullptr, nullptr, nullptr, carry.ptr, body.ptr, d_text);
  }
  unsigned long long h = 0;
  STB_CUDA(ctx, cudaMemcpyAsync(&h, d_len.ptr, 8, cudaMemcpyDeviceToHost, ctx.stream));
  STB_CUDA(ctx, cudaStreamSynchronize(ctx.stream));
  STB_CUDA(ctx, cudaGetLastError());
  *body_len = h;
  return STB_OK;
}

// ---- the same, chunk by chunk (streaming ingest from host memory) --------------------------------
// The text arrives in chunks of a multiple of FX_TILE bytes in ONE contiguous device buffer; the
// automaton's state at the end of a chunk (FxStream, on the device) is where the next chunk starts.
struct FastaStreamImpl {
  DevBuf<RunSum> run_sums, run_carry;
  DevBuf<TileSum> tile_sums;
  DevBuf<FxCarry> carry;
  DevBuf<FxStream> state;
  DevBuf<unsigned long long> d_len;
};

int fasta_stream_begin(Ctx& ctx, FastaStream& fs, uint64_t max_chunk_bytes) {
  if (!fs.impl) fs.impl = std::make_shared<FastaStreamImpl>();
  auto& im = *static_cast<FastaStreamImpl*>(fs.impl.get());
  const uint64_t tiles = ceil_div(max_chunk_bytes, FX_TILE) + 1;
  STB_CUDA(ctx, im.run_sums.ensure(tiles, ctx.stream));
  STB_CUDA(ctx, im.run_carry.ensure(tiles, ctx.stream));
  STB_CUDA(ctx, im.tile_sums.ensure(tiles, ctx.stream));
  STB_CUDA(ctx, im.carry.ensure(tiles, ctx.stream));
  STB_CUDA(ctx, im.state.ensure(1, ctx.stream));
  STB_CUDA(ctx, im.d_len.ensure(1, ctx.stream));
  const FxStream init{RunSum{1u, 0u}, 0u, 0u, 0ull};
  STB_CUDA(ctx, cudaMemcpyAsync(im.state.ptr, &init, sizeof(init), cudaMemcpyHostToDevice, ctx.stream));
  STB_CUDA(ctx, cudaStreamSynchronize(ctx.stream));  // `init` is stack memory
  return STB_OK;
}

int fasta_stream_chunk(Ctx& ctx, FastaStream& fs, const char* d_text_begin, uint64_t first, uint64_t len, char* d_body) {
  auto& im = *static_cast<FastaStreamImpl*>(fs.impl.get());
  if (len == 0) return STB_OK;
  if (first % FX_TILE) return ctx.fail(STB_ERR_INVALID_ARG, "fasta stream: chunk start must be a multiple of the extraction tile");
  const char* text = d_text_begin + first;
  const uint32_t tiles = (uint32_t)ceil_div(len, FX_TILE);
  cudaStream_t st = ctx.stream;
  {
    Launch l(ctx, "fasta_lines");
    fasta_tile_kernel<0><<<tiles, FX_THREADS, 0, st>>>(text, len, im.run_sums.ptr, nullptr, nullptr, nullptr, nullptr, d_text_begin);
  }
  {
    Launch l(ctx, "fasta_scan");
    fasta_scan_run_kernel<<<1, FX_THREADS, 0, st>>>(im.run_sums.ptr, tiles, im.run_carry.ptr, im.state.ptr);
  }
  {
    Launch l(ctx, "fasta_count");
    fasta_tile_kernel<1><<<tiles, FX_THREADS, 0, st>>>(text, len, nullptr, im.run_carry.ptr, im.tile_sums.ptr, nullptr, nullptr, d_text_begin);
  }
  {
    Launch l(ctx, "fasta_scan");
    fasta_scan_skip_kernel<<<1, FX_THREADS, 0, st>>>(im.tile_sums.ptr, im.run_carry.ptr, tiles, im.carry.ptr, im.d_len.ptr, im.state.ptr);
  }
  {
    Launch l(ctx, "fasta_emit");
    fasta_tile_kernel<2><<<tiles, FX_THREADS, 0, st>>>(text, len, nullptr, nullptr, nullptr, im.carry.ptr, d_body, d_text_begin);
  }
  return STB_OK;
}

int fasta_stream_body_len(Ctx& ctx, FastaStream& fs, uint64_t* body_len) {
  auto& im = *static_cast<FastaStreamImpl*>(fs.impl.get());
  unsigned long long h = 0;
  STB_CUDA(ctx, cudaMemcpyAsync(&h, im.d_len.ptr, 8, cudaMemcpyDeviceToHost, ctx.stream));
  STB_CUDA(ctx, cudaStreamSynchronize(ctx.stream));
  STB_CUDA(ctx, cudaGetLastError());
  *body_len = h;
  return STB_OK;
}

}  // namespace stb
