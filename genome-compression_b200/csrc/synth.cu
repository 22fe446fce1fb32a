// synth.cu — the synthetic genome-shaped workload of BASELINE.json configs 3-5,
// generated on the device.  Not part of the reference; the oracle has
// an independent CPU definition (oracle/oracle.c: orc_synth_*) that tests compare
// against byte for byte.
//
// sequence[i] = ACGT[ 2 bits of splitmix64(seed, i/32) ]  (i.i.d. uniform), then
// planted repeats: disjoint destination intervals, each a copy (or reverse-complement
// copy) of a stretch of the UNDERLYING i.i.d. sequence, so every output base is a
// pure function of its position.
#include <algorithm>
#include <vector>

#include "tree.h"

namespace stb {

struct Repeat {
  unsigned long long dst, src, len;
  uint32_t rc, pad;
};

__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__host__ __device__ __forceinline__ uint32_t base_code(unsigned long long seed, unsigned long long i) {
  const unsigned long long w = splitmix64(seed ^ ((i >> 5) * 0xD1B54A32D192ED03ull));
  return (uint32_t)((w >> (2 * (i & 31))) & 3ull);
}

constexpr unsigned long long SYNTH_ALIGN = 12ull * 1024ull;

// Destination intervals in increasing order; lengths 300*2^k + r (k in 0..9), capped at
// 200000; gaps uniform with mean len*(1-f)/f; of every 8 copies 4 are forward copies at a
// distance that is a multiple of 12*1024 bases, 1 is a reverse-complement copy aligned the
// same way, 3 are at an arbitrary distance.
static std::vector<Repeat> make_repeats(uint64_t n_bases, uint64_t seed, uint32_t permille) {
  std::vector<Repeat> out;
  if (permille == 0 || permille >= 1000) return out;
  unsigned long long state = splitmix64(seed ^ 0x5eedc0de5eedc0deull);
  unsigned long long cursor = 0;
  for (;;) {
    state = splitmix64(state);
    const unsigned k = (unsigned)(state % 10);
    state = splitmix64(state);
    unsigned long long len = (300ull << k) + state % (300ull << k);
    if (len > 200000) len = 200000;
    state = splitmix64(state);
    const unsigned long long span = 2 * len * (1000 - permille) / permille + 1;
    const unsigned long long dst = cursor + state % span;
    if (dst + len > n_bases) break;
    state = splitmix64(state);
    unsigned long long src = state % (n_bases - len + 1);
    state = splitmix64(state);
    const unsigned cls = (unsigned)(state & 7);
    uint32_t rc = 0;
    if (cls < 4) {
      const unsigned long long d = dst > src ? dst - src : src - dst;
      const unsigned long long da = d - d % SYNTH_ALIGN;
      src = dst > src ? dst - da : (dst + da + len <= n_bases ? dst + da : dst);
    } else if (cls == 4) {
      rc = 1;
      const unsigned long long want = (SYNTH_ALIGN - (src + len) % SYNTH_ALIGN + dst % SYNTH_ALIGN) % SYNTH_ALIGN;
      if (src + want + len <= n_bases) src += want;
      else if (src >= SYNTH_ALIGN - want) src -= SYNTH_ALIGN - want;
    }
    out.push_back(Repeat{dst, src, len, rc, 0});
    cursor = dst + len;
  }
  return out;
}

constexpr int SY_THREADS = 256;
constexpr int SY_PER_THREAD = 16;

__global__ void __launch_bounds__(SY_THREADS)
synth_kernel(char* __restrict__ out, unsigned long long first, unsigned long long count, unsigned long long seed,
             const Repeat* __restrict__ reps, uint32_t n_reps) {
  const unsigned long long j0 = ((unsigned long long)blockIdx.x * SY_THREADS + threadIdx.x) * SY_PER_THREAD;
  if (j0 >= count) return;
  const unsigned long long i0 = first + j0;
  // first repeat whose end lies beyond i0
  uint32_t lo = 0, hi = n_reps;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (reps[mid].dst + reps[mid].len <= i0) lo = mid + 1;
    else hi = mid;
  }
  uint32_t r = lo;
  Repeat cur{~0ull, 0, 0, 0, 0};
  if (r < n_reps) cur = reps[r];
  __align__(16) char buf[SY_PER_THREAD];
#pragma unroll
  for (int k = 0; k < SY_PER_THREAD; ++k) {
    const unsigned long long i = i0 + k;
    while (r < n_reps && cur.dst + cur.len <= i) {
      ++r;
      cur = r < n_reps ? reps[r] : Repeat{~0ull, 0, 0, 0, 0};
    }
    uint32_t code;
    if (r < n_reps && i >= cur.dst) {
      const unsigned long long off = i - cur.dst;
      code = cur.rc ? 3u - base_code(seed, cur.src + cur.len - 1 - off) : base_code(seed, cur.src + off);
    } else {
      code = base_code(seed, i);
    }
    buf[k] = "ACGT"[code];
  }
  if (j0 + SY_PER_THREAD <= count && ((reinterpret_cast<uintptr_t>(out + j0) & 15) == 0)) {
    *reinterpret_cast<uint4*>(out + j0) = *reinterpret_cast<const uint4*>(buf);
  } else {
    for (int k = 0; k < SY_PER_THREAD && j0 + k < count; ++k) out[j0 + k] = buf[k];
  }
}

int synth_genome(Ctx& ctx, char* d_out, uint64_t n_bases, uint64_t first, uint64_t count, uint64_t seed,
                 uint32_t repeat_permille) {
  if (first + count > n_bases) return ctx.fail(STB_ERR_INVALID_ARG, "synth range exceeds n_bases");
  if (count == 0) return STB_OK;
  const std::vector<Repeat> reps = make_repeats(n_bases, seed, repeat_permille);
  DevBuf<Repeat> d_reps;
  STB_CUDA(ctx, d_reps.alloc(reps.size(), ctx.stream));
  if (!reps.empty())
    STB_CUDA(ctx, cudaMemcpyAsync(d_reps.ptr, reps.data(), reps.size() * sizeof(Repeat), cudaMemcpyHostToDevice, ctx.stream));
  {
    Launch l(ctx, "synth_genome");
    const uint64_t threads = ceil_div(count, SY_PER_THREAD);
    synth_kernel<<<(unsigned)ceil_div(threads, SY_THREADS), SY_THREADS, 0, ctx.stream>>>(d_out, first, count, seed, d_reps.ptr, (uint32_t)reps.size());
  }
  STB_CUDA(ctx, cudaStreamSynchronize(ctx.stream));
  STB_CUDA(ctx, cudaGetLastError());
  return STB_OK;
}

// "Real genome" variant of the workload: runs of N (unsequenced stretches) and soft-masked (lower-case)
// stretches laid over the text, both pure functions of the position.  Per 64 Ki-base block: with
// probability 8 % one run of 1..16384 N starting in the block's first 48 Ki bases (about 1 % of all
// bases); per 4 Ki-base block: lower case with probability 1/2.  CPU twin: orc_synth_mask.
__host__ __device__ __forceinline__ char masked_base(char c, unsigned long long i, unsigned long long seed) {
  const unsigned long long b = i >> 16;
  const unsigned long long h = splitmix64(seed ^ 0x4e4e4e4e4e4e4e4eull ^ (b * 0x9E3779B97F4A7C15ull));
  if (h % 100ull < 8ull) {
    const unsigned long long start = (b << 16) + ((h >> 8) % 49152ull), len = 1ull + ((h >> 32) % 16384ull);
    if (i >= start && i < start + len) c = 'N';
  }
  const unsigned long long h2 = splitmix64(seed ^ 0x6c6f776572636173ull ^ ((i >> 12) * 0xD1B54A32D192ED03ull));
  if (h2 & 1ull) c = (char)(c | 0x20);
  return c;
}

__global__ void __launch_bounds__(256) synth_mask_kernel(char* __restrict__ text, unsigned long long first, unsigned long long count, unsigned long long seed) {
  const unsigned long long j = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
  if (j < count) text[j] = masked_base(text[j], first + j, seed);
}

int synth_mask(Ctx& ctx, char* d_text, uint64_t first, uint64_t count, uint64_t seed) {
  if (count == 0) return STB_OK;
  Launch l(ctx, "synth_mask");
  synth_mask_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, ctx.stream>>>(d_text, first, count, seed);
  STB_CUDA(ctx, cudaStreamSynchronize(ctx.stream));
  STB_CUDA(ctx, cudaGetLastError());
  return STB_OK;
}

}  // namespace stb
