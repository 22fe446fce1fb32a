// Microbenchmark 2: (a) ld+store into slots previously modified by atomics; (b) effect of
// cudaLimitMaxL2FetchGranularity on random 4-byte loads; (c) store only to a second array.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct __align__(16) Slot { unsigned long long key; uint32_t minpos, id; };
__device__ __forceinline__ uint32_t h32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void k_atomic(Slot* tab, uint32_t cap, uint32_t n) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; if (p >= n) return;
  uint32_t s = __umulhi(h32(p), cap);
  atomicCAS(&tab[s].key, 0xffffffffffffffffull, (unsigned long long)p * 77u);
  atomicMin(&tab[s].minpos, p);
}
template <int MODE>
__global__ void k(Slot* tab, uint32_t cap, uint32_t n, uint2* uniq, uint32_t* tmp) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; if (p >= n) return;
  uint32_t s = __umulhi(h32(p), cap);
  if (MODE == 0) { unsigned long long kk = __ldcg(&tab[s].key); tab[s].id = p; uniq[p] = make_uint2(kk >> 32, kk); tmp[p] = p; }
  if (MODE == 1) { unsigned long long kk = __ldcg(&tab[s].key); uniq[p] = make_uint2(kk >> 32, kk); tmp[p] = p; }  // no slot store
  if (MODE == 2) { tab[s].id = p; tmp[p] = p; }                                                                   // no key load
  if (MODE == 3) { uint32_t v = __ldcg(&tab[s].minpos); if (v == 12345u) tmp[p] = 1; }
  if (MODE == 4) { uint32_t v = __ldcg(&tmp[__umulhi(h32(p), n)]); if (v == 0x12345u) tmp[p] = 1; }              // random 4B loads in dense array
}
template <int MODE> float run(Slot* tab, uint32_t cap, uint32_t n, uint2* uniq, uint32_t* tmp, bool atom) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9;
  for (int r = 0; r < 3; ++r) {
    cudaMemset(tab, 0xff, (size_t)cap * 16);
    if (atom) k_atomic<<<(n + 255) / 256, 256>>>(tab, cap, n);
    cudaEventRecord(a); k<MODE><<<(n + 255) / 256, 256>>>(tab, cap, n, uniq, tmp); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
  }
  return best;
}
int main() {
  uint32_t n = 20000000u, cap = 2 * n; Slot* tab; uint2* uniq; uint32_t* tmp;
  cudaMalloc(&tab, (size_t)cap * 16); cudaMalloc(&uniq, (size_t)n * 8); cudaMalloc(&tmp, (size_t)n * 4);
  for (int gran : {0, 32, 64, 128}) {
    if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("L2 fetch granularity limit = %zu\n", g);
    printf("  after memset : ld+3stores %.3f  ld+2stores(no slot st) %.3f  slot st only %.3f  ld minpos %.3f  dense 4B ld %.3f ms\n",
           run<0>(tab, cap, n, uniq, tmp, false), run<1>(tab, cap, n, uniq, tmp, false), run<2>(tab, cap, n, uniq, tmp, false), run<3>(tab, cap, n, uniq, tmp, false), run<4>(tab, cap, n, uniq, tmp, false));
    printf("  after atomics: ld+3stores %.3f  ld+2stores(no slot st) %.3f  slot st only %.3f  ld minpos %.3f ms\n",
           run<0>(tab, cap, n, uniq, tmp, true), run<1>(tab, cap, n, uniq, tmp, true), run<2>(tab, cap, n, uniq, tmp, true), run<3>(tab, cap, n, uniq, tmp, true));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
