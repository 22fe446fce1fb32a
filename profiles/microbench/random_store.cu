// Microbenchmark: why is a random 4-byte store into a 16-byte hash slot slow?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o random_store random_store.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct __align__(16) Slot { unsigned long long key; uint32_t minpos, id; };
__device__ __forceinline__ uint32_t h32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
template <int MODE>
__global__ void k(Slot* tab, uint32_t cap, uint32_t n, uint32_t* sink) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t s = __umulhi(h32(p), cap);
  if (MODE == 0) { tab[s].id = p; }                                             // store only
  if (MODE == 1) { unsigned long long kk = __ldcg(&tab[s].key); tab[s].id = (uint32_t)kk + p; }   // ld.cg then store
  if (MODE == 2) { unsigned long long kk = tab[s].key; tab[s].id = (uint32_t)kk + p; }           // plain ld then store
  if (MODE == 3) { atomicMin(&tab[s].minpos, p); }                               // atomic only
  if (MODE == 4) { uint32_t v = __ldcg(&tab[s].minpos); if (v == p) sink[0] = 1; }               // load only
  if (MODE == 5) { atomicExch(&tab[s].id, p); }                                  // atomic exch (returns)
  if (MODE == 6) { unsigned long long kk = __ldcg(&tab[s].key); uint4 v = make_uint4((uint32_t)kk, (uint32_t)(kk>>32), p, p); *reinterpret_cast<uint4*>(&tab[s]) = v; } // full 16B store
  if (MODE == 7) { asm volatile("st.global.cg.u32 [%0], %1;" :: "l"(&tab[s].id), "r"(p)); }   // st.cg
  if (MODE == 8) { asm volatile("red.global.max.u32 [%0], %1;" :: "l"(&tab[s].id), "r"(p)); }  // red (no return)
}
template <int MODE> float run(Slot* tab, uint32_t cap, uint32_t n, uint32_t* sink) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaMemset(tab, 0xff, (size_t)cap * 16);
  k<MODE><<<(n + 255) / 256, 256>>>(tab, cap, n, sink);
  cudaMemset(tab, 0xff, (size_t)cap * 16);
  cudaEventRecord(a); k<MODE><<<(n + 255) / 256, 256>>>(tab, cap, n, sink); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  const char* names[] = {"store4", "ldcg+store4", "ld+store4", "atomicMin", "ldcg only", "atomicExch", "ldcg+store16", "st.cg 4", "red.max"};
  uint32_t* sink; cudaMalloc(&sink, 4);
  for (uint32_t n : {100000u, 20000000u}) {
    uint32_t cap = 2 * n; Slot* tab; cudaMalloc(&tab, (size_t)cap * 16);
    float ms[9] = {run<0>(tab, cap, n, sink), run<1>(tab, cap, n, sink), run<2>(tab, cap, n, sink), run<3>(tab, cap, n, sink), run<4>(tab, cap, n, sink), run<5>(tab, cap, n, sink), run<6>(tab, cap, n, sink), run<7>(tab, cap, n, sink), run<8>(tab, cap, n, sink)};
    for (int i = 0; i < 9; ++i) printf("n=%u cap=%u %-14s %8.3f ms  %7.2f G/s\n", n, cap, names[i], ms[i], n / ms[i] / 1e6);
    cudaFree(tab);
  }
  cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
}
