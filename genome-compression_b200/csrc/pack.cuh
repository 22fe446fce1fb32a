// pack.cuh — ASCII body -> 4-bit packed leaves, staged through shared memory.
//
// Replaces to_nac / dna::dna(string_view) / dna::set (reference src/dna.cpp:25-49,
// :79-84, :187-197).  One CTA stages a tile of TILE_LEAVES*S text bytes with
// coalesced 16-byte loads, then every thread packs leaves out of shared memory
// (stride-S reads: conflict-free for S = 12, three 32-bit words per leaf) through a
// 256-entry code table, also in shared memory.
#pragma once

#include "common.cuh"

namespace stb {

constexpr int PACK_THREADS = 256;
constexpr int PACK_LEAVES_PER_THREAD = 4;
constexpr int PACK_TILE_LEAVES = PACK_THREADS * PACK_LEAVES_PER_THREAD;  // 1024

// code table: 0..15 = nucleotide code (reference include/dna.h:20-32), 0xff = unknown
__device__ __forceinline__ uint8_t ascii_code(uint32_t c) {
  if (c >= 'a' && c <= 'z') c -= 32;  // std::toupper, "C" locale
  switch (c) {
    case 'A': return 0x1;
    case 'C': return 0x2;
    case 'G': return 0x4;
    case 'T': return 0x8;
    case 'R': return 0x3;
    case 'Y': return 0xC;
    case 'K': return 0x7;
    case 'M': return 0xE;
    case 'S': return 0x0;
    case 'W': return 0x9;
    case 'B': return 0x5;
    case 'D': return 0xB;
    case 'H': return 0xD;
    case 'V': return 0xA;
    case 'N': return 0x6;
    case '-': return 0xF;
    default: return 0xFF;
  }
}

struct PackSmem {
  uint8_t lut[256];
  // text tile follows (dynamic), 16-byte aligned
};

// Fills the code table and stages `tile_bytes` of text starting at `src` (16-byte
// aligned) into `tile`.  Must be called by all PACK_THREADS threads.
__device__ __forceinline__ void stage_text_tile(uint8_t* lut, uint8_t* tile, const char* __restrict__ src,
                                                uint32_t tile_bytes) {
  lut[threadIdx.x] = ascii_code(threadIdx.x);
  const uint32_t vecs = tile_bytes >> 4;
  const uint4* __restrict__ src4 = reinterpret_cast<const uint4*>(src);
  uint4* tile4 = reinterpret_cast<uint4*>(tile);
  for (uint32_t i = threadIdx.x; i < vecs; i += PACK_THREADS) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(src4 + i));
    tile4[i] = v;
  }
  for (uint32_t i = (vecs << 4) + threadIdx.x; i < tile_bytes; i += PACK_THREADS) tile[i] = (uint8_t)src[i];
  __syncthreads();
}

// Packs leaf `j` of the staged tile.  S_T > 0 fixes the leaf length at compile time.
// On an unknown symbol, *bad_off receives the in-tile byte offset (smallest one).
template <int S_T>
__device__ __forceinline__ unsigned long long pack_leaf(const uint8_t* lut, const uint8_t* tile, uint32_t j, int S_rt,
                                                        uint32_t& bad_off) {
  const int S = S_T > 0 ? S_T : S_rt;
  unsigned long long v = 0;
  const uint32_t base = j * (uint32_t)S;
  if (S_T > 0 && (S_T % 4) == 0) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(tile + base);
#pragma unroll
    for (int q = 0; q < S_T / 4; ++q) {
      const uint32_t word = w[q];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint32_t code = lut[(word >> (8 * b)) & 0xFFu];
        if (code > 15u && bad_off == 0xFFFFFFFFu) bad_off = base + 4 * q + b;
        v |= (unsigned long long)(code & 15u) << (4 * (4 * q + b));
      }
    }
  } else {
#pragma unroll 4
    for (int i = 0; i < S; ++i) {
      const uint32_t code = lut[tile[base + i]];
      if (code > 15u && bad_off == 0xFFFFFFFFu) bad_off = base + i;
      v |= (unsigned long long)(code & 15u) << (4 * i);
    }
  }
  return v;
}

static inline size_t pack_smem_bytes(int S) { return 256 + (size_t)PACK_TILE_LEAVES * S + 16; }

}  // namespace stb
