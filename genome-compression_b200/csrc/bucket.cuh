// bucket.cuh — interface of the on-chip level deduplication (bucket.cu).
#pragma once

#include <algorithm>

#include "tree.h"

namespace stb {

struct BucketPlan {
  uint64_t n = 0;      // positions of the level
  int b1 = 1, b2 = 1;  // hash bits consumed by the two partition passes
  uint32_t cap1 = 0;   // records a first-pass bucket can hold
  uint32_t cap2 = 0;   // records a final bucket can hold (<= the dedup kernel's shared-memory tile)
  bool usable = false; // false: the level is too large for final buckets of cap2 records
};

struct BucketWorkspace {
  DevBuf<unsigned long long> keys1, keys2;
  DevBuf<uint32_t> pos1, pos2, counters;
};

BucketPlan bucket_plan(uint64_t n, const Options& opt);
int bucket_reserve(Ctx& ctx, BucketWorkspace& ws, const BucketPlan& plan);

// Deduplicates the node level whose children are cur[0, n_cur): afterwards first_bits marks the
// first occurrences, aux[p] = flags(p) | position of p's first occurrence (later occurrences) or
// flags(p) alone (first occurrences), multi_bits marks the first occurrences whose key occurs
// again.  child_first / child_multi (optional): the child level's bitmaps; a position with a child
// that never repeats is a certified singleton and makes no record (build.cu: children_both_repeat).
// first_bits (tile-rounded) and multi_bits must be zero on entry.  *overflow_out is a
// device flag: non-zero means a bucket overflowed and nothing of the above was produced.
int bucket_dedup_level(Ctx& ctx, BucketWorkspace& ws, const BucketPlan& plan, const uint32_t* cur, uint32_t n_cur, uint32_t n_next,
                       const uint32_t* child_first, const uint32_t* child_multi, uint32_t* aux, uint32_t* first_bits, uint32_t* multi_bits, uint32_t** overflow_out);

}  // namespace stb
