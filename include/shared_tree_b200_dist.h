/*
 * shared_tree_b200_dist.h — stage entry points of the multi-GPU build (C ABI).
 *
 * The reference is a single process (no collectives).  The sharded build of
 * BASELINE.json config 4 keeps its semantics exactly: node ids are first-occurrence ranks in
 * GLOBAL position order.  One process per GPU; every rank owns a contiguous, power-of-two
 * aligned range of leaf positions (the analogue of the reference's own 2^22 / 2^25-leaf
 * segments, include/shared_tree.h:305-316, src/shared_tree.cpp:719-763).  Per level:
 *
 *   1. partition   each rank canonicalises its positions (dna::canonical / node::canonical)
 *                  and splits (key, global position) records by hash owner          [stage]
 *   2. exchange    all-to-all of the records (NCCL, done by the caller)
 *   3. owner       the owner dedups its keys with min-position, answers every record with the
 *                  global position of the key's first occurrence and marks that position in
 *                  a bitmap over the level's global positions                        [stage]
 *   4. exchange    all-to-all of the answers back; all-reduce (sum = or: bits are disjoint)
 *                  of the bitmap
 *   5. rank index  exclusive popcount prefix per bitmap word: id(q) = rank of bit q  [stage]
 *   6. finish      first occurrences append their item to the rank's slice of the layer
 *                  (ids [base, base+count)), every position gets pointer{id(q), flags} [stage]
 *
 * The collectives live in the caller (genome-compression_b200/dist.py uses
 * torch.distributed); these stages only see device pointers.  All buffers are caller
 * allocated; `ctx` is any handle from stb_create (device, stream, dna_size, error text).
 */
#ifndef SHARED_TREE_B200_DIST_H
#define SHARED_TREE_B200_DIST_H

#include "shared_tree_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* body (device ASCII, no header/newline) -> packed leaves (device). dna.cpp:79-84. */
int stb_dist_pack_body(stb_tree* ctx, const char* body_dev, uint64_t n_leaves, uint64_t* leaves_dev);

/* Stage 1.  kind 0: items are packed leaves (uint64[n_items]); kind 1: items are the child
 * pointer array (uint32[n_items]) and positions are pairs of children (odd tail -> null).
 * n_positions = n_items (kind 0) or ceil(n_items/2) (kind 1).  Outputs, in owner order
 * (owner 0's records first, stable within an owner): keys[n_positions], gpos[n_positions]
 * (= gpos0 + local position), meta[n_positions] (= local position | flags << 29),
 * counts[world] (records per owner). */
int stb_dist_partition(stb_tree* ctx, int kind, const void* items_dev, uint64_t n_items, uint64_t gpos0, int world,
                       uint64_t* keys_dev, uint32_t* gpos_dev, uint32_t* meta_dev, uint32_t* counts_dev);

/* Stage 3.  table_dev: (cap + 1) * 16 bytes, cap >= 2 * n_records (cleared here);
 * answers_dev[n_records] = global position of the first occurrence of the record's key;
 * bitmap_dev (caller-zeroed, ceil(n_level_positions/32) words): bit q set for every first
 * occurrence q owned here. */
int stb_dist_owner(stb_tree* ctx, const uint64_t* keys_dev, const uint32_t* gpos_dev, uint64_t n_records, void* table_dev,
                   uint32_t cap, uint32_t* answers_dev, uint32_t* bitmap_dev);

/* Stage 5.  word_prefix_dev[n_words + 1]: exclusive popcount prefix (last entry = total);
 * scratch_dev: ceil(n_words/1024) + 1 words. */
int stb_dist_rank_index(stb_tree* ctx, const uint32_t* bitmap_dev, uint64_t n_words, uint32_t* word_prefix_dev,
                        uint32_t* scratch_dev);

/* Stage 6.  Same kind / items / gpos0 as stage 1.  answers_dev / meta_dev are in the
 * rank's SEND order (what stage 1 produced, answers as returned by the exchange).
 * Outputs: pointers_dev[n_positions]; layer_slice_dev: the items first seen in this rank's
 * range, in id order (uint64 leaves or uint2 nodes), capacity n_positions;
 * base_count_dev[2] = {first id of the slice, number of items}. */
int stb_dist_finish(stb_tree* ctx, int kind, const void* items_dev, uint64_t n_items, uint64_t gpos0,
                    const uint32_t* bitmap_dev, const uint32_t* word_prefix_dev, uint64_t n_level_positions,
                    const uint32_t* meta_dev, const uint32_t* answers_dev, uint32_t* pointers_dev, void* layer_slice_dev,
                    uint32_t* base_count_dev);

/* Leaf level for ACGT-only input with dna_size <= 12: every rank lowers a replicated
 * direct-addressed table (4^dna_size uint32, caller-filled with 0x7f7f7f7f) with its GLOBAL
 * leaf positions straight from the text; the caller all-reduces the table with MIN (as int32);
 * then every rank derives all leaf ids locally.  *non_acgt = 1: the shard holds other symbols,
 * use the generic stages (stb_dist_pack_body + partition ...) for the leaf level instead.
 * tmp_dev[n_local] carries slot | flags between the two calls. */
int stb_dist_leaf_direct_minpos(stb_tree* ctx, const char* body_dev, uint64_t n_local, uint64_t gpos0, uint32_t* table_dev,
                                uint32_t* tmp_dev, int* non_acgt);
/* bitmap_dev: caller-zeroed, ceil(n_level_positions/32) words; word_prefix_dev: that + 1;
 * scratch_dev as in stb_dist_rank_index; ids_dev: 4^dna_size words; leaves_out_dev receives the
 * WHOLE leaf table in id order (capacity min(n_level_positions, 4^dna_size)); total =
 * word_prefix_dev[n_words]. */
int stb_dist_leaf_direct_finish(stb_tree* ctx, const uint32_t* table_dev, uint64_t n_level_positions, const uint32_t* tmp_dev,
                                uint64_t n_local, uint32_t* bitmap_dev, uint32_t* word_prefix_dev, uint32_t* scratch_dev,
                                uint32_t* ids_dev, uint32_t* pointers_dev, uint64_t* leaves_out_dev);

/* Remaining (small) levels on one rank: builds node layers from a pointer array until one
 * pointer is left.  The handle afterwards holds ONLY those layers (stb_layer_count,
 * stb_copy_layer, stb_root work; leaf_count is 0).  leaf_pointers != 0: the array is the leaf
 * level itself, so at least one node layer is built even for a single pointer (a lone leaf is
 * wrapped as node{leaf, null}, include/shared_tree.h:282-299). */
int stb_dist_upper_levels(stb_tree* tree, const uint32_t* pointers_dev, uint64_t n_pointers, int leaf_pointers);

/* Assembles a complete tree from device arrays (copies them): the gathered result of a
 * sharded build, or any externally produced tree.  layer_counts[n_layers],
 * layers_dev[n_layers] (uint2 nodes each). */
int stb_assemble(stb_tree* tree, const uint64_t* leaves_dev, uint64_t n_leaves, uint64_t n_layers,
                 const uint64_t* layer_counts, const void* const* layers_dev, uint32_t root, uint64_t width);

/* ---- peer exchange: stages 1-4 with the record exchange fused into the kernels ---------------
 *
 * Every rank allocates one ARENA (stb_dist_peer_alloc: plain cudaMalloc + a 64-byte CUDA IPC
 * handle), the ranks swap handles (any host channel; dist.py uses all_gather_object) and map each
 * other's arenas (stb_dist_peer_open).  `arenas[world]` below are those mapped base pointers,
 * arenas[rank] the rank's own.  region_cap = the largest number of positions any rank holds at
 * the level (every rank must pass the same value; the arena must hold
 * stb_dist_peer_arena_bytes(world, region_cap)).  Per level:
 *
 *   stb_dist_peer_scatter   stage 1 + 2: records are grouped by owner in shared memory and
 *                           written straight into the owners' arenas (NVLink stores)
 *   -- barrier A: any collective on the same stream (dist.py: 1-element all-reduce) --
 *   stb_dist_peer_owner     stage 3 + 4a: dedup; LATER occurrences get their answer written into
 *                           the source's arena; first occurrences are marked in the bitmap only
 *   -- all-reduce(sum) of the bitmap: also barrier B --
 *   stb_dist_rank_index, stb_dist_peer_finish
 *
 * Nothing is read back by the host inside a level.  */
uint64_t stb_dist_peer_arena_bytes(int world, uint64_t region_cap);
int stb_dist_peer_alloc(stb_tree* ctx, uint64_t bytes, void** ptr_out, unsigned char* handle_out /* [64] */);
int stb_dist_peer_open(stb_tree* ctx, const unsigned char* handle /* [64] */, void** ptr_out);
int stb_dist_peer_close(stb_tree* ctx, void* ptr);
int stb_dist_peer_free(stb_tree* ctx, void* ptr);
/* meta_dev[world * region_cap]: local position | flags of every record sent, in send order
 * (region o: the records that went to owner o). */
int stb_dist_peer_scatter(stb_tree* ctx, int kind, const void* items_dev, uint64_t n_items, uint64_t gpos0, int world, int rank,
                          void* const* arenas, uint64_t region_cap, uint32_t* meta_dev);
/* expected_records: the rank's fair share (sizes grids and the singleton filter; any number of
 * records up to world * region_cap is handled).  table_dev: table_slots * 16 bytes with
 * table_slots >= 2 * world * region_cap + 1 (only 2 * records of them are touched), filled with
 * 0xff bytes ONCE by the caller and never cleared again: slots carry an epoch tag, `serial` must
 * be a value this table has not seen before (1, 2, 3, ...);
 * slot_scratch_dev: world * region_cap words; planes_dev (may be NULL): planes_words words for the
 * singleton filter; bitmap_dev: caller-zeroed, ceil(n_level_positions / 32) words. */
int stb_dist_peer_owner(stb_tree* ctx, int world, int rank, void* const* arenas, uint64_t region_cap, uint64_t expected_records,
                        void* table_dev, uint64_t table_slots, uint32_t serial, uint32_t* slot_scratch_dev, uint32_t* planes_dev,
                        uint64_t planes_words, uint32_t* bitmap_dev);
/* stb_dist_finish for the peer exchange: meta_dev as written by stb_dist_peer_scatter, the
 * answers and the per-owner record counts are read from the rank's own arena. */
int stb_dist_peer_finish(stb_tree* ctx, int kind, const void* items_dev, uint64_t n_items, uint64_t gpos0,
                         const uint32_t* bitmap_dev, const uint32_t* word_prefix_dev, uint64_t n_level_positions,
                         const uint32_t* meta_dev, const void* arena, int world, uint64_t region_cap, uint32_t* pointers_dev,
                         void* layer_slice_dev, uint32_t* base_count_dev);
/* A free area of the arena between levels (the record regions: world * region_cap * 8 bytes),
 * and a stream-ordered copy into it (dst may be a peer's): used to collect the last sharded
 * level's pointers on rank 0. */
void* stb_dist_peer_payload(void* arena, int world, uint64_t region_cap);
int stb_dist_peer_put(stb_tree* ctx, void* dst_dev, const void* src_dev, uint64_t bytes);

/* ==== the sharded build behind ONE call per rank (csrc/shard.cu) ===============================
 *
 * tree_constructor::reduce (src/shared_tree.cpp:719-736) is one call; so is a rank's share of the
 * sharded build.  The library holds the communicator (NCCL, looked up at run time), maps the ranks'
 * arenas into each other (CUDA IPC) and enqueues every level without reading anything back.
 * World sizes: powers of two up to 16, all ranks on one node.  ACGT text, dna_size <= 12 (anything
 * else: use the single-GPU build).  The stage calls above remain for callers that bring their own
 * collectives (genome-compression_b200/dist.py). */
typedef struct stb_shard stb_shard; /* opaque: one rank */
#define STB_SHARD_ID_BYTES 128

/* Rank 0 makes the communicator id; the caller hands the bytes to every rank (any transport). */
int stb_shard_unique_id(uint8_t id[STB_SHARD_ID_BYTES]);
/* Collective over the ranks of `id`. */
int stb_shard_create(stb_shard** out, int device, int dna_size, void* cuda_stream, int rank, int world,
                     const uint8_t id[STB_SHARD_ID_BYTES]);
/* `world` virtual ranks in THIS process on one device (each with its own stream), for tests and for boxes
 * with fewer GPUs than ranks: out[0..world).  Drive each from its own thread. */
int stb_shard_create_local(stb_shard** out, int world, int device, int dna_size);
int stb_shard_destroy(stb_shard* shard);
/* The options of stb_set_option, plus "cut": levels with at most this many positions are finished on rank 0. */
int stb_shard_set_option(stb_shard* shard, const char* name, uint64_t value);
/* Which bases of a body of n_bases_total this rank builds from (a contiguous, power-of-two aligned leaf range). */
int stb_shard_range(const stb_shard* shard, uint64_t n_bases_total, uint64_t* first_base, uint64_t* base_count);
/* Collective.  body_local: this rank's bases (stb_shard_range), host or device memory. */
int stb_shard_build_from_body(stb_shard* shard, const char* body_local, uint64_t n_bases_total, int memory);
/* Unique items of the sharded levels (leaves first), known to every rank after a build. */
int stb_shard_layer_totals(const stb_shard* shard, uint64_t* out, uint64_t cap, uint64_t* count);
/* Collective.  Rank 0 receives the complete tree in `out` (any handle from stb_create on its device): the
 * other ranks pass NULL.  sort_tree / serialize / decode then run on it as on a single-GPU tree. */
int stb_shard_gather(stb_shard* shard, stb_tree* out);
const char* stb_shard_last_error(const stb_shard* shard);
/* Per-kernel device timing of this rank: on = 1 / 0 switches it, on < 0 resets the totals; the totals so far
 * are read into names / total_ms / launches (up to cap entries; *count = classes). */
int stb_shard_profile(stb_shard* shard, int on, const char** names, double* total_ms, uint64_t* launches, uint64_t cap,
                      uint64_t* count);

#ifdef __cplusplus
}
#endif
#endif
