/*
 * oracle/oracle.h — CPU restatement of the reference's shared_tree hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this; the product
 * (genome-compression_b200/) never does and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every
 * function below against the unmodified reference compiled into
 * oracle/_ref/libref.so (oracle/ref_shim.cpp), and tests/test_oracle_golden.py
 * checks it against tests/golden/ (values produced by that same reference,
 * generator committed as oracle/gen_golden.py).
 *
 * Raw pointer layout used everywhere in this repo (include/shared_tree.h:73-76
 * of the reference, confirmed by execution): bits 0-28 index, bit 29 mirror,
 * bit 30 transpose, bit 31 invariant.  NULL = 0x9fffffff.
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_IDX_MASK 0x1fffffffu
#define ORC_MIRROR 0x20000000u
#define ORC_TRANSPOSE 0x40000000u
#define ORC_INVARIANT 0x80000000u
#define ORC_KEY31 0x7fffffffu
#define ORC_NULL 0x9fffffffu

/* ---- leaves (reference include/dna.h, src/dna.cpp) ---- */
int orc_code(int ch);                                  /* 4-bit code, -1 if unknown */
uint64_t orc_pack(const char* text, int S, int* bad);  /* *bad = offending byte or -1 */
uint64_t orc_transposed(uint64_t v);
uint64_t orc_mirrored(uint64_t v, int S);
uint64_t orc_inverted(uint64_t v, int S);
/* flags: bit0 mirror, bit1 transpose, bit2 invariant */
uint64_t orc_leaf_canonical(uint64_t v, int S, int* flags);
void orc_leaf_to_ascii(uint64_t v, int S, char* out);

/* ---- pointers / nodes (reference include/shared_tree.h, src/shared_tree.cpp) ---- */
uint32_t orc_ptr(uint64_t index, int m, int t, int inv);
uint32_t orc_compose(uint32_t p, int m, int t);
int orc_node_canonical(uint32_t l, uint32_t r, uint32_t out[2]);
int orc_ptr_bytes(uint32_t raw);
int orc_ptr_serialize(uint32_t raw, uint8_t* out);

/* ---- ingest (reference src/fasta_reader.cpp:40-68) ---- */
/* Extracts the FASTA body (header / blank-line skipping, newline removal).
 * Returns the body length (before tail truncation). out needs len bytes. */
uint64_t orc_fasta_body(const char* text, uint64_t len, char* out);
/* body -> leaves; returns the leaf count = floor(body_len / S). On an unknown
 * symbol inside the kept part returns UINT64_MAX and sets *bad to the
 * upper-cased byte (what the reference prints before exit(1)). */
uint64_t orc_fasta_to_leaves(const char* text, uint64_t len, int S, uint64_t* out, uint64_t cap,
                             int* bad);

/* ---- tree ---- */
typedef struct orc_tree {
  int S;
  uint32_t root;
  uint64_t n_leaves;
  uint64_t* leaves;
  int n_layers;
  uint64_t* layer_size; /* nodes per layer */
  uint32_t** layers;    /* layers[k][2*i], [2*i+1] = left, right raw */
} orc_tree;

orc_tree* orc_build(const uint64_t* leaves, uint64_t n, int S);
/* As orc_build but also dumps every level's pointer array back to back into
 * out (cap entries); level_sizes[k] = length of level k's array. */
orc_tree* orc_build_levels(const uint64_t* leaves, uint64_t n, int S, uint32_t* out, uint64_t cap,
                           uint64_t* level_sizes, int max_levels, int* n_levels);
void orc_free(orc_tree* t);
void orc_sort(orc_tree* t);
uint64_t orc_bytes(const orc_tree* t);
uint64_t orc_serialize(const orc_tree* t, uint8_t* out, uint64_t cap);
orc_tree* orc_deserialize(const uint8_t* bytes, uint64_t len, int S);
uint64_t orc_width(const orc_tree* t);
uint64_t orc_node_count(const orc_tree* t);
uint64_t orc_decode(const orc_tree* t, uint64_t* out, uint64_t cap);
void orc_random_access(const orc_tree* t, const uint64_t* idx, uint64_t q, uint64_t* out);
/* child layer c: 0 = leaves (referenced from node layer 0), c>0 = node layer c-1 */
void orc_histogram(const orc_tree* t, int layer, uint64_t* out);

/* accessors for ctypes */
uint64_t orc_tree_leaf_count(const orc_tree* t);
int orc_tree_layers(const orc_tree* t);
uint64_t orc_tree_layer_size(const orc_tree* t, int layer);
const uint64_t* orc_tree_leaves(const orc_tree* t);
const uint32_t* orc_tree_layer(const orc_tree* t, int layer);
uint32_t orc_tree_root(const orc_tree* t);

/* ---- synthetic genome (this repo's own workload definition) ---- */
typedef struct orc_repeat {
  uint64_t dst, src, len;
  uint32_t rc; /* 1 = reverse-complement copy */
  uint32_t pad;
} orc_repeat;
uint64_t orc_synth_repeats(uint64_t n_bases, uint64_t seed, uint32_t repeat_permille,
                           orc_repeat* out, uint64_t cap);
void orc_synth_mask(char* text, uint64_t first, uint64_t count, uint64_t seed);
void orc_synth_fill(char* out, uint64_t first, uint64_t count, uint64_t seed,
                    const orc_repeat* reps, uint64_t n_reps);

#ifdef __cplusplus
}
#endif
#endif
