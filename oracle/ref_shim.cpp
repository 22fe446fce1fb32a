/*
 * oracle/ref_shim.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * A C-ABI probe around the UNMODIFIED reference implementation.  It is
 * compiled together with the reference's own translation units, in place, from
 * /root/reference (see oracle/Makefile, target `ref`), and the outputs go to
 * oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).
 * Nothing of the reference is copied into this repository: this file only
 * *calls* the public API declared in /root/reference/include/shared_tree.h,
 * dna.h and fasta_reader.h.
 *
 * Uses:
 *   - tests/ and oracle/gen_golden.py: pin the C restatement (oracle/oracle.c)
 *     and the CUDA path against the real reference;
 *   - bench.py `--impl reference` and the `cpu_baseline` leg: time the
 *     reference's CPU build (kind "reference").
 * The product (genome-compression_b200/) never links or loads this.
 */
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "shared_tree.h"
#include "dna.h"
#include "fasta_reader.h"

namespace {

struct ref_tree {
  shared_tree tree;
  int dna_size;
};

inline std::uint32_t raw_of(const pointer& p) {
  // Layout used throughout this repo: bits 0-28 index, 29 mirror,
  // 30 transpose, 31 invariant.  Rebuilt from the public accessors.
  const std::uint32_t index = p.empty() ? 0x1fffffffu : (std::uint32_t)p.index();
  return index | ((std::uint32_t)p.is_mirrored() << 29) |
         ((std::uint32_t)p.is_transposed() << 30) |
         ((std::uint32_t)p.is_invariant() << 31);
}

inline pointer pointer_of(std::uint32_t raw) {
  // Reconstructs a pointer from the raw layout without passing through the
  // normalising 4-argument constructor unless the input is already normalised.
  const auto index = raw & 0x1fffffffu;
  const bool m = (raw >> 29) & 1, t = (raw >> 30) & 1, inv = (raw >> 31) & 1;
  if (index == 0x1fffffffu && !m && !t && inv) return pointer{nullptr};
  return pointer{index, m, t, inv};
}

}  // namespace

extern "C" {

/* ---- leaf (dna) primitives: include/dna.h, src/dna.cpp ------------------ */

std::uint64_t ref_dna_pack(const char* text, int dna_size) {
  dna::size(dna_size);
  return dna{std::string_view{text, (std::size_t)dna_size}}.to_ullong();
}

std::uint64_t ref_dna_transposed(std::uint64_t v, int dna_size) {
  dna::size(dna_size);
  return dna{(unsigned long long)v}.transposed().to_ullong();
}

std::uint64_t ref_dna_mirrored(std::uint64_t v, int dna_size) {
  dna::size(dna_size);
  return dna{(unsigned long long)v}.mirrored().to_ullong();
}

std::uint64_t ref_dna_inverted(std::uint64_t v, int dna_size) {
  dna::size(dna_size);
  return dna{(unsigned long long)v}.inverted().to_ullong();
}

/* out_flags: bit0 mirror, bit1 transpose, bit2 invariant. */
std::uint64_t ref_dna_canonical(std::uint64_t v, int dna_size, int* out_flags) {
  dna::size(dna_size);
  const auto [canonical, mirror, transpose, invariant] = dna{(unsigned long long)v}.canonical();
  *out_flags = (int)mirror | ((int)transpose << 1) | ((int)invariant << 2);
  return canonical.to_ullong();
}

/* ---- pointer / node primitives: include/shared_tree.h ------------------- */

std::uint32_t ref_pointer_make(std::uint64_t index, int m, int t, int inv) {
  return raw_of(pointer{(std::size_t)index, (bool)m, (bool)t, (bool)inv});
}

std::uint32_t ref_pointer_null() { return raw_of(pointer{nullptr}); }

std::uint32_t ref_pointer_compose(std::uint32_t raw, int m, int t) {
  return raw_of(pointer{pointer_of(raw), (bool)m, (bool)t});
}

/* Canonical form of node (l, r); out[0], out[1] = canonical children raw,
 * returns flags bit0 mirror, bit1 transpose, bit2 = emplace_node's invariant. */
int ref_node_canonical(std::uint32_t l, std::uint32_t r, std::uint32_t* out) {
  const auto left = pointer_of(l), right = pointer_of(r);
  const auto [canonical, mirror, transpose] = node{left, right}.canonical();
  out[0] = raw_of(canonical.left());
  out[1] = raw_of(canonical.right());
  const bool invariant = (left == right.mirrored());
  return (int)mirror | ((int)transpose << 1) | ((int)invariant << 2);
}

int ref_pointer_serialize(std::uint32_t raw, std::uint8_t* out) {
  std::ostringstream os;
  pointer_of(raw).serialize(os);
  const auto s = os.str();
  std::memcpy(out, s.data(), s.size());
  return (int)s.size();
}

/* ---- ingest: include/fasta_reader.h -------------------------------------- */

/* read_genome(path) (src/fasta_reader.cpp:108).  Returns the number of leaves;
 * copies min(count, cap) of them. */
std::uint64_t ref_read_genome(const char* path, int dna_size, std::uint64_t* out, std::uint64_t cap) {
  dna::size(dna_size);
  const auto data = read_genome(path);
  for (std::uint64_t i = 0; i < data.size() && i < cap; ++i) out[i] = data[i].to_ullong();
  return data.size();
}

/* ---- shared_tree --------------------------------------------------------- */

void* ref_tree_from_file(const char* path, int dna_size) {
  dna::size(dna_size);
  auto* t = new ref_tree{shared_tree{std::filesystem::path{path}}, dna_size};
  return t;
}

void* ref_tree_from_leaves(const std::uint64_t* leaves, std::uint64_t n, int dna_size) {
  dna::size(dna_size);
  std::vector<dna> data;
  data.reserve(n);
  for (std::uint64_t i = 0; i < n; ++i) data.emplace_back((unsigned long long)leaves[i]);
  return new ref_tree{shared_tree{data}, dna_size};
}

void* ref_tree_deserialize(const std::uint8_t* bytes, std::uint64_t len, int dna_size) {
  dna::size(dna_size);
  std::istringstream is{std::string{(const char*)bytes, (std::size_t)len}};
  return new ref_tree{shared_tree::deserialize(is), dna_size};
}

void ref_tree_free(void* h) { delete (ref_tree*)h; }

std::uint64_t ref_tree_depth(void* h) { return ((ref_tree*)h)->tree.depth(); }
std::uint64_t ref_tree_width(void* h) {
  auto* t = (ref_tree*)h;
  dna::size(t->dna_size);
  return t->tree.width();
}
std::uint64_t ref_tree_leaf_count(void* h) { return ((ref_tree*)h)->tree.leaf_count(); }
std::uint64_t ref_tree_node_count(void* h) { return ((ref_tree*)h)->tree.node_count(); }
std::uint64_t ref_tree_layer_count(void* h, std::uint64_t layer) {
  return ((ref_tree*)h)->tree.node_count(layer);
}

void ref_tree_sort(void* h) {
  auto* t = (ref_tree*)h;
  dna::size(t->dna_size);
  t->tree.sort_tree(false);
}

std::uint64_t ref_tree_bytes(void* h) {
  auto* t = (ref_tree*)h;
  dna::size(t->dna_size);
  return t->tree.bytes();
}

/* Serializes into out (if cap suffices); returns the stream length. */
std::uint64_t ref_tree_serialize(void* h, std::uint8_t* out, std::uint64_t cap) {
  auto* t = (ref_tree*)h;
  dna::size(t->dna_size);
  std::ostringstream os;
  t->tree.serialize(os);
  const auto s = os.str();
  if (s.size() <= cap) std::memcpy(out, s.data(), s.size());
  return s.size();
}

/* Copies the stored (canonical) leaf table. */
std::uint64_t ref_tree_leaves(void* h, std::uint64_t* out, std::uint64_t cap) {
  auto* t = (ref_tree*)h;
  const auto n = t->tree.leaf_count();
  for (std::uint64_t i = 0; i < n && i < cap; ++i)
    out[i] = t->tree.access_leaf(pointer{(std::size_t)i, false, false, false}).to_ullong();
  return n;
}

/* Copies one node layer as raw (left, right) pairs. */
std::uint64_t ref_tree_layer(void* h, std::uint64_t layer, std::uint32_t* out, std::uint64_t cap_nodes) {
  auto* t = (ref_tree*)h;
  const auto n = t->tree.node_count(layer);
  for (std::uint64_t i = 0; i < n && i < cap_nodes; ++i) {
    const auto nd = t->tree.access_node(layer, pointer{(std::size_t)i, false, false, false});
    out[2 * i] = raw_of(nd.left());
    out[2 * i + 1] = raw_of(nd.right());
  }
  return n;
}

std::uint64_t ref_tree_histogram(void* h, std::uint64_t layer, std::uint64_t* out, std::uint64_t cap) {
  auto* t = (ref_tree*)h;
  const auto hist = t->tree.histogram(layer);
  for (std::uint64_t i = 0; i < hist.size() && i < cap; ++i) out[i] = hist[i];
  return hist.size();
}

/* Sequential decode with the reference iterator. Returns number of leaves. */
std::uint64_t ref_tree_decode(void* h, std::uint64_t* out, std::uint64_t cap) {
  auto* t = (ref_tree*)h;
  dna::size(t->dna_size);
  std::uint64_t i = 0;
  for (dna d : t->tree) {
    if (i < cap) out[i] = d.to_ullong();
    ++i;
  }
  return i;
}

void ref_tree_random_access(void* h, const std::uint64_t* idx, std::uint64_t q, std::uint64_t* out) {
  auto* t = (ref_tree*)h;
  dna::size(t->dna_size);
  for (std::uint64_t i = 0; i < q; ++i) out[i] = t->tree[idx[i]].to_ullong();
}

/* Per-level pointer arrays straight from tree_constructor (public API,
 * include/shared_tree.h:254-269): level 0 = reduce_leaves, then reduce_nodes.
 * Writes every level's pointer array back to back into out (raw layout);
 * level_sizes receives the array lengths.  Only valid for n <= 2^25 leaves
 * (one reduce_segment).  Returns number of levels written. */
std::uint64_t ref_build_levels(const std::uint64_t* leaves, std::uint64_t n, int dna_size,
                               std::uint32_t* out, std::uint64_t cap, std::uint64_t* level_sizes,
                               std::uint64_t max_levels) {
  dna::size(dna_size);
  std::vector<dna> data;
  for (std::uint64_t i = 0; i < n; ++i) data.emplace_back((unsigned long long)leaves[i]);
  shared_tree tree;
  tree_constructor ctor{tree};
  auto layer = ctor.reduce_leaves(data);
  std::uint64_t levels = 0, written = 0;
  auto dump = [&](const std::vector<pointer>& l) {
    if (levels < max_levels) level_sizes[levels] = l.size();
    for (const auto& p : l) {
      if (written < cap) out[written] = raw_of(p);
      ++written;
    }
    ++levels;
  };
  dump(layer);
  for (auto index = 1u; layer.size() > 1; ++index) {
    layer = ctor.reduce_nodes(layer, index);
    dump(layer);
  }
  return levels;
}

}  // extern "C"
