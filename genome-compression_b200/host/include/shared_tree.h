// shared_tree.h — the reference's class surface (include/shared_tree.h) as a shim over the
// B200 C ABI (include/shared_tree_b200.h).  The reference's compress.cpp and tests/test.cpp
// compile against this header unchanged; every tree operation below is one C-ABI call into the
// CUDA library.  `pointer` and `node` are host value types over the same raw 32-bit layout the
// kernels use (bits 0-28 index, 29 mirror, 30 transpose, 31 invariant).
#pragma once

#include <algorithm>
#include <array>
#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iostream>
#include <iterator>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "dna.h"
#include "fasta_reader.h"
#include "shared_tree_b200.h"
#include "utility.h"

class pointer {
  static constexpr std::uint32_t kIndex = 0x1fffffffu, kMirror = 1u << 29, kTranspose = 1u << 30, kInvariant = 1u << 31;

 public:
  static constexpr auto address_bits = std::array{4, 12, 20, 28};

  // null: all index bits set, invariant (so mirroring and transposing leave it alone)
  pointer(std::nullptr_t = nullptr) noexcept : raw_(kIndex | kInvariant) {}
  // re-annotate `other` (src/shared_tree.cpp:76-80)
  pointer(const pointer& other, bool mirror = false, bool transpose = false) noexcept : raw_(other.raw_ & (kIndex | kInvariant)) {
    if (mirror != other.is_mirrored() && !other.is_invariant()) raw_ |= kMirror;
    if (transpose != other.is_transposed() && other != nullptr) raw_ |= kTranspose;
  }
  pointer(std::size_t index, bool mirror, bool transpose, bool invariant) noexcept
      : raw_((std::uint32_t(index) & kIndex) | (mirror && !invariant ? kMirror : 0) | (transpose ? kTranspose : 0) |
             (invariant ? kInvariant : 0)) {}
  static pointer from_raw(std::uint32_t raw) noexcept {
    pointer p;
    p.raw_ = raw;
    return p;
  }

  pointer(pointer&&) noexcept = default;
  pointer& operator=(const pointer&) noexcept = default;
  pointer& operator=(pointer&&) noexcept = default;

  bool empty() const noexcept { return *this == nullptr; }
  std::uint32_t canonical() const noexcept { return raw_ & kIndex; }
  std::size_t index() const noexcept { return raw_ & kIndex; }
  std::uint32_t raw() const noexcept { return raw_; }

  // comparisons ignore the invariant bit (include/shared_tree.h:54-57)
  unsigned long to_ulong() const noexcept { return raw_ & ~kInvariant; }
  bool operator==(const pointer& o) const noexcept { return to_ulong() == o.to_ulong(); }
  bool operator!=(const pointer& o) const noexcept { return to_ulong() != o.to_ulong(); }
  bool operator<(const pointer& o) const noexcept { return to_ulong() < o.to_ulong(); }
  operator bool() const noexcept { return *this != nullptr; }

  bool is_mirrored() const noexcept { return raw_ & kMirror; }
  bool is_transposed() const noexcept { return raw_ & kTranspose; }
  bool is_inverted() const noexcept { return is_mirrored() && is_transposed(); }
  bool is_invariant() const noexcept { return raw_ & kInvariant; }
  pointer mirrored() const noexcept { return pointer{*this, true, false}; }
  pointer transposed() const noexcept { return pointer{*this, false, true}; }
  pointer inverted() const noexcept { return pointer{*this, true, true}; }

  // wire format: 2 bits segment, transpose, mirror, then 4/12/20/28 offset bits
  std::size_t bytes() const noexcept { return segment() + 1; }
  void serialize(std::ostream& os) const {
    const unsigned seg = segment();
    const std::uint32_t off = index() == kIndex ? 0xfffffffu : std::uint32_t(index()) - kStart[seg];
    os.put(char((off >> (8 * seg)) | (is_mirrored() << 4) | (is_transposed() << 5) | (seg << 6)));
    for (unsigned b = seg; b-- > 0;) os.put(char(off >> (8 * b)));
  }
  static pointer deserialize(std::istream& is) {
    const unsigned head = (unsigned char)is.get();
    const unsigned seg = head >> 6;
    std::uint32_t off = head & 0xf;
    for (unsigned b = 0; b < seg; ++b) off = (off << 8) | (unsigned char)is.get();
    const std::size_t index = (seg == 3 && off == 0xfffffffu) ? kIndex : kStart[seg] + off;
    return pointer{index, bool(head & 0x10), bool(head & 0x20), false};
  }

 private:
  static constexpr std::uint32_t kStart[4] = {0u, 16u, 4112u, 1052688u};
  unsigned segment() const noexcept {
    const std::uint32_t i = raw_ & kIndex;
    return (i >= kStart[1]) + (i >= kStart[2]) + (i >= kStart[3]);
  }
  std::uint32_t raw_;
};

inline std::ostream& operator<<(std::ostream& os, const pointer& p) {
  if (p.empty()) return os << "empty";
  return os << '(' << p.index() << ": " << p.is_mirrored() << p.is_transposed() << p.is_invariant() << ')';
}

namespace std {
template <>
struct hash<pointer> {
  std::size_t operator()(const pointer& p) const noexcept { return std::hash<unsigned long>()(p.to_ulong()); }
};
}  // namespace std

class node {
 public:
  node(pointer left, pointer right = nullptr) : children{left, right} {}
  node(const node&) noexcept = default;
  node(node&&) noexcept = default;
  node& operator=(const node&) noexcept = default;
  node& operator=(node&&) noexcept = default;

  bool operator==(const node& o) const noexcept { return children == o.children; }
  bool operator!=(const node& o) const noexcept { return !(*this == o); }
  bool operator<(const node& o) const noexcept { return children < o.children; }

  pointer left() const noexcept { return children[0]; }
  pointer right() const noexcept { return children[1]; }
  node mirrored() const noexcept { return node{children[1].mirrored(), children[0].mirrored()}; }
  node transposed() const noexcept { return node{children[0].transposed(), children[1].transposed()}; }
  node inverted() const noexcept { return node{children[1].inverted(), children[0].inverted()}; }
  // lexicographic minimum of (node, mirror, transpose) over the four variants
  std::tuple<node, bool, bool> canonical() const noexcept {
    return variadic_min(std::tuple{*this, false, false}, std::tuple{mirrored(), true, false},
                        std::tuple{transposed(), false, true}, std::tuple{inverted(), true, true});
  }

  std::size_t bytes() const noexcept { return left().bytes() + right().bytes(); }
  void serialize(std::ostream& os) const {
    left().serialize(os);
    right().serialize(os);
  }
  static node deserialize(std::istream& is) {
    const pointer l = pointer::deserialize(is);
    const pointer r = pointer::deserialize(is);
    return node{l, r};
  }

 private:
  std::array<pointer, 2> children;
};

inline std::ostream& operator<<(std::ostream& os, const node& n) { return os << "node<" << n.left() << ", " << n.right() << '>'; }

namespace std {
template <>
struct hash<node> {
  std::size_t operator()(const node& n) const noexcept {
    return detail::hash(std::hash<pointer>()(n.left()), std::hash<pointer>()(n.right()));
  }
};
}  // namespace std

class shared_tree {
 public:
  shared_tree() = default;
  shared_tree(std::filesystem::path path) : shared_tree{fasta_reader{path}} {}
  shared_tree(fasta_reader file, bool verbose = false) {
    (void)verbose;
    create();
    if (file.untouched()) {
      check(stb_build_from_fasta(h_, file.text().data(), file.text().size(), STB_HOST));
    } else {  // some buffers were already taken: build from what is left, as the reference would
      std::vector<dna> rest, buffer;
      while (file.read_into(buffer)) rest.insert(rest.end(), buffer.begin(), buffer.end());
      build(rest);
    }
  }
  shared_tree(std::vector<dna>& data, bool verbose = false) {
    (void)verbose;
    create();
    build(data);
  }
  shared_tree(const shared_tree& o) { *this = o; }
  shared_tree(shared_tree&& o) noexcept { *this = std::move(o); }
  shared_tree& operator=(const shared_tree& o) {
    if (this == &o) return *this;
    reset();
    if (o.h_) check_on(o.h_, stb_clone(o.h_, &h_));
    return *this;
  }
  shared_tree& operator=(shared_tree&& o) noexcept {
    if (this != &o) {
      reset();
      h_ = o.h_;
      o.h_ = nullptr;
    }
    return *this;
  }
  ~shared_tree() { reset(); }

  std::size_t depth() const { return h_ ? get(stb_depth) : 1; }
  std::size_t width() const { return get(stb_width); }
  std::size_t node_count() const { return get(stb_node_count); }
  std::size_t node_count(std::size_t layer) const {
    uint64_t v = 0;
    check(stb_layer_count(h_, layer, &v));
    return v;
  }
  std::size_t leaf_count() const noexcept { return h_ ? get(stb_leaf_count) : 0; }

  dna access_leaf(pointer p) const {
    cache();
    dna leaf{(unsigned long long)cache_->leaves[p.index()]};
    if (p.is_mirrored()) leaf = leaf.mirrored();
    if (p.is_transposed()) leaf = leaf.transposed();
    return leaf;
  }
  node access_node(std::size_t layer, pointer p) const {
    cache();
    const auto& l = cache_->layers[layer];
    return node{pointer::from_raw(l[2 * p.index()]), pointer::from_raw(l[2 * p.index() + 1])};
  }
  std::size_t children(std::size_t layer, pointer p) const {
    if (p.empty()) return 0;
    const node n = access_node(layer, p);
    if (layer == 0) return !n.left().empty() + !n.right().empty();
    return children(layer - 1, n.left()) + children(layer - 1, n.right());
  }
  // operator[] (src/shared_tree.cpp:268): one batched-random-access call with a single query
  dna operator[](std::uint64_t index) const {
    uint64_t out = 0;
    check(stb_random_access(h_, &index, 1, &out, STB_HOST));
    return dna{(unsigned long long)out};
  }

  std::vector<std::size_t> histogram(std::size_t layer) const {
    const std::size_t n = layer == 0 ? leaf_count() : node_count(layer - 1);
    std::vector<uint64_t> raw(n);
    check(stb_histogram(h_, layer, raw.data(), n, STB_HOST));
    return std::vector<std::size_t>(raw.begin(), raw.end());
  }
  // src/shared_tree.cpp:332-345: per layer, frequencies sorted descending, 1000 per line
  void store_histogram(std::filesystem::path path) const {
    std::ofstream file{path};
    for (std::size_t layer = 0; layer + 1 < depth(); ++layer) {
      auto freq = histogram(layer);
      std::sort(freq.begin(), freq.end(), std::greater<>());
      for (std::size_t i = 0; i < freq.size(); ++i) {
        file << freq[i] << ',';
        if (i % 1000 == 999 || i + 1 == freq.size()) file << '\n';
      }
      file << '\n';
    }
  }

  void sort_tree(bool verbose = false) {
    if (verbose) std::cout << progress_bar("Sorting nodes", 0, 1) << std::flush;
    check(stb_sort_tree(h_));
    cache_.reset();
    if (verbose) std::cout << "\rSorting nodes: done." << spaces(100) << '\n';
  }

  std::size_t bytes() const noexcept { return get(stb_bytes); }
  void serialize(std::ostream& os) const {
    std::vector<uint8_t> buf(bytes());
    uint64_t written = 0;
    check(stb_serialize(h_, buf.data(), buf.size(), STB_HOST, &written));
    os.write(reinterpret_cast<const char*>(buf.data()), (std::streamsize)written);
  }
  static shared_tree deserialize(std::istream& is) {
    const std::string bytes{std::istreambuf_iterator<char>(is), std::istreambuf_iterator<char>()};
    shared_tree t;
    t.create();
    t.check(stb_deserialize(t.h_, reinterpret_cast<const uint8_t*>(bytes.data()), bytes.size()));
    return t;
  }
  void save(std::filesystem::path path) const {
    std::ofstream file{path, std::ios::binary};
    serialize(file);
  }

  // Sequential decode: the whole sequence is expanded on the GPU once per begin().
  struct iterator {
    std::shared_ptr<std::vector<uint64_t>> data;
    std::size_t at = 0;
    dna operator*() const noexcept { return dna{(unsigned long long)(*data)[at]}; }
    iterator& operator++() {
      ++at;
      return *this;
    }
    bool operator!=(const iterator&) const { return data && at < data->size(); }
  };
  using const_iterator = iterator;
  iterator begin() {
    auto data = std::make_shared<std::vector<uint64_t>>(width());
    check(stb_decode_leaves(h_, 0, data->size(), data->data(), STB_HOST));
    return iterator{data, 0};
  }
  iterator end() { return iterator{}; }

  stb_tree* handle() const { return h_; }
  friend std::ostream& operator<<(std::ostream& os, const shared_tree& tree);

 private:
  struct host_copy {
    std::vector<uint64_t> leaves;
    std::vector<std::vector<uint32_t>> layers;
  };

  void create() {
    const int st = stb_create(&h_, 0, (int)dna::size(), nullptr);
    if (st != STB_OK) {
      std::cerr << "shared_tree_b200: " << stb_status_string(st) << '\n';
      std::exit(1);
    }
  }
  void build(const std::vector<dna>& data) {
    std::vector<uint64_t> raw(data.size());
    for (std::size_t i = 0; i < data.size(); ++i) raw[i] = data[i].to_ullong();
    check(stb_build_from_leaves(h_, raw.data(), raw.size(), STB_HOST));
  }
  void reset() {
    if (h_) stb_destroy(h_);
    h_ = nullptr;
    cache_.reset();
  }
  void check(int st) const { check_on(h_, st); }
  static void check_on(const stb_tree* h, int st) {
    if (st == STB_OK) return;
    // the reference reports bad input on stderr and exits with status 1 (src/dna.cpp:44-47)
    if (st == STB_ERR_UNKNOWN_SYMBOL) std::cerr << stb_last_error(h) << '\n';
    else std::cerr << "shared_tree_b200: " << stb_status_string(st) << ": " << (h ? stb_last_error(h) : "") << '\n';
    std::exit(1);
  }
  template <typename F>
  std::size_t get(F fn) const {
    uint64_t v = 0;
    check(fn(h_, &v));
    return v;
  }
  void cache() const {
    if (cache_) return;
    cache_ = std::make_shared<host_copy>();
    cache_->leaves.resize(leaf_count());
    check(stb_copy_leaves(h_, cache_->leaves.data(), cache_->leaves.size(), STB_HOST));
    for (std::size_t k = 0; k + 1 < depth(); ++k) {
      cache_->layers.emplace_back(2 * node_count(k));
      check(stb_copy_layer(h_, k, cache_->layers.back().data(), node_count(k), STB_HOST));
    }
  }

  stb_tree* h_ = nullptr;
  mutable std::shared_ptr<host_copy> cache_;
};

inline std::ostream& operator<<(std::ostream& os, const shared_tree& tree) {
  tree.cache();
  os << "Leaves (" << tree.cache_->leaves.size() << "):";
  for (auto v : tree.cache_->leaves) os << ' ' << dna{(unsigned long long)v};
  os << '\n';
  for (const auto& layer : tree.cache_->layers) {
    os << "Layer (" << layer.size() / 2 << "):";
    for (std::size_t i = 0; i < layer.size(); i += 2) os << ' ' << node{pointer::from_raw(layer[i]), pointer::from_raw(layer[i + 1])};
    os << '\n';
  }
  return os;
}
