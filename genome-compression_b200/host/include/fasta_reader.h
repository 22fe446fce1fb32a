// fasta_reader.h — the reference's ingest surface (include/fasta_reader.h) on top of the C ABI.
// The file is read once; header/newline removal and 4-bit packing run on the GPU
// (stb_pack_fasta), reproducing src/fasta_reader.cpp:40-68 (see csrc/ingest.cu).
#pragma once

#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "dna.h"
#include "shared_tree_b200.h"

class fasta_reader {
 public:
  using value_type = dna;

  fasta_reader(std::filesystem::path path, std::size_t buffer_size = (1 << 22)) : path_(path), buffer_size_(buffer_size) {
    std::ifstream file{path, std::ios::binary};
    if (!file.is_open()) {
      std::cerr << "Unable to open file, aborting...\n";  // src/fasta_reader.cpp:16
      std::exit(1);
    }
    text_.assign(std::istreambuf_iterator<char>(file), std::istreambuf_iterator<char>());
  }
  fasta_reader(const fasta_reader&) = delete;
  fasta_reader(fasta_reader&&) = delete;

  bool eof() const { return packed_ && next_ >= leaves_.size(); }
  std::size_t size() const { return text_.size(); }
  std::size_t buffers() const { return size() / (buffer_size_ * dna::size()); }
  const std::string& text() const { return text_; }
  bool untouched() const { return next_ == 0; }

  // Hands out the next buffer_size leaves; false once nothing is left.
  bool read_into(std::vector<dna>& out) {
    pack();
    out.clear();
    const std::size_t stop = std::min(leaves_.size(), next_ + buffer_size_);
    out.reserve(stop - next_);
    for (; next_ < stop; ++next_) out.emplace_back((unsigned long long)leaves_[next_]);
    return !out.empty();
  }

 private:
  void pack() {
    if (packed_) return;
    packed_ = true;
    stb_tree* h = nullptr;
    int st = stb_create(&h, 0, (int)dna::size(), nullptr);
    if (st == STB_OK) {
      uint64_t n = 0;
      st = stb_pack_fasta(h, text_.data(), text_.size(), STB_HOST, nullptr, 0, STB_HOST, &n);
      if (st == STB_OK) {
        leaves_.resize(n);
        st = stb_pack_fasta(h, text_.data(), text_.size(), STB_HOST, leaves_.data(), n, STB_HOST, &n);
      }
    }
    if (st != STB_OK) {
      std::cerr << (h && st == STB_ERR_UNKNOWN_SYMBOL ? stb_last_error(h) : stb_status_string(st)) << '\n';
      std::exit(1);  // src/dna.cpp:44-47
    }
    stb_destroy(h);
  }

  std::filesystem::path path_;
  std::size_t buffer_size_;
  std::string text_;
  std::vector<uint64_t> leaves_;
  std::size_t next_ = 0;
  bool packed_ = false;
};

// src/fasta_reader.cpp:108-122
inline std::vector<dna> read_genome(const std::filesystem::path path) {
  if (!std::filesystem::is_regular_file(path)) {
    std::cerr << "Non-existent path, aborting...\n";
    std::exit(1);
  }
  std::vector<dna> all, buffer;
  fasta_reader file{path};
  while (file.read_into(buffer)) all.insert(all.end(), buffer.begin(), buffer.end());
  return all;
}
