// dna.h — host value type for one leaf (the reference's `dna`, include/dna.h:39-75), with the
// same public surface, as a thin wrapper over the 64-bit word the CUDA path uses.  The hot path
// never calls these; they exist so callers written against the reference keep compiling.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <string_view>
#include <tuple>

#include "utility.h"

// 4-bit nucleotide codes (include/dna.h:20-32): complement = bit-reversed nibble
enum class nac : char {
  A = 0x1, T = 0x8, C = 0x2, G = 0x4, R = 0x3, Y = 0xC, K = 0x7, M = 0xE,
  B = 0x5, V = 0xA, D = 0xB, H = 0xD, S = 0x0, W = 0x9, N = 0x6, Indeterminate = 0xF
};

class dna {
 public:
  dna() = default;
  dna(unsigned long long value) noexcept : word_(value) {}
  dna(const std::string_view strand) : word_(0) {
    for (std::size_t i = 0; i < length_ && i < strand.size(); ++i) word_ |= std::uint64_t(encode(strand[i])) << (4 * i);
  }

  static std::size_t size() noexcept { return length_; }
  static std::size_t size(std::size_t n) noexcept { return length_ = n; }
  static std::size_t bytes() noexcept { return (length_ + 1) / 2; }

  // test helper of the reference (src/dna.cpp:90-95): low dna::size() BITS of two rand() calls
  static dna random(unsigned seed = 0) {
    std::srand(seed);
    const auto r = static_cast<unsigned long long>(rand() | (std::uint64_t(rand()) << 32));
    return dna{r & ((1u << size()) - 1)};
  }

  dna transposed() const noexcept {
    std::uint64_t v = word_;
    v = ((v >> 1) & 0x5555555555555555ull) | ((v & 0x5555555555555555ull) << 1);
    v = ((v >> 2) & 0x3333333333333333ull) | ((v & 0x3333333333333333ull) << 2);
    return dna{v};
  }
  dna mirrored() const noexcept {
    std::uint64_t out = 0;
    for (std::size_t i = 0; i < length_; ++i) out |= ((word_ >> (4 * (length_ - 1 - i))) & 0xf) << (4 * i);
    return dna{out};
  }
  dna inverted() const noexcept { return transposed().mirrored(); }
  bool invariant() const noexcept { return *this == mirrored(); }
  // minimum over (value, mirror, transpose), earliest wins (src/dna.cpp:135-143)
  std::tuple<dna, bool, bool, bool> canonical() const noexcept {
    const bool inv = invariant();
    return variadic_min(std::tuple{*this, false, false, inv}, std::tuple{transposed(), false, true, inv},
                        std::tuple{mirrored(), true, false, inv}, std::tuple{inverted(), true, true, inv});
  }

  void serialize(std::ostream& os) const { binary_write(os, word_, bytes()); }
  static dna deserialize(std::istream& is) {
    std::uint64_t v = 0;
    binary_read(is, v, bytes());
    return dna{v};
  }

  nac code(std::size_t i) const { return static_cast<nac>((word_ >> (4 * i)) & 0xf); }
  char nucleotide(std::size_t i) const { return "SACRGBNKTWVDYHM-"[(word_ >> (4 * i)) & 0xf]; }

  bool operator==(const dna& o) const noexcept { return word_ == o.word_; }
  bool operator!=(const dna& o) const noexcept { return word_ != o.word_; }
  bool operator<(const dna& o) const noexcept { return word_ < o.word_; }
  operator std::uint64_t() const noexcept { return word_; }
  std::uint64_t to_ullong() const noexcept { return word_; }

 private:
  // src/dna.cpp:25-49: case-insensitive, anything else aborts the program
  static unsigned encode(char ch) {
    const int c = (ch >= 'a' && ch <= 'z') ? ch - 32 : ch;
    switch (c) {
      case 'A': return 0x1; case 'C': return 0x2; case 'G': return 0x4; case 'T': return 0x8;
      case 'R': return 0x3; case 'Y': return 0xC; case 'K': return 0x7; case 'M': return 0xE;
      case 'S': return 0x0; case 'W': return 0x9; case 'B': return 0x5; case 'D': return 0xB;
      case 'H': return 0xD; case 'V': return 0xA; case 'N': return 0x6; case '-': return 0xF;
    }
    std::cerr << "Encountered unknown symbol: " << c << " (ASCII code " << c << ")\n";
    std::exit(1);
  }

  std::uint64_t word_ = 0;
  inline static std::size_t length_ = 12;
};

inline std::ostream& operator<<(std::ostream& os, const dna& strand) {
  for (std::size_t i = 0; i < dna::size(); ++i) os << strand.nucleotide(i);
  return os;
}

namespace std {
template <>
struct hash<dna> {
  std::size_t operator()(const dna& d) const noexcept { return std::hash<std::uint64_t>()(d.to_ullong()); }
};
}  // namespace std
