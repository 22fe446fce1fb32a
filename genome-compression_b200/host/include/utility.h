// utility.h — small host helpers with the reference's names (include/utility.h) so that
// its compress.cpp and tests/test.cpp build against this shim unchanged.  Written for this
// repository; only the behaviour that results depend on is mirrored:
//   * variadic_min keeps the EARLIEST of equal candidates (include/utility.h:164-170)
//   * binary_write / binary_read are big-endian (include/utility.h:178-194)
#pragma once

#include <array>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <iterator>
#include <sstream>
#include <string>
#include <string_view>
#include <type_traits>
#include <utility>

// (1,2,3,4,5) -> f(1,2) f(3,4) g(5)
template <typename Range, typename Pair, typename Single>
void foreach_pair(Range&& range, Pair on_pair, Single on_single) {
  auto it = std::begin(range);
  const auto last = std::end(range);
  while (it != last) {
    auto&& first = *it++;
    if (it == last) {
      on_single(first);
      return;
    }
    on_pair(first, *it++);
  }
}

namespace detail {
inline std::size_t hash() noexcept { return 0; }
template <typename Head, typename... Tail>
std::size_t hash(const Head& head, const Tail&... tail) noexcept {
  return ((std::size_t{1} << (sizeof...(tail) + 1)) + 1) * static_cast<std::size_t>(head) + hash(tail...);
}
}  // namespace detail

template <typename T>
constexpr decltype(auto) variadic_min(T&& only) noexcept { return std::forward<T>(only); }
template <typename A, typename B, typename... Rest>
constexpr decltype(auto) variadic_min(A&& a, B&& b, Rest&&... rest) noexcept {
  if (b < a) return variadic_min(b, std::forward<Rest>(rest)...);
  return variadic_min(a, std::forward<Rest>(rest)...);
}

// range of consecutive sub-ranges of at most `size` elements
template <typename Range>
class chunk_view {
  using iter = decltype(std::begin(std::declval<Range&>()));
  Range& range_;
  std::size_t size_;

 public:
  struct chunk {
    iter first, last;
    iter begin() const { return first; }
    iter end() const { return last; }
    std::size_t size() const { return (std::size_t)std::distance(first, last); }
  };
  struct iterator {
    iter cur, last;
    std::size_t size;
    chunk operator*() const {
      iter stop = cur;
      for (std::size_t i = 0; i < size && stop != last; ++i) ++stop;
      return chunk{cur, stop};
    }
    iterator& operator++() {
      for (std::size_t i = 0; i < size && cur != last; ++i) ++cur;
      return *this;
    }
    bool operator!=(const iterator& o) const { return cur != o.cur; }
  };
  chunk_view(Range& r, std::size_t size) : range_(r), size_(size) {}
  iterator begin() { return {std::begin(range_), std::end(range_), size_}; }
  iterator end() { return {std::end(range_), std::end(range_), size_}; }
};
template <typename Range>
chunk_view<std::remove_reference_t<Range>> chunks(Range&& range, std::size_t size) { return {range, size}; }

template <typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>>
void binary_write(std::ostream& os, T value, std::size_t bytes = sizeof(T)) {
  for (std::size_t b = bytes; b-- > 0;) os.put(static_cast<char>((value >> (8 * b)) & 0xff));
}
template <typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>>
void binary_read(std::istream& is, T& value, std::size_t bytes = sizeof(T)) {
  value = 0;
  for (std::size_t b = bytes; b-- > 0;) {
    char c = 0;
    is.get(c);
    value |= static_cast<T>(static_cast<unsigned char>(c)) << static_cast<T>(8 * b);
  }
}

template <typename T>
std::string bytes_to_string(T bytes) {
  static constexpr const char* unit[] = {"B", "KB", "MB", "GB", "TB", "PB", "EB"};
  double v = static_cast<double>(bytes);
  std::size_t u = 0;
  for (; v >= 1000.0 && u + 1 < std::size(unit); ++u) v /= 1000.0;
  std::ostringstream out;
  out.precision(3);
  out << v << ' ' << unit[u];
  return out.str();
}

inline std::string progress_bar(std::string_view name, unsigned current, unsigned end) {
  const double done = end ? double(current) / double(end) : 1.0;
  std::ostringstream out;
  out << '\r' << name << ": [" << std::string(unsigned(done * 60), '#') << std::string(60 - unsigned(done * 60), ' ') << "] "
      << unsigned(done * 100) << '%';
  return out.str();
}
inline std::string spaces(unsigned n) { return std::string(n, ' '); }
