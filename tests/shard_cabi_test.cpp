// Two ranks of the sharded build driven through the C ABI alone (include/shared_tree_b200_dist.h): host
// buffers, one std::thread per rank, no CUDA headers.  The gathered stream must equal the single-GPU one.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "shared_tree_b200_dist.h"

#define CHECK(expr)                                                           \
  do {                                                                        \
    const int rc__ = (expr);                                                  \
    if (rc__ != STB_OK) {                                                     \
      std::fprintf(stderr, "%s failed: %s\n", #expr, stb_status_string(rc__)); \
      std::exit(1);                                                           \
    }                                                                         \
  } while (0)

static std::vector<uint8_t> stream_of(stb_tree* t) {
  uint64_t n = 0, written = 0;
  CHECK(stb_bytes(t, &n));
  std::vector<uint8_t> out(n);
  CHECK(stb_serialize(t, out.data(), n, STB_HOST, &written));
  return out;
}

int main() {
  // a pseudo-random ACGT body with a planted repeat, deterministic
  const uint64_t n_bases = 3000000;
  std::string body(n_bases, 'A');
  uint64_t x = 88172645463325252ull;
  for (uint64_t i = 0; i < n_bases; ++i) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    body[i] = "ACGT"[x & 3];
  }
  for (uint64_t i = 0; i < 240000; ++i) body[1200000 + i] = body[12000 + i];

  stb_tree* single = nullptr;
  CHECK(stb_create(&single, 0, 12, nullptr));
  CHECK(stb_build_from_body(single, body.data(), n_bases, STB_HOST));
  const auto want = stream_of(single);

  const int world = 2;
  stb_shard* ranks[2] = {nullptr, nullptr};
  CHECK(stb_shard_create_local(ranks, world, 0, 12));
  stb_tree* gathered = nullptr;
  CHECK(stb_create(&gathered, 0, 12, nullptr));
  std::vector<std::thread> threads;
  for (int r = 0; r < world; ++r)
    threads.emplace_back([&, r] {
      CHECK(stb_shard_set_option(ranks[r], "cut", 4096));
      uint64_t first = 0, count = 0;
      CHECK(stb_shard_range(ranks[r], n_bases, &first, &count));
      const int rc = stb_shard_build_from_body(ranks[r], body.data() + first, n_bases, STB_HOST);
      if (rc != STB_OK) {
        std::fprintf(stderr, "rank %d: %s\n", r, stb_shard_last_error(ranks[r]));
        std::exit(1);
      }
      CHECK(stb_shard_gather(ranks[r], r == 0 ? gathered : nullptr));
    });
  for (auto& t : threads) t.join();
  const auto got = stream_of(gathered);
  if (got != want) {
    std::fprintf(stderr, "streams differ: %zu vs %zu bytes\n", got.size(), want.size());
    return 1;
  }
  uint64_t totals[64], levels = 0;
  CHECK(stb_shard_layer_totals(ranks[0], totals, 64, &levels));
  for (int r = 0; r < world; ++r) stb_shard_destroy(ranks[r]);
  stb_destroy(single);
  stb_destroy(gathered);
  std::printf("shard_cabi_test ok: %zu stream bytes, %llu sharded levels, %llu leaves\n", got.size(), (unsigned long long)levels,
              (unsigned long long)totals[0]);
  return 0;
}
