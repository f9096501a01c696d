// Compiled and run by tests/test_host_rng.py (g++, no GPU): the host RNG building blocks against each other.
//   * copies of a ReplayStream share the generated outputs but keep their own cursor, and return the std::mt19937
//     sequence at every position (randomsequence.hh:19-138 semantics);
//   * normal_fill_cached() == successive calls of a fresh BoxMullerNormal, bit for bit, on hit and on miss, for even
//     and odd lengths, and it consumes the same number of raw outputs.
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "rng.hpp"

using namespace pf::host;

static int fails = 0;
#define CHECK(c)                                               \
    do {                                                       \
        if (!(c)) {                                            \
            std::printf("FAILED line %d: %s\n", __LINE__, #c); \
            ++fails;                                           \
        }                                                      \
    } while (0)

int main() {
    // --- shared replay stream
    std::mt19937 ref(5489u);
    std::vector<std::uint32_t> seq(300000);
    for (auto& v : seq) v = ref();
    ReplayStream a;
    ReplayStream b(a), c(a);  // copies made BEFORE anything was drawn
    for (int i = 0; i < 1000; ++i) CHECK(a() == seq[i]);
    CHECK(b.position() == 0 && b() == seq[0]);  // own cursor
    c.set_position(250000);                     // far ahead: slides the shared window
    for (int i = 0; i < 100; ++i) CHECK(c() == seq[250000 + i]);
    a.set_position(249000);                     // still inside the window
    CHECK(a() == seq[249000]);
    b.set_position(5);                          // older than the window: regenerated from the seed
    for (int i = 0; i < 100; ++i) CHECK(b() == seq[5 + i]);
    c.set_position(260000);
    CHECK(c() == seq[260000]);
    ReplayStream d(b);                          // a copy inherits the cursor
    CHECK(d.position() == b.position() && d() == seq[b.position()]);

    // --- cached Box-Muller products
    for (std::size_t n : {1u, 2u, 7u, 16u, 1001u, 1024u}) {
        for (double sigma : {0.1, 1.7}) {
            NoiseCache cache;
            ReplayStream g1, g2, g3;
            g1.set_position(123);
            g2.set_position(123);
            g3.set_position(123);
            std::vector<double> want(n), miss(n), hit(n);
            BoxMullerNormal scalar(0.25, sigma);
            for (std::size_t i = 0; i < n; ++i) want[i] = scalar(g1);
            normal_fill_cached(cache, g2, 0.25, sigma, miss.data(), n);  // computes
            normal_fill_cached(cache, g3, 0.25, sigma, hit.data(), n);   // served from the cache
            CHECK(std::memcmp(want.data(), miss.data(), n * sizeof(double)) == 0);
            CHECK(std::memcmp(want.data(), hit.data(), n * sizeof(double)) == 0);
            CHECK(g1.position() == g2.position() && g1.position() == g3.position());
            CHECK(g1.position() == 123 + NoiseCache::draws(n));
            // another sigma / mean at the same position reuses the same products
            std::vector<double> other(n), other_want(n);
            ReplayStream g4, g5;
            g4.set_position(123);
            g5.set_position(123);
            BoxMullerNormal scalar2(-1.0, 3.0 * sigma);
            for (std::size_t i = 0; i < n; ++i) other_want[i] = scalar2(g4);
            normal_fill_cached(cache, g5, -1.0, 3.0 * sigma, other.data(), n);
            CHECK(std::memcmp(other_want.data(), other.data(), n * sizeof(double)) == 0);
        }
    }
    std::printf(fails ? "host rng check: %d failures\n" : "host rng check: ok\n", fails);
    return fails ? 1 : 0;
}
