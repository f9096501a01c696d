// host/rng.hpp -- the random streams the chain depends on.
//
// Chain identity with elaine84/fetching requires the same draws in the same order:
//   * ReplayStream: a mt19937 stream addressable by POSITION (number of raw 32-bit outputs
//     consumed), so a proposal can be re-derived anywhere in the speculative tree by rewinding to
//     the position recorded for its parent.  Plays the role of RandomSequence<>
//     (randomsequence.hh:19-138; pinned by test_random_sequence, testmaster.cc:295-332).
//   * the Boost.Random distributions the reference applies to those streams, restated from
//     their published algorithms (Boost 1.47-1.63: uniform_01 = one 32-bit draw * 2^-32;
//     normal_distribution = Box-Muller with a cached second variate; uniform_int = bucket
//     rejection).  Call sites: master.hh:39-42, master.cc:438-445, normalsampler.hh:73-75,115-116,
//     boltzmannsampler.hh:51-52, normal.cc:41-46.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <random>
#include <vector>

namespace pf { namespace host {

using Mt19937 = std::mt19937;  // == boost::random::mt19937, default seed 5489
using Mt11213b = std::mersenne_twister_engine<std::uint32_t, 32, 351, 175, 19, 0xccab8ee7u, 11, 0xffffffffu, 7,
                                              0x31b6ab00u, 15, 0xffe50000u, 17, 1812433253u>;

// A generator whose output sequence can be revisited by index.
class ReplayStream {
  public:
    using result_type = std::uint32_t;
    explicit ReplayStream(std::uint32_t seed = 5489u) : seed_(seed), gen_(seed) {}

    std::size_t position() const { return pos_; }

    void set_position(std::size_t p) {
        if (p < first_) {  // older than the window: regenerate from the seed
            gen_.seed(seed_);
            window_.clear();
            first_ = 0;
        }
        while (first_ + window_.size() < p) produce();
        pos_ = p;
    }

    result_type operator()() {
        if (pos_ == first_ + window_.size()) produce();
        return window_[pos_++ - first_];
    }

    static constexpr result_type min() { return Mt19937::min(); }
    static constexpr result_type max() { return Mt19937::max(); }

  private:
    static constexpr std::size_t kKeep = 65536;  // outputs kept for cheap rewinds (32 proposals of a 1000-dimensional model)
    void produce() {
        if (window_.size() >= 2 * kKeep) {
            window_.erase(window_.begin(), window_.begin() + kKeep);
            first_ += kKeep;
        }
        window_.push_back(gen_());
    }
    std::uint32_t seed_;
    Mt19937 gen_;
    std::vector<result_type> window_;
    std::size_t first_ = 0;  // stream index of window_[0]
    std::size_t pos_ = 0;
};

// uniform_01<double> on a 32-bit engine: one draw, scaled by 2^-32
template <class Gen>
inline double uniform01(Gen& g) {
    for (;;) {
        const double r = static_cast<double>(static_cast<std::uint32_t>(g() - Gen::min())) * (1.0 / 4294967296.0);
        if (r < 1.0) return r;
    }
}

// normal_distribution<double>(mean, sigma): Box-Muller, second variate cached in the OBJECT
class BoxMullerNormal {
  public:
    explicit BoxMullerNormal(double mean = 0.0, double sigma = 1.0) : mean_(mean), sigma_(sigma) {}
    template <class Gen>
    double operator()(Gen& g) {
        if (!have_) {
            r1_ = uniform01(g);
            const double r2 = uniform01(g);
            rho_ = std::sqrt(-2.0 * std::log(1.0 - r2));
            have_ = true;
            // cos now, sin for the cached variate: one sincos() (glibc computes both with the kernels of sin() and
            // cos(), so the values are the ones two separate calls return -- the chain-parity tests compare the
            // thetas bit for bit against the reference build, which calls them separately).  Proposals of
            // 1000-dimensional models are the host's hot spot: this is a quarter of their cost.
            ::sincos(2.0 * 3.14159265358979323846 * r1_, &sin_, &cos_);
            return rho_ * cos_ * sigma_ + mean_;
        }
        have_ = false;
        return rho_ * sin_ * sigma_ + mean_;
    }

  private:
    double mean_, sigma_, r1_ = 0.0, rho_ = 0.0, sin_ = 0.0, cos_ = 1.0;
    bool have_ = false;
};

// uniform_int_distribution<int>(0, range) on a 32-bit engine
template <class Gen>
inline std::uint32_t uniform_int(Gen& g, std::uint32_t range) {
    if (range == 0) return 0;
    if (range == 0xffffffffu) return static_cast<std::uint32_t>(g() - Gen::min());
    std::uint32_t bucket = 0xffffffffu / (range + 1u);
    if (0xffffffffu % (range + 1u) == range) ++bucket;
    for (;;) {
        const std::uint32_t r = static_cast<std::uint32_t>(g() - Gen::min()) / bucket;
        if (r <= range) return r;
    }
}

// numpy's legacy random_sample(): 53-bit double from two consecutive 32-bit draws
template <class Gen>
inline double numpy_random_sample(Gen& g) {
    const std::uint32_t a = static_cast<std::uint32_t>(g()) >> 5, b = static_cast<std::uint32_t>(g()) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

} }  // namespace pf::host
