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
#include <memory>
#include <random>
#include <vector>

namespace pf { namespace host {

using Mt19937 = std::mt19937;  // == boost::random::mt19937, default seed 5489
using Mt11213b = std::mersenne_twister_engine<std::uint32_t, 32, 351, 175, 19, 0xccab8ee7u, 11, 0xffffffffu, 7,
                                              0x31b6ab00u, 15, 0xffe50000u, 17, 1812433253u>;

// A generator whose output sequence can be revisited by index.  Copies SHARE the outputs generated so far (one
// Mersenne Twister per family of copies) and keep their own cursor: every virtual worker of the frontier communicator
// owns a copy of the executor, and with private streams each of the 64 workers regenerated the whole sequence up to
// the position of its next job -- for a 1000-dimensional model more host time than the proposal itself.
class ReplayStream {
  public:
    using result_type = std::uint32_t;
    explicit ReplayStream(std::uint32_t seed = 5489u) : s_(std::make_shared<Shared>(seed)) {}

    std::size_t position() const { return pos_; }

    void set_position(std::size_t p) {
        Shared& s = *s_;
        if (p < s.first) {  // older than the window: regenerate from the seed
            s.gen.seed(s.seed);
            s.window.clear();
            s.first = 0;
        }
        while (s.first + s.window.size() < p) produce(s);
        pos_ = p;
    }

    result_type operator()() {
        Shared& s = *s_;
        if (pos_ < s.first) set_position(pos_);  // another copy slid the window past this cursor
        if (pos_ == s.first + s.window.size()) produce(s);
        return s.window[pos_++ - s.first];
    }

    static constexpr result_type min() { return Mt19937::min(); }
    static constexpr result_type max() { return Mt19937::max(); }

  private:
    static constexpr std::size_t kKeep = 65536;  // outputs kept for cheap rewinds (32 proposals of a 1000-dimensional model)
    struct Shared {
        explicit Shared(std::uint32_t sd) : seed(sd), gen(sd) {}
        std::uint32_t seed;
        Mt19937 gen;
        std::vector<result_type> window;
        std::size_t first = 0;  // stream index of window[0]
    };
    static void produce(Shared& s) {
        if (s.window.size() >= 2 * kKeep) {
            s.window.erase(s.window.begin(), s.window.begin() + kKeep);
            s.first += kKeep;
        }
        s.window.push_back(s.gen());
    }
    std::shared_ptr<Shared> s_;
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

// Box-Muller products shared between the nodes of a speculative tree.
//
// A node draws its proposal at the stream position where its parent's proposal ended, so BOTH children of a node --
// and, as long as every proposal consumes the same number of draws, ALL nodes of one depth -- start at the same
// position and see the same uniforms.  What differs between them is the state they perturb and (for the adaptive
// samplers) the proposal scale.  The expensive half of normal_distribution -- rho cos(2 pi r1), rho sin(2 pi r1)
// with rho = sqrt(-2 log(1 - r2)) -- therefore depends on the position alone: a full depth-6 frontier of a
// 1024-dimensional model needs 7 sets of them, not 64.  A value is still formed as (rho trig) * sigma + mean, the
// operation order of BoxMullerNormal, so proposals are bit-identical with or without the cache.
class NoiseCache {
  public:
    // the n products for a FRESH normal_distribution starting at stream position `pos`, or nullptr
    const std::vector<double>* find(std::size_t pos, std::size_t n) const {
        for (const Entry& e : ring_)
            if (e.n == n && e.pos == pos && !e.w.empty()) return &e.w;
        return nullptr;
    }
    // draws the 2 * ceil(n / 2) uniforms from `g` (which must stand at `pos`) and keeps the products
    template <class Gen>
    const std::vector<double>& make(Gen& g, std::size_t pos, std::size_t n) {
        Entry& e = ring_[next_];
        next_ = (next_ + 1) % kEntries;
        e.pos = pos;
        e.n = n;
        e.w.resize(n);
        for (std::size_t i = 0; i < n; i += 2) {
            const double r1 = uniform01(g);
            const double r2 = uniform01(g);
            const double rho = std::sqrt(-2.0 * std::log(1.0 - r2));
            double sn, cs;
            ::sincos(2.0 * 3.14159265358979323846 * r1, &sn, &cs);
            e.w[i] = rho * cs;
            if (i + 1 < n) e.w[i + 1] = rho * sn;
        }
        return e.w;
    }
    static std::size_t draws(std::size_t n) { return 2 * ((n + 1) / 2); }  // raw outputs consumed by n variates

  private:
    static constexpr std::size_t kEntries = 32;
    struct Entry {
        std::size_t pos = 0, n = 0;
        std::vector<double> w;
    };
    Entry ring_[kEntries];
    std::size_t next_ = 0;
};

// n variates of a fresh normal_distribution(mean, sigma) on `g`, through the cache: out[i] = w[i] * sigma + mean
template <class Gen>
inline void normal_fill_cached(NoiseCache& cache, Gen& g, double mean, double sigma, double* out, std::size_t n) {
    const std::size_t pos = g.position();
    const std::vector<double>* w = cache.find(pos, n);
    if (w)
        g.set_position(pos + NoiseCache::draws(n));  // consume what the draws would have consumed
    else
        w = &cache.make(g, pos, n);
    for (std::size_t i = 0; i < n; ++i) out[i] = (*w)[i] * sigma + mean;
}

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
