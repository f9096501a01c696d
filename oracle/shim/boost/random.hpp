/* oracle/shim/boost/random.hpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Restatement of the pieces of Boost.Random the reference calls.  Boost is an
 * UNVENDORED, UNPINNED third-party dependency of the reference (GNUmakefile:10 links
 * the system package; README.md:25 says `port install boost +openmpi`), and its
 * sources are not in this image.  What is restated here is the published algorithm
 * of Boost 1.47 - 1.63 (the releases contemporary with the reference, 2012-2014):
 *
 *   mt19937, mt11213b      the standard Mersenne-twister parameterisations
 *                          (boost/random/mersenne_twister.hpp), default seed 5489
 *   uniform_01<double>     (eng() - eng.min()) * 1/(max-min+1); one 32-bit draw,
 *                          retry if the product rounds to 1 (uniform_01.hpp)
 *   normal_distribution    polar-free Box-Muller with a cached second variate:
 *                          r1,r2 = two uniform_01; rho = sqrt(-2 log(1-r2));
 *                          first call rho*cos(2 pi r1), second rho*sin(2 pi r1)
 *                          (normal_distribution.hpp before the 1.64 ziggurat)
 *   uniform_int_distribution  bucket rejection on a 32-bit engine
 *                          (uniform_int_distribution.hpp, generate_uniform_int)
 *   uniform_real / lognormal  only named by samplers/python.cc, never instantiated
 *                          by anything the oracle builds; provided for completeness.
 *
 * PARITY NOTE: the only reference test that touches this boundary is
 * test_random_sequence (testmaster.cc:295-332), which pins the raw mt19937 stream.
 * No reference test pins the distributions, so proposal / log_u *values* are pinned
 * against this restatement (and SURVEY.md section 8c.3's probe goldens), not against
 * an arbitrary real Boost build.  Call sites: master.hh:39-42,129,131; master.cc:438,445;
 * normalsampler.hh:73-75,115-116; boltzmannsampler.hh:51-52; normal.cc:41-46;
 * testmaster.cc:297,487-492; randomsequence.hh:19. */
#ifndef PF_SHIM_BOOST_RANDOM_HPP
#define PF_SHIM_BOOST_RANDOM_HPP
#include <cmath>
#include <cstdint>
#include <limits>
#include <random>

namespace boost { namespace random {

typedef std::mt19937 mt19937;
typedef std::mersenne_twister_engine<std::uint32_t, 32, 351, 175, 19, 0xccab8ee7u, 11,
                                     0xffffffffu, 7, 0x31b6ab00u, 15, 0xffe50000u, 17,
                                     1812433253u> mt11213b;

template <typename RealType = double>
class uniform_01 {
  public:
    typedef RealType result_type;
    template <typename Engine> result_type operator()(Engine& eng) {
        for (;;) {
            typedef typename Engine::result_type base_result;
            result_type factor = 1 / (result_type((eng.max)() - (eng.min)()) + 1);
            result_type result = result_type(base_result(eng() - (eng.min)())) * factor;
            if (result < result_type(1))
                return result;
        }
    }
};

template <typename RealType = double>
class normal_distribution {
  public:
    typedef RealType result_type;
    explicit normal_distribution(RealType mean = 0, RealType sigma = 1)
        : mean_(mean), sigma_(sigma), r1_(0), r2_(0), cached_rho_(0), valid_(false) {}
    template <typename Engine> result_type operator()(Engine& eng) {
        using std::sqrt; using std::log; using std::sin; using std::cos;
        if (!valid_) {
            r1_ = uniform_01<RealType>()(eng);
            r2_ = uniform_01<RealType>()(eng);
            cached_rho_ = sqrt(-result_type(2) * log(result_type(1) - r2_));
            valid_ = true;
        } else
            valid_ = false;
        const result_type pi = result_type(3.14159265358979323846);
        return cached_rho_ * (valid_ ? cos(result_type(2) * pi * r1_)
                                     : sin(result_type(2) * pi * r1_)) * sigma_ + mean_;
    }
  private:
    RealType mean_, sigma_, r1_, r2_, cached_rho_;
    bool valid_;
};

template <typename IntType = int>
class uniform_int_distribution {
  public:
    typedef IntType result_type;
    explicit uniform_int_distribution(IntType min_arg = 0,
                                      IntType max_arg = (std::numeric_limits<IntType>::max)())
        : min_(min_arg), max_(max_arg) {}
    template <typename Engine> result_type operator()(Engine& eng) const {
        typedef typename std::make_unsigned<IntType>::type range_type;
        typedef typename Engine::result_type base_result;
        const range_type range = range_type(max_) - range_type(min_);
        const base_result bmin = (eng.min)();
        const base_result brange = (eng.max)() - (eng.min)();
        if (range == 0)
            return min_;
        if (brange == range)
            return IntType(range_type(eng() - bmin) + range_type(min_));
        /* brange > range for every call site (32-bit engine, int range) */
        base_result bucket_size = brange / (base_result(range) + 1);
        if (brange % (base_result(range) + 1) == base_result(range))
            ++bucket_size;
        for (;;) {
            base_result result = base_result(eng() - bmin) / bucket_size;
            if (result <= base_result(range))
                return IntType(range_type(result) + range_type(min_));
        }
    }
  private:
    IntType min_, max_;
};

template <typename RealType = double>
class uniform_real_distribution {
  public:
    typedef RealType result_type;
    explicit uniform_real_distribution(RealType a = 0, RealType b = 1) : a_(a), b_(b) {}
    template <typename Engine> result_type operator()(Engine& eng) const {
        return uniform_01<RealType>()(eng) * (b_ - a_) + a_;
    }
  private:
    RealType a_, b_;
};

template <typename RealType = double>
class lognormal_distribution {
  public:
    typedef RealType result_type;
    explicit lognormal_distribution(RealType m = 0, RealType s = 1) : n_(m, s) {}
    template <typename Engine> result_type operator()(Engine& eng) { return std::exp(n_(eng)); }
  private:
    normal_distribution<RealType> n_;
};

} /* namespace random */
using random::mt19937;
using random::mt11213b;
} /* namespace boost */
#endif
