// host/executors.hpp -- the host half of the Executor concept for every sampler.
//
// Executor concept (worker.hh:28-36,150): std::string load(const JobInfo&);
// void prepare(); double log_prior(); size_t random_position() const -- proposal generation,
// prior and random-stream bookkeeping, all cheap and all on the host.  The expensive half,
// evaluate()/evaluate_many(), is NOT here: the frontier communicator sends the thetas of a whole
// scheduling burst to the GPU engine in one call (include/pfetch.h).
//
// theta is carried as the raw bytes of its doubles (protocol.hpp) in the layout the engine expects.
#pragma once
#include <cmath>
#include <cstddef>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "protocol.hpp"
#include "rng.hpp"
#include "stride_index.hpp"

namespace pf { namespace host {

// ---------------------------------------------------------------------------------------------
// Gaussian target: NormalSampler<D> / NormalTheta<D> (normalsampler.hh:28-165, normaltheta.hh)
// ---------------------------------------------------------------------------------------------
template <int D>
class NormalExecutor {
  public:
    static constexpr int theta_len = D + D * (D + 1) / 2;
    explicit NormalExecutor(std::size_t data_size) : data_size_(data_size) {}

    // normalsampler.hh:66-79.  The initial theta is mean 0 / identity covariance
    // (normaltheta.hh:55-62); a proposal adds N(0, 0.1) to each mean from a FRESH distribution.
    std::string load(const JobInfo& job) {
        if (job.need_initial()) {
            for (double& v : theta_) v = 0.0;
            for (int i = 0; i < D; ++i) theta_[D + i * (i + 1) / 2 + i] = 1.0;
        } else {
            engine_.set_position(job.random_position);
            load_doubles(job.theta, theta_, theta_len);
            if (job.need_proposal()) {
                BoxMullerNormal step(0, 0.1);
                for (int i = 0; i < D; ++i) theta_[i] = theta_[i] + step(engine_);
            }
        }
        return save_doubles(theta_, theta_len);
    }
    // normalsampler.hh:115-119: the per-node visiting order; the two draws must be consumed even
    // when the whole data set is scored in one pass, because positions are recorded around them.
    void prepare() {
        const std::uint32_t range = (std::uint32_t) (data_size_ - 1);
        const std::size_t stride = uniform_int(engine_, range);
        const std::size_t offset = uniform_int(engine_, range);
        order_ = StrideIndex(data_size_, StrideIndex::valid_stride(data_size_, stride), offset);
    }
    double log_prior() { return 0.0; }
    std::size_t random_position() const { return engine_.position(); }
    const StrideIndex& order() const { return order_; }
    std::size_t data_size() const { return data_size_; }

    static std::string unparse(const std::string& theta) {  // normaltheta.hh:152-167
        double v[theta_len];
        load_doubles(theta, v, theta_len);
        std::ostringstream s;
        const char* sep = "[";
        for (int i = 0; i < D; ++i) {
            s << sep << v[i];
            sep = ", ";
        }
        sep = "; ";
        for (int i = 0; i < D; ++i)
            for (int j = 0; j <= i; ++j) {
                s << sep << v[D + i * (i + 1) / 2 + j];
                sep = ", ";
            }
        s << "]";
        return s.str();
    }

  private:
    std::size_t data_size_;
    ReplayStream engine_;
    StrideIndex order_;
    double theta_[theta_len];
};

// ---------------------------------------------------------------------------------------------
// scalar Boltzmann (boltzmannsampler.hh:18-82): theta >= 0, log p = -theta / T per item
// ---------------------------------------------------------------------------------------------
class BoltzmannScalarExecutor {
  public:
    static constexpr int theta_len = 1;
    explicit BoltzmannScalarExecutor(std::size_t data_size) : data_size_(data_size) {}
    std::string load(const JobInfo& job) {  // boltzmannsampler.hh:40-58
        if (job.need_initial())
            theta_ = 0.25;
        else {
            load_doubles(job.theta, &theta_, 1);
            if (job.need_proposal()) {
                engine_.set_position(job.random_position);
                BoxMullerNormal step(theta_, 0.1);
                const double tentative = step(engine_);
                if (tentative >= 0) theta_ = tentative;
            }
        }
        return save_doubles(&theta_, 1);
    }
    void prepare() {}
    double log_prior() { return 0.0; }
    std::size_t random_position() const { return engine_.position(); }
    std::size_t data_size() const { return data_size_; }
    static std::string unparse(const std::string& theta) {
        double v = 0;
        load_doubles(theta, &v, 1);
        std::ostringstream s;
        s << v;
        return s.str();
    }

  private:
    std::size_t data_size_;
    double theta_ = 0.25;
    ReplayStream engine_;
};

// ---------------------------------------------------------------------------------------------
// shared by the models the reference ran through its Python bridge (python.cc:141-192):
// load() ALWAYS rewinds the stream first (python.cc:142); theta printed as a nested list
// ---------------------------------------------------------------------------------------------
inline std::string unparse_list(const double* v, std::size_t n, std::size_t row) {
    std::ostringstream s;
    s.precision(17);
    s << "[";
    for (std::size_t i = 0; i < n; ++i) {
        if (row && i % row == 0) s << (i ? "], [" : "[");
        else if (i) s << ", ";
        s << v[i];
    }
    if (row) s << "]";
    s << "]";
    return s.str();
}

// Gaussian mixture (pysamplers/pymixture.py).  first_proposal reproduces the reference exactly
// (np.random.seed(seed); 4 * np.random.random((K, d)) - 2, pymixture.py:69-73: numpy's legacy
// generator is MT19937 + 53-bit doubles).  next_proposal keeps the reference's RULE -- permute
// the component labels, add N(0, sqrt(exp(scale))) to every coordinate (pymixture.py:81-84) --
// but draws from the replayable stream itself instead of reseeding numpy's global generator with
// int(rand.random() * 10**12) (pymixture.py:77-78), which is unpinned: a
// documented divergence; see DESIGN.md.
class MixtureExecutor {
  public:
    MixtureExecutor(std::size_t data_size, int K, int d, std::uint32_t seed = 0)
        : data_size_(data_size), K_(K), d_(d), seed_(seed), theta_((std::size_t) K * d) {}
    int theta_len() const { return K_ * d_; }

    std::string load(const JobInfo& job) {
        engine_.set_position(job.random_position);
        if (job.need_initial()) {
            Mt19937 np(seed_);
            for (double& v : theta_) v = 4.0 * numpy_random_sample(np) - 4.0 / 2.0;
        } else {
            load_doubles(job.theta, theta_.data(), theta_.size());
            if (job.need_proposal()) {
                std::vector<int> perm(K_);
                for (int i = 0; i < K_; ++i) perm[i] = i;
                for (int i = K_ - 1; i > 0; --i) std::swap(perm[i], perm[uniform_int(engine_, (std::uint32_t) i)]);
                BoxMullerNormal jump(0.0, std::sqrt(std::exp(job.scale)));
                std::vector<double> next(theta_.size());
                for (int k = 0; k < K_; ++k)
                    for (int j = 0; j < d_; ++j) next[k * d_ + j] = theta_[perm[k] * d_ + j] + jump(engine_);
                theta_.swap(next);
            }
        }
        return save_doubles(theta_.data(), theta_.size());
    }
    void prepare() {}                      // the model has no `prepare` (python.cc:159-166)
    double log_prior() { return 0.0; }     // ... and no `log_prior` (python.cc:168-176)
    std::size_t random_position() const { return engine_.position(); }
    std::size_t data_size() const { return data_size_; }
    std::string unparse(const std::string& theta) const {
        std::vector<double> v((std::size_t) K_ * d_);
        load_doubles(theta, v.data(), v.size());
        return unparse_list(v.data(), v.size(), d_);
    }

  private:
    std::size_t data_size_;
    int K_, d_;
    std::uint32_t seed_;
    std::vector<double> theta_;
    ReplayStream engine_;
};

// Bayesian lasso (pysamplers/pyblasso.py): theta = (sigma, beta).  log_prior is the reference's
// (pyblasso.py:123-133); next_proposal keeps its rule -- add N(0, sqrt(exp(scale))) to every
// coordinate, redraw until sigma > 0 (pyblasso.py:117-121) -- on the replayable stream (same
// documented divergence as the mixture).  The reference's initial theta is a 57-vector fitted to
// its own (absent) OPV data (pyblasso.py:83-102); here it is supplied by the caller, default
// sigma = 1, beta = 0.
class BLassoExecutor {
  public:
    BLassoExecutor(std::size_t data_size, int p, double tau = 5.0, const double* initial = nullptr)
        : data_size_(data_size), p_(p), tau_(tau), theta_((std::size_t) p + 1, 0.0), initial_((std::size_t) p + 1, 0.0) {
        initial_[0] = 1.0;
        if (initial)
            for (int i = 0; i <= p; ++i) initial_[i] = initial[i];
    }
    int theta_len() const { return p_ + 1; }

    std::string load(const JobInfo& job) {
        engine_.set_position(job.random_position);
        if (job.need_initial())
            theta_ = initial_;
        else {
            load_doubles(job.theta, theta_.data(), theta_.size());
            if (job.need_proposal()) {
                // first attempt through the shared cache (the nodes of a depth draw at the same position, rng.hpp); a
                // redraw (sigma <= 0, rare) continues the SAME distribution object, whose cached second variate of the
                // last pair comes first when the vector length is odd: replay the attempt on a scalar distribution
                const double sigma = std::sqrt(std::exp(job.scale));
                const std::size_t n = theta_.size(), start = engine_.position();
                std::vector<double> next(n), z(n);
                normal_fill_cached(*noise_, engine_, 0.0, sigma, z.data(), n);
                for (std::size_t i = 0; i < n; ++i) next[i] = theta_[i] + z[i];
                if (!(next[0] > 0.0)) {
                    engine_.set_position(start);
                    BoxMullerNormal jump(0.0, sigma);
                    do {
                        for (std::size_t i = 0; i < n; ++i) next[i] = theta_[i] + jump(engine_);
                    } while (!(next[0] > 0.0));
                }
                theta_.swap(next);
            }
        }
        return save_doubles(theta_.data(), theta_.size());
    }
    void prepare() {}
    double log_prior() {  // pyblasso.py:123-133, same operation order
        const double sigma = theta_[0];
        double l1 = 0.0;
        for (int j = 0; j < p_; ++j) l1 += std::fabs(theta_[1 + j]);
        const double log_prior_sigma = -2.0 * std::log(sigma);
        const double log_laplace = p_ * std::log(tau_ / 2 / sigma) - (tau_ * l1) / sigma;
        return log_prior_sigma + log_laplace;
    }
    std::size_t random_position() const { return engine_.position(); }
    std::size_t data_size() const { return data_size_; }
    std::string unparse(const std::string& theta) const {
        std::vector<double> v((std::size_t) p_ + 1);
        load_doubles(theta, v.data(), v.size());
        return unparse_list(v.data(), v.size(), 0);
    }

  private:
    std::size_t data_size_;
    int p_;
    double tau_;
    std::vector<double> theta_, initial_;
    ReplayStream engine_;
    std::shared_ptr<NoiseCache> noise_ = std::make_shared<NoiseCache>();  // shared by the copies (one per virtual worker)
};

// Dense Boltzmann (BASELINE config 2, new workload): state x in R^D, log p = -x'Wx/T, one item.
// Proposal rule in the style of the reference's samplers: x' = x + N(0, 0.1) per unit; initial
// state 0.25 in every unit (boltzmannsampler.hh:41-43,51).
class BoltzmannDenseExecutor {
  public:
    explicit BoltzmannDenseExecutor(int D) : D_(D), theta_((std::size_t) D, 0.25) {}
    int theta_len() const { return D_; }
    std::string load(const JobInfo& job) {
        engine_.set_position(job.random_position);
        if (job.need_initial())
            theta_.assign((std::size_t) D_, 0.25);
        else {
            load_doubles(job.theta, theta_.data(), theta_.size());
            if (job.need_proposal()) {
                z_.resize(theta_.size());
                normal_fill_cached(*noise_, engine_, 0.0, 0.1, z_.data(), z_.size());  // see NoiseCache (rng.hpp)
                for (std::size_t i = 0; i < theta_.size(); ++i) theta_[i] += z_[i];
            }
        }
        return save_doubles(theta_.data(), theta_.size());
    }
    void prepare() {}
    double log_prior() { return 0.0; }
    std::size_t random_position() const { return engine_.position(); }
    std::size_t data_size() const { return 1; }
    std::string unparse(const std::string& theta) const {
        std::vector<double> v((std::size_t) D_);
        load_doubles(theta, v.data(), v.size());
        return unparse_list(v.data(), v.size() < 8 ? v.size() : 8, 0);
    }

  private:
    int D_;
    std::vector<double> theta_, z_;
    ReplayStream engine_;
    std::shared_ptr<NoiseCache> noise_ = std::make_shared<NoiseCache>();  // shared by the copies (one per virtual worker)
};

} }  // namespace pf::host
