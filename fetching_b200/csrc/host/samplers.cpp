// host/samplers.cpp -- the sampler plugins behind the registry of sampler_type.hpp.
//
// One SamplerType subclass per model, signed up by a static constructor under the name the reference uses
// (samplers/normal.cc:59-62 "normal"; python.cc:220-225 "py" / "python", which loads the MODULE named by its first
// argument: `fetching py pymixture train_size=...`).  set_arguments() parses `key=value` the way the Python models'
// sampler(args) factories do (pymixture.py:193-204, pyblasso.py:254-267): an argument that is not exactly one
// key=value pair, or a non-integer value for an integer key, is an error (-1; complaint on stderr when asked).
// work() = what a reference worker does before its loop -- build the data set -- followed by handing it to the GPU
// engine (pf_engine_create) and running the chain through it (pf_chain_run); the chain rows are printed by
// pf_chain_run's completion hook exactly like SamplerType::notify (stype.cc:67-84).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "../../../include/pfetch.h"
#include "executors.hpp"
#include "rng.hpp"
#include "sampler_type.hpp"

namespace pf { namespace host {
namespace {

// dict([arg.split('=') for arg in args]): every argument must split into exactly two parts
int parse_kv(const std::vector<const char*>& argv, std::size_t from, std::map<std::string, std::string>& kv,
             bool complain) {
    for (std::size_t i = from; i < argv.size(); ++i) {
        const char* a = argv[i];
        const char* eq = strchr(a, '=');
        if (!eq || strchr(eq + 1, '=')) {
            if (complain) std::cerr << "sampler argument \"" << a << "\": expected key=value\n";
            return -1;
        }
        kv[std::string(a, eq - a)] = eq + 1;
    }
    return 0;
}

// int(a['key']): the whole value must be an integer
int kv_long(const std::map<std::string, std::string>& kv, const char* key, long& value, bool complain) {
    auto it = kv.find(key);
    if (it == kv.end()) return 0;
    char* end = nullptr;
    const long v = strtol(it->second.c_str(), &end, 10);
    if (end == it->second.c_str() || *end != '\0') {
        if (complain) std::cerr << "sampler argument " << key << "=" << it->second << ": not an integer\n";
        return -1;
    }
    value = v;
    return 0;
}

// engine + chain, shared by every plugin's work()
int run_on_engine(pf_model_desc& d, int device, const pf_chain_config& cfg, pf_chain_stats* stats) {
    d.device = device;
    pf_engine* engine = nullptr;
    if (pf_engine_create(&d, &engine) != PF_OK) return 1;
    const int rc = pf_chain_run(engine, &cfg, nullptr, nullptr, nullptr, stats);
    pf_engine_destroy(engine);
    return rc == PF_OK ? 0 : 1;
}

// ---------------------------------------------------------------------------------------- normal
class NormalSamplerType : public SamplerType {  // samplers/normal.cc:23-62
  public:
    int set_arguments(std::vector<const char*>& argv, bool complain) override {
        std::map<std::string, std::string> kv;
        if (parse_kv(argv, 0, kv, complain) < 0) return -1;
        long n = (long) ndata_;
        if (kv_long(kv, "ndata", n, complain) < 0 || n < 1) return -1;  // (the reference's NDATA is a constant 10^6)
        ndata_ = (std::size_t) n;
        return 0;
    }
    int work(int device, const pf_chain_config& cfg, pf_chain_stats* stats) override {
        Mt19937 engine;  // default seeded; one long-lived distribution object (normal.cc:41-46)
        BoxMullerNormal gauss;
        std::vector<double> data(ndata_ * 2);
        for (std::size_t i = 0; i != ndata_ * 2; ++i) data[i] = gauss(engine) + (i % 2 ? 100 : 200);
        pf_model_desc d;
        memset(&d, 0, sizeof d);
        d.model = PF_MODEL_NORMAL, d.n_rows = ndata_, d.dim = 2, d.data = data.data();
        return run_on_engine(d, device, cfg, stats);
    }
    std::string unparse(const std::string& theta) const override { return NormalExecutor<2>::unparse(theta); }

  private:
    std::size_t ndata_ = 1000000;  // NDATA, normal.cc:21
};

// ---------------------------------------------------------------------------------------- scalar Boltzmann
class BoltzmannSamplerType : public SamplerType {  // boltzmannsampler.hh (samplers/boltzmann.cc is not in the build)
  public:
    int set_arguments(std::vector<const char*>& argv, bool complain) override {
        std::map<std::string, std::string> kv;
        if (parse_kv(argv, 0, kv, complain) < 0) return -1;
        long n = (long) data_size_;
        if (kv_long(kv, "data_size", n, complain) < 0 || n < 1) return -1;
        data_size_ = (std::size_t) n;
        return 0;
    }
    int work(int device, const pf_chain_config& cfg, pf_chain_stats* stats) override {
        pf_model_desc d;
        memset(&d, 0, sizeof d);
        d.model = PF_MODEL_BOLTZMANN_SCALAR, d.n_rows = data_size_, d.dim = 1;
        return run_on_engine(d, device, cfg, stats);
    }
    std::string unparse(const std::string& theta) const override { return BoltzmannScalarExecutor::unparse(theta); }

  private:
    std::size_t data_size_ = 1;
};

// ---------------------------------------------------------------------------------------- pymixture
// pymixture.py:38-54: component positions in the unit hypercube (scaled to length 4, centred, at :54)
const double kMixturePos[8][8] = {
    {0.2456, 0.8211, 0.3065, 0.9171, 0.9674, 0.5055, 0.535, 0.7781}, {0.1852, 0.774, 0.9248, 0.8285, 0.7948, 0.460, 0.9904, 0.6430},
    {0.7135, 0.8969, 0.7882, 0.7179, 0.8707, 0.1549, 0.364, 0.7309}, {0.3507, 0.8099, 0.0669, 0.2366, 0.7635, 0.5878, 0.5188, 0.7846},
    {0.186, 0.3913, 0.7746, 0.3846, 0.1483, 0.4110, 0.5936, 0.5528},  {0.2550, 0.7924, 0.5779, 0.5291, 0.2643, 0.7684, 0.3859, 0.9556},
    {0.3698, 0.1247, 0.1504, 0.8657, 0.9061, 0.2281, 0.9170, 0.9552}, {0.354, 0.3176, 0.2076, 0.0267, 0.6507, 0.0931, 0.2434, 0.2387}};

class PyMixtureSamplerType : public SamplerType {  // pymixture.py: Mixture, sampler(args) :193-204
  public:
    int set_arguments(std::vector<const char*>& argv, bool complain) override {
        std::map<std::string, std::string> kv;
        if (parse_kv(argv, 0, kv, complain) < 0) return -1;
        train_ = 1000, seed_ = 0;
        if (kv_long(kv, "train_size", train_, complain) < 0 || train_ < 1) return -1;
        batch_ = train_ / 100;  // pymixture.py:195 (integer division), unless given
        if (kv_long(kv, "seed", seed_, complain) < 0 || kv_long(kv, "batch_size", batch_, complain) < 0) return -1;
        if (batch_ < 1) batch_ = 1;
        return 0;
    }
    int work(int device, const pf_chain_config& cfg_in, pf_chain_stats* stats) override {
        Mt19937 engine(0);  // np.random.seed(0) in __init__ (:33); rows drawn natively: component ~ U{0..7}, x ~ N(pos, I)
        BoxMullerNormal gauss;
        std::vector<double> data((std::size_t) train_ * 8);
        for (long r = 0; r < train_; ++r) {
            const int c = (int) (numpy_random_sample(engine) * 8);
            for (int j = 0; j < 8; ++j) data[(std::size_t) r * 8 + j] = 4.0 * kMixturePos[c][j] - 2.0 + gauss(engine);
        }
        pf_model_desc d;
        memset(&d, 0, sizeof d);
        d.model = PF_MODEL_MIXTURE, d.n_rows = (uint64_t) train_, d.dim = 8, d.components = 8;
        d.item_rows = (uint64_t) batch_, d.data = data.data();
        pf_chain_config cfg = cfg_in;
        cfg.seed = (uint32_t) seed_;  // seed of first_proposal (pymixture.py:71)
        return run_on_engine(d, device, cfg, stats);
    }
    std::string unparse(const std::string& theta) const override {
        std::vector<double> v(theta.size() / sizeof(double));
        load_doubles(theta, v.data(), v.size());
        return unparse_list(v.data(), v.size(), 8);
    }

  private:
    long train_ = 1000, seed_ = 0, batch_ = 10;
};

// ---------------------------------------------------------------------------------------- pyblasso
class PyBlassoSamplerType : public SamplerType {  // pyblasso.py: BLasso, sampler(args) :254-267
  public:
    int set_arguments(std::vector<const char*>& argv, bool complain) override {
        std::map<std::string, std::string> kv;
        if (parse_kv(argv, 0, kv, complain) < 0) return -1;
        train_ = 100000, p_ = 56, batch_ = 18000, seed_ = 0;
        if (kv_long(kv, "train_size", train_, complain) < 0 || kv_long(kv, "p", p_, complain) < 0 ||
            kv_long(kv, "batch_size", batch_, complain) < 0 || kv_long(kv, "seed", seed_, complain) < 0)
            return -1;
        if (train_ < 1 || p_ < 1 || batch_ < 1 || batch_ > train_) {
            if (complain) std::cerr << "pyblasso: need train_size >= batch_size >= 1 and p >= 1\n";
            return -1;
        }
        return 0;
    }
    // The OPV .npy files the reference reads (pyblasso.py:37-45) are not in its repository: the design matrix is a
    // synthetic stand-in with the reference's preprocessing (columns standardised, response centred, :55-61).
    int work(int device, const pf_chain_config& cfg_in, pf_chain_stats* stats) override {
        const long train = train_, p = p_;
        Mt19937 engine((uint32_t) seed_);
        BoxMullerNormal gauss;
        std::vector<double> beta(p), data((std::size_t) train * p), resp(train), mean(p, 0.0), var(p, 0.0);
        for (long j = 0; j < p; ++j) beta[j] = (j % 3 == 0) ? gauss(engine) : 0.0;
        for (double& v : data) v = gauss(engine);
        for (long r = 0; r < train; ++r)
            for (long j = 0; j < p; ++j) mean[j] += data[(std::size_t) r * p + j] / train;
        for (long r = 0; r < train; ++r)
            for (long j = 0; j < p; ++j) var[j] += std::pow(data[(std::size_t) r * p + j] - mean[j], 2) / train;
        double ymean = 0.0;
        for (long r = 0; r < train; ++r) {
            double yv = gauss(engine);
            for (long j = 0; j < p; ++j) {
                double& x = data[(std::size_t) r * p + j];
                x = (x - mean[j]) / std::sqrt(var[j]);
                yv += beta[j] * x;
            }
            resp[r] = yv;
            ymean += yv / train;
        }
        for (double& v : resp) v -= ymean;
        std::vector<double> init(p + 1, 0.0);
        init[0] = 1.0;
        for (long j = 0; j < p; ++j) init[1 + j] = beta[j];
        pf_model_desc d;
        memset(&d, 0, sizeof d);
        d.model = PF_MODEL_BLASSO, d.n_rows = (uint64_t) train, d.dim = (uint32_t) p, d.item_rows = (uint64_t) batch_;
        d.data = data.data(), d.response = resp.data();
        pf_chain_config cfg = cfg_in;
        cfg.initial_theta = init.data();
        return run_on_engine(d, device, cfg, stats);
    }
    std::string unparse(const std::string& theta) const override {
        std::vector<double> v(theta.size() / sizeof(double));
        load_doubles(theta, v.data(), v.size());
        return unparse_list(v.data(), v.size(), 0);
    }

  private:
    long train_ = 100000, p_ = 56, batch_ = 18000, seed_ = 0;
};

// ---------------------------------------------------------------------------------------- py MODULE args...
// python.cc:96-139: the first argument names the module, the rest go to its sampler(args).  The two modules the
// reference ships are native here; anything else is "No module named ..." as Python would say.
class PythonSamplerType : public SamplerType {
  public:
    int set_arguments(std::vector<const char*>& argv, bool complain) override {
        if (argv.size() < 1) {
            if (complain) std::cerr << "too few arguments, expected 'fetching py MODULE\n";  // python.cc:98-99
            return -1;
        }
        module_ = SamplerType::find(argv[0]);
        if (!module_ || module_ == this || (strcmp(argv[0], "pymixture") != 0 && strcmp(argv[0], "pyblasso") != 0)) {
            if (complain) std::cerr << "ImportError: No module named " << argv[0] << "\n";
            module_ = nullptr;
            return -1;
        }
        std::vector<const char*> rest(argv.begin() + 1, argv.end());
        return module_->set_arguments(rest, complain);
    }
    int work(int device, const pf_chain_config& cfg, pf_chain_stats* stats) override {
        return module_ ? module_->work(device, cfg, stats) : -1;
    }
    std::string unparse(const std::string& theta) const override {
        return module_ ? module_->unparse(theta) : SamplerType::unparse(theta);
    }

  private:
    SamplerType* module_ = nullptr;
};

// heap-allocated, never freed, signed up before main() -- like every reference plugin (normal.cc:59-62)
__attribute__((constructor)) void sign_up_samplers() {
    SamplerType::signup("normal", new NormalSamplerType);
    SamplerType::signup("boltzmann", new BoltzmannSamplerType);
    SamplerType::signup("pymixture", new PyMixtureSamplerType);
    SamplerType::signup("pyblasso", new PyBlassoSamplerType);
    SamplerType* py = new PythonSamplerType;  // python.cc:220-225: one instance under two names
    SamplerType::signup("py", py);
    SamplerType::signup("python", py);
}

}  // namespace
} }  // namespace pf::host
