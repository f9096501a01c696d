// host/pfetch_main.cpp -- `pfetch`: the command line of the reference's `fetching` binary
// (tree.cc:43-170) in front of the GPU engine.
//
//   pfetch [-n NP] [-g DEVICE] [-S NAME] [-l CHAINLENGTH] [-b BATCH] [-d DIM] [-r RATIO] [-u USEC]
//          [-y STYLE] [-s POLICY] [-c CORR] [-t] [-p] [-h]  [sampler [key=value ...]]
//
// Same flags with the same meaning and defaults (tree.cc:45-48,62-110); `-n NP` replaces
// `mpirun -np NP` (NP - 1 = nodes scored per kernel pass; default 2 as in the README example).
// -b BATCH / -y STYLE turn on partial updates every BATCH items (predictive scheduling from partial sums,
// worker.hh:141-175); the default -b 0 scores nodes whole, one kernel pass per scheduling burst.  -u is accepted
// and ignored (it injects artificial latency in the reference).
// stdout: the chain TSV of stype.cc:67-84; stderr: the command echo (tree.cc:145), the -p table
// (master.cc:324-345) and the final `OK sample/s = ...` line (tree.cc:163-168).
//
// Samplers: normal (samplers/normal.cc: 10^6 points, same recipe), boltzmann
// (boltzmannsampler.hh; data_size=N), pymixture | py pymixture (pymixture.py: K=8, d=8, the
// hard-coded component positions; train_size= seed= batch_size=; data drawn natively, see
// DESIGN.md), pyblasso (synthetic design: train_size= p= batch_size=).
#include <getopt.h>
#include <sys/time.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../../include/pfetch.h"
#include "rng.hpp"

namespace {

double now() {
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    return tv.tv_sec + tv.tv_usec * 1e-6;
}

void usage() {  // tree.cc:26-41
    fputs("`pfetch` is a prefetching Metropolis-Hastings sampler (B200 engine)\n\n"
          "Usage: pfetch [-l CHAINLENGTH]\n\n"
          "Options:\n"
          " -h              Print this help statement and exit.\n"
          " -n NP           Like `mpirun -np NP`: NP - 1 nodes are scored per kernel pass [2].\n"
          " -g DEVICE       CUDA device [0].\n"
          " -S NAME         Draw samples from NAME (default is normal).\n"
          " -l SAMPLES      Draw SAMPLES number of samples.\n"
          " -d DIM          Dimension of theta (initial proposal scale).\n"
          " -r REASSIGNRATIO Reassign if REASSIGNRATIO* better work exists [1.1].\n"
          " -s POLICY       Scheduling policy: 0 naive, 1 constant, 2 predictive [2].\n"
          " -c CORRELATION  [0.99]   -t threshold mode   -p per-level table on stderr\n"
          "Example:\n"
          " pfetch -n 9 -l 1024 normal\n",
          stdout);
}

std::map<std::string, std::string> parse_kv(const std::vector<const char*>& args) {
    std::map<std::string, std::string> kv;
    for (const char* a : args) {
        const char* eq = strchr(a, '=');
        if (eq) kv[std::string(a, eq - a)] = eq + 1;
    }
    return kv;
}

long kv_long(const std::map<std::string, std::string>& kv, const char* k, long dflt) {
    auto it = kv.find(k);
    return it == kv.end() ? dflt : strtol(it->second.c_str(), nullptr, 10);
}

// pymixture.py:38-54: component positions, length 4 hypercube
const double kMixturePos[8][8] = {
    {0.2456, 0.8211, 0.3065, 0.9171, 0.9674, 0.5055, 0.535, 0.7781}, {0.1852, 0.774, 0.9248, 0.8285, 0.7948, 0.460, 0.9904, 0.6430},
    {0.7135, 0.8969, 0.7882, 0.7179, 0.8707, 0.1549, 0.364, 0.7309}, {0.3507, 0.8099, 0.0669, 0.2366, 0.7635, 0.5878, 0.5188, 0.7846},
    {0.186, 0.3913, 0.7746, 0.3846, 0.1483, 0.4110, 0.5936, 0.5528},  {0.2550, 0.7924, 0.5779, 0.5291, 0.2643, 0.7684, 0.3859, 0.9556},
    {0.3698, 0.1247, 0.1504, 0.8657, 0.9061, 0.2281, 0.9170, 0.9552}, {0.354, 0.3176, 0.2076, 0.0267, 0.6507, 0.0931, 0.2434, 0.2387}};

}  // namespace

int main(int argc, char** argv) {
    using namespace pf::host;
    const char* sampler_name = nullptr;
    int np = 2, device = 0, chain_length = 100, batch_size = 0, usleep_time = 0, dimension = 1;
    double reassign_ratio = 1.1, correlation = 0.99;
    bool threshold = false, quiet = true;
    int update_style = 1, scheduling_policy = 2;
    const double tic = now();

    for (int optc; (optc = getopt(argc, argv, "htfpS:l:D:b:d:r:u:y:s:c:n:g:")) != -1;) {
        switch (optc) {
            case 'h': usage(); return 0;
            case 'S': sampler_name = optarg; break;
            case 'D': fputs("Prefer -l CHAINLENGTH argument.\n", stderr); /* fallthru */
            case 'l': chain_length = strtol(optarg, 0, 10); break;
            case 'b': batch_size = strtol(optarg, 0, 10); break;
            case 'd': dimension = strtol(optarg, 0, 10); break;
            case 'r': reassign_ratio = strtod(optarg, 0); break;
            case 'u': usleep_time = strtol(optarg, 0, 10); break;
            case 'y': update_style = strtol(optarg, 0, 10); break;
            case 's': scheduling_policy = strtol(optarg, 0, 10); break;
            case 'c': correlation = strtod(optarg, 0); break;
            case 't': threshold = true; break;
            case 'p': quiet = false; break;
            case 'n': np = strtol(optarg, 0, 10); break;
            case 'g': device = strtol(optarg, 0, 10); break;
            default: return 1;
        }
    }
    std::vector<const char*> sampler_args;
    if (!sampler_name && optind < argc) sampler_name = argv[optind++];
    while (optind < argc) sampler_args.push_back(argv[optind++]);
    if (!sampler_name) sampler_name = "normal";
    std::string name = sampler_name;
    if ((name == "py" || name == "python") && !sampler_args.empty()) {  // `fetching py MODULE args` (python.cc:96-101)
        name = sampler_args[0];
        sampler_args.erase(sampler_args.begin());
    }
    if (np < 2) np = 2;
    const auto kv = parse_kv(sampler_args);

    pf_model_desc d;
    memset(&d, 0, sizeof d);
    d.device = device;
    std::vector<double> data, resp, init;
    pf_chain_config cfg;
    memset(&cfg, 0, sizeof cfg);
    if (name == "normal") {  // samplers/normal.cc:38-48
        const size_t n = (size_t) kv_long(kv, "ndata", 1000000);
        Mt19937 engine;
        BoxMullerNormal gauss;
        data.resize(n * 2);
        for (size_t i = 0; i != n * 2; ++i) data[i] = gauss(engine) + (i % 2 ? 100 : 200);
        d.model = PF_MODEL_NORMAL, d.n_rows = n, d.dim = 2, d.data = data.data();
    } else if (name == "boltzmann") {
        d.model = PF_MODEL_BOLTZMANN_SCALAR, d.n_rows = (uint64_t) kv_long(kv, "data_size", 1), d.dim = 1;
    } else if (name == "pymixture") {  // pymixture.py:32-64,193-204
        const long train = kv_long(kv, "train_size", 1000);
        const long batch = kv_long(kv, "batch_size", train / 100 > 0 ? train / 100 : 1);
        cfg.seed = (uint32_t) kv_long(kv, "seed", 0);
        Mt19937 engine(0);  // np.random.seed(0) in __init__; rows drawn natively (component ~ U{0..7}, x ~ N(pos, I))
        BoxMullerNormal gauss;
        data.resize((size_t) train * 8);
        for (long r = 0; r < train; ++r) {
            const int c = (int) (numpy_random_sample(engine) * 8);
            for (int j = 0; j < 8; ++j) data[(size_t) r * 8 + j] = 4.0 * kMixturePos[c][j] - 2.0 + gauss(engine);
        }
        d.model = PF_MODEL_MIXTURE, d.n_rows = train, d.dim = 8, d.components = 8, d.item_rows = batch;
        d.data = data.data();
    } else if (name == "pyblasso") {  // synthetic stand-in for the absent OPV data (pyblasso.py:37-69)
        const long train = kv_long(kv, "train_size", 100000), p = kv_long(kv, "p", 56);
        const long batch = kv_long(kv, "batch_size", 18000);
        Mt19937 engine(0);
        BoxMullerNormal gauss;
        std::vector<double> beta(p);
        for (long j = 0; j < p; ++j) beta[j] = (j % 3 == 0) ? gauss(engine) : 0.0;
        data.resize((size_t) train * p);
        resp.resize(train);
        std::vector<double> mean(p, 0.0), var(p, 0.0);
        for (double& v : data) v = gauss(engine);
        for (long r = 0; r < train; ++r)
            for (long j = 0; j < p; ++j) mean[j] += data[(size_t) r * p + j] / train;
        for (long r = 0; r < train; ++r)
            for (long j = 0; j < p; ++j) var[j] += std::pow(data[(size_t) r * p + j] - mean[j], 2) / train;
        double ymean = 0.0;
        for (long r = 0; r < train; ++r) {  // standardise X, centre y (pyblasso.py:60-61)
            double yv = gauss(engine);
            for (long j = 0; j < p; ++j) {
                double& x = data[(size_t) r * p + j];
                x = (x - mean[j]) / std::sqrt(var[j]);
                yv += beta[j] * x;
            }
            resp[r] = yv;
            ymean += yv / train;
        }
        for (double& v : resp) v -= ymean;
        init.assign(p + 1, 0.0);
        init[0] = 1.0;
        for (long j = 0; j < p; ++j) init[1 + j] = beta[j];
        cfg.initial_theta = init.data();
        d.model = PF_MODEL_BLASSO, d.n_rows = train, d.dim = (uint32_t) p, d.item_rows = batch;
        d.data = data.data(), d.response = resp.data();
    } else {
        fprintf(stderr, "No such sampler \"%s\"\nnormal\nboltzmann\npymixture\npyblasso\n", sampler_name);
        return 1;
    }

    {
        char host[1000];
        gethostname(host, sizeof host);
        fprintf(stderr, "# 0 = %s\n", host);  // tree.cc:133-139
    }
    fprintf(stderr, "# pfetch -n %d -S %s -l %d -b %d -d %d -r %g -c %g -u %d%s -y %d -s %d\n", np, name.c_str(),
            chain_length, batch_size, dimension, reassign_ratio, correlation, usleep_time, threshold ? " -t" : "",
            update_style, scheduling_policy);  // tree.cc:145

    pf_engine* engine = nullptr;
    if (pf_engine_create(&d, &engine) != PF_OK) {
        fprintf(stderr, "pfetch: %s\n", pf_last_error());
        return 1;
    }
    cfg.frontier = np - 1;
    cfg.scheduling_policy = scheduling_policy;
    cfg.chain_length = chain_length;
    cfg.correlation = correlation;
    cfg.reassign_ratio = reassign_ratio;
    cfg.dimension = dimension;
    cfg.threshold = threshold;
    cfg.verbose = !quiet;
    cfg.print_tsv = 1;
    cfg.batch_items = batch_size > 0 ? (uint64_t) batch_size : 0;
    cfg.update_style = (uint32_t) update_style;
    pf_chain_stats st;
    memset(&st, 0, sizeof st);
    const int rc = pf_chain_run(engine, &cfg, nullptr, nullptr, nullptr, &st);
    if (rc != PF_OK) fprintf(stderr, "pfetch: chain failed: %s\n", pf_last_error());
    pf_engine_destroy(engine);
    const double toc = now();
    fflush(stdout);
    fprintf(stderr, "OK sample/s = %-7.3f samples = %-6d seconds = %-7.3f N = %d\n", chain_length / (toc - tic),
            chain_length, toc - tic, np);  // tree.cc:163-168
    fprintf(stderr, "# chain only: %.3f sample/s, %llu engine calls, %llu node passes, %llu node x item evaluations, "
                    "%llu nodes dropped early, largest frontier %llu\n",
            st.seconds > 0 ? st.levels / st.seconds : 0.0, (unsigned long long) st.engine_calls,
            (unsigned long long) st.nodes_scored, (unsigned long long) st.items_scored,
            (unsigned long long) st.abandoned, (unsigned long long) st.max_frontier);
    return rc == PF_OK ? 0 : 1;
}
