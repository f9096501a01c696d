// host/pfetch_main.cpp -- `pfetch`: the command line of the reference's `fetching` binary
// (tree.cc:43-170) in front of the GPU engine.
//
//   pfetch [-n NP] [-g DEVICE] [-S NAME] [-l CHAINLENGTH] [-b BATCH] [-d DIM] [-r RATIO] [-u USEC]
//          [-y STYLE] [-s POLICY] [-c CORR] [-t] [-p] [-h]  [sampler [key=value ...]]
//
// Same flags with the same meaning and defaults (tree.cc:45-48,62-110); `-n NP` replaces
// `mpirun -np NP` (NP - 1 = nodes scored per kernel pass; default 2 as in the README example).
// -b BATCH (default 1000, tree.cc:45) / -y STYLE: partial updates every BATCH items -- predictive scheduling from
// partial sums, reassignment and ABANDON (worker.hh:141-175); a BATCH of 0, or one that covers the data set, scores
// nodes whole (one kernel pass per scheduling burst).  -u is accepted and ignored: it injects artificial per-batch
// latency in the reference's workers (tree.cc:89-91), and there are no worker processes here.
// stdout: the chain TSV of stype.cc:67-84; stderr: the command echo (tree.cc:145), the -p table
// (master.cc:324-345) and the final `OK sample/s = ...` line (tree.cc:163-168).
//
// The sampler is looked up in the plugin registry (SamplerType::find, stype.cc:36-43; tree.cc:122-131) and gets the
// remaining arguments through set_arguments(); an unknown name lists the registered ones (print_list).  Plugins:
// host/samplers.cpp.
#include <getopt.h>
#include <sys/time.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../../../include/pfetch.h"
#include "sampler_type.hpp"

namespace {

double now() {
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    return tv.tv_sec + tv.tv_usec * 1e-6;
}

void usage() {  // tree.cc:26-41
    fputs("`pfetch` is a prefetching Metropolis-Hastings sampler (B200 engine)\n\n"
          "Usage: pfetch [-l CHAINLENGTH] [sampler [key=value ...]]\n\n"
          "Options:\n"
          " -h              Print this help statement and exit.\n"
          " -n NP           Like `mpirun -np NP`: NP - 1 nodes are scored per kernel pass [2].\n"
          " -g DEVICE       CUDA device [0].\n"
          " -S NAME         Draw samples from NAME (default is normal).\n"
          " -l SAMPLES      Draw SAMPLES number of samples [100].\n"
          " -b BATCH        Report partial sums every BATCH data items [1000]; 0 scores nodes whole.\n"
          " -y STYLE        Update style: 0 fast/slow, 1 incremental, 2 logarithmic [1].\n"
          " -u USEC         Accepted and ignored (artificial worker latency in the reference).\n"
          " -d DIM          Dimension of theta (initial proposal scale).\n"
          " -r REASSIGNRATIO Reassign if REASSIGNRATIO* better work exists [1.1].\n"
          " -s POLICY       Scheduling policy: 0 naive, 1 constant, 2 predictive [2].\n"
          " -c CORRELATION  [0.99]   -t threshold mode   -p per-level table on stderr\n"
          "Example:\n"
          " pfetch -n 9 -l 1024 normal\n",
          stdout);
}

}  // namespace

int main(int argc, char** argv) {
    using namespace pf::host;
    const char* sampler_name = nullptr;
    int np = 2, device = 0, chain_length = 100, batch_size = 1000, usleep_time = 0, dimension = 1;  // tree.cc:45-48
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
            default: return 1;  // (also -f: in the reference's getopt string without a case, tree.cc:62,107-108)
        }
    }
    // Additional arguments are passed to the sampler (tree.cc:112-119).
    std::vector<const char*> sampler_args;
    if (!sampler_name && optind < argc) sampler_name = argv[optind++];
    while (optind < argc) sampler_args.push_back(argv[optind++]);
    if (!sampler_name) sampler_name = "normal";
    if (np < 2) np = 2;
    if (update_style < 0 || update_style > 2) {
        fprintf(stderr, "pfetch: -y %d: update style must be 0, 1 or 2\n", update_style);
        return 1;
    }

    // Lookup sampler in list of registered samplers (tree.cc:121-131).
    SamplerType* sampler = SamplerType::find(sampler_name);
    if (!sampler) {
        std::cerr << "No such sampler \"" << sampler_name << "\"\n";
        SamplerType::print_list(std::cerr);
        return 1;
    }
    if (sampler->set_arguments(sampler_args, true) < 0) return 1;

    {
        char host[1000];
        gethostname(host, sizeof host);
        fprintf(stderr, "# 0 = %s\n", host);  // tree.cc:133-139
    }
    fprintf(stderr, "# pfetch -n %d -S %s -l %d -b %d -d %d -r %g -c %g -u %d%s -y %d -s %d\n", np, sampler_name,
            chain_length, batch_size, dimension, reassign_ratio, correlation, usleep_time, threshold ? " -t" : "",
            update_style, scheduling_policy);  // tree.cc:145

    pf_chain_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.frontier = np - 1;
    cfg.scheduling_policy = scheduling_policy;
    cfg.chain_length = chain_length;
    cfg.correlation = correlation;
    cfg.reassign_ratio = reassign_ratio;
    cfg.dimension = dimension;
    cfg.threshold = threshold;
    cfg.verbose = !quiet;
    cfg.print_tsv = 1;
    cfg.batch_items = batch_size > 0 ? (uint64_t) batch_size : 0;  // >= data_size: whole nodes (FrontierCommunicator)
    cfg.update_style = (uint32_t) update_style;
    pf_chain_stats st;
    memset(&st, 0, sizeof st);
    const int rc = sampler->work(device, cfg, &st);
    fflush(stdout);
    if (rc != 0) {  // no `OK` line after a failed run
        fprintf(stderr, "pfetch: %s\n", rc < 0 ? "this sampler cannot work" : pf_last_error());
        return 1;
    }
    const double toc = now();
    fprintf(stderr, "OK sample/s = %-7.3f samples = %-6d seconds = %-7.3f N = %d\n", chain_length / (toc - tic),
            chain_length, toc - tic, np);  // tree.cc:163-168
    fprintf(stderr, "# chain only: %.3f sample/s, %llu engine calls, %llu node passes, %llu node x item evaluations, "
                    "%llu nodes dropped early, largest frontier %llu\n",
            st.seconds > 0 ? st.levels / st.seconds : 0.0, (unsigned long long) st.engine_calls,
            (unsigned long long) st.nodes_scored, (unsigned long long) st.items_scored,
            (unsigned long long) st.abandoned, (unsigned long long) st.max_frontier);
    return 0;
}
