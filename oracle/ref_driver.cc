/* oracle/ref_driver.cc -- TEST INFRASTRUCTURE ONLY.  Never linked into the product.
 *
 * Builds the REFERENCE's own code, in place, from /root/reference (nothing is copied into
 * this repository): testmaster.cc (which textually includes masterloop.cc and every
 * header on the path) plus master.cc and stype.cc, against oracle/shim/.  This file adds
 * a main() around them that
 *
 *   unit                         runs the reference's three assert tests
 *                                (testmaster.cc:295-426; they are commented out of its main)
 *   chain  SAMPLER ...           runs the reference JobTree + run_master + MasterTester<S>
 *                                (testmaster.cc:57-293) for L levels and prints every chain
 *                                row the reference would print (master.cc:347-355), at full
 *                                precision, with the raw theta bytes
 *   normal-eval IN OUT           NormalSampler<2>::load/prepare/evaluate_many
 *                                (normalsampler.hh:66-165) over a binary dataset
 *   boltz-eval  IN OUT           BoltzmannSampler::evaluate_many (boltzmannsampler.hh:71-82)
 *   normal-bench N B REPS        times evaluate_many for B nodes (CPU baseline, kind
 *                                "reference", one thread per node like one MPI worker per node)
 *   data N                       prints the first values of the normal.cc:41-46 dataset
 *   mpirun -np N SAMPLER ...     the reference's parallel program, threads for ranks: run_master(world) on rank 0,
 *                                the registered sampler's work() -> worker<>() on ranks 1..N-1, over the
 *                                thread-mailbox boost::mpi::communicator of oracle/shim (the CPU baseline of C1)
 *   frontier SAMPLER ...         the reference JobTree + run_master driven by the product's
 *                                FrontierCommunicator (whole scheduling bursts scored at once by the
 *                                reference's own evaluate_many): prints every JobInfo handed out
 *                                (J rows) and every chain level (R rows) -- tests/golden/frontier_trace_*.tsv
 *
 * Output of `chain` is the golden-vector format of tests/golden/chain_*.tsv:
 *   depth <TAB> accept <TAB> metric_sum(%.17g) <TAB> theta(hex bytes) <TAB> unparse(theta)
 */
#define main testmaster_reference_main
#include "testmaster.cc"
#undef main

#include <chrono>
#include <thread>
#include <vector>

// The PRODUCT's frontier communicator (header-only, generic over the tree / record types),
// instantiated here with the REFERENCE's JobTree, records and samplers and a CPU evaluator: the
// traces it prints are the goldens the product (its own tree + the GPU engine) must reproduce.
#include "frontier_comm.hpp"

namespace {

std::string hex(const std::string& s) {
    static const char* d = "0123456789abcdef";
    std::string out;
    for (unsigned char c : s) {
        out += d[c >> 4];
        out += d[c & 15];
    }
    return out;
}

size_t g_stop_depth = 0;

/* A completion hook (stype.hh:37-39) that prints what the reference passes to notify()
 * (master.cc:347-355) without the wall-clock column, then ends the process once the
 * requested depth has been emitted (testmaster never exits on its own). */
template <typename T>
class GoldenHook : public SamplerType {
  public:
    std::string unparse(const std::string& theta) const { return load_unparse_theta<T>(theta); }
    void notify(const JobTree&, size_t depth, const double metric_sum, const std::string& theta,
                bool accept) {
        printf("%zu\t%d\t%.17g\t%s\t%s\n", depth, (int) accept, metric_sum, hex(theta).c_str(),
               unparse(theta).c_str());
        if (depth >= g_stop_depth) {
            fflush(stdout);
            exit(0);
        }
    }
};

double* normal_dataset(size_t n) {
    /* recipe of samplers/normal.cc:41-46 == testmaster.cc:487-492 */
    boost::random::mt19937 engine;
    boost::random::normal_distribution<> dist;
    double* data = new double[n * 2];
    for (size_t i = 0; i != n * 2; ++i)
        data[i] = dist(engine) + (i % 2 ? 100 : 200);
    return data;
}

int cmd_chain(int argc, char** argv) {
    const char* sampler_name = "normal";
    size_t cores = 9, data_size = 10000, batch_size = 1000, levels = 20, policy = PREDICTIVE,
           dimension = 1;
    double reassign_ratio = 1.1, correlation = 0.99;
    bool threshold = false;
    for (int i = 0; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return i + 1 < argc ? argv[++i] : "0"; };
        if (a == "-n") cores = strtoull(next(), 0, 0);
        else if (a == "-d") data_size = strtoull(next(), 0, 0);
        else if (a == "-b") batch_size = strtoull(next(), 0, 0);
        else if (a == "-l") levels = strtoull(next(), 0, 0);
        else if (a == "-s") policy = strtoull(next(), 0, 0);
        else if (a == "-D") dimension = strtoull(next(), 0, 0);
        else if (a == "-r") reassign_ratio = strtod(next(), 0);
        else if (a == "-c") correlation = strtod(next(), 0);
        else if (a == "-t") threshold = true;
        else sampler_name = argv[i];
    }
    g_stop_depth = levels;
    bool boltz = strcmp(sampler_name, "boltzmann") == 0;
    SamplerType* hook = boltz ? (SamplerType*) new GoldenHook<double>
                              : (SamplerType*) new GoldenHook<NormalTheta<2> >;
    /* same construction sequence as tree.cc:147-150 */
    tree = new JobTree(cores, levels, hook, policy, correlation, dimension);
    tree->set_reassign_ratio(reassign_ratio);
    tree->set_require_comparison_theta(false);
    tree->set_threshold(threshold);
    double t0 = timestamp();
    (void) t0;
    if (boltz) {
        MasterTester<BoltzmannSampler> mt(cores, 0, data_size, batch_size, BoltzmannSampler());
        mt.run(*tree);
    } else {
        double* data = normal_dataset(data_size);
        NormalSampler<2> sampler(data, data_size);
        MasterTester<NormalSampler<2> > mt(cores, 0, data_size, batch_size, sampler);
        mt.run(*tree);
    }
    return 0;
}

// ---- reference tree under the product's frontier communicator ---------------------------------
struct RefTypes {
    using JobInfo = ::JobInfo;
    using ProposalInfo = ::ProposalInfo;
    using MetricChange = ::MetricChange;
    using AbandonInfo = ::AbandonInfo;
};
static_assert(MSG_W2M_REGISTER == (int) pf::host::FC_REGISTER && MSG_W2M_WANTWORK == (int) pf::host::FC_WANTWORK &&
                  MSG_W2M_SETPROPOSAL == (int) pf::host::FC_SETPROPOSAL && MSG_W2M_UPDATE == (int) pf::host::FC_UPDATE &&
                  MSG_W2M_QUIT == (int) pf::host::FC_QUIT,
              "message tags differ from the reference's (protocol.hh:32-41)");

struct RefTrace {
    void job(const JobInfo& in, const ProposalInfo& out) {
        printf("J\t%zu\t%zu\t%d\t%s\t%zu\t%zu\t%zu\t%.17g\t%s\n", in.id, in.generation, in.state,
               in.path.empty() ? "-" : in.path.c_str(), in.random_position, out.random_position,
               out.next_random_position, out.log_prior, hex(out.proposal).c_str());
    }
};

template <typename S>
struct RefEvaluator {  // what MasterTester::check_update does per node (testmaster.cc:170-184), whole node at once
    S sampler;
    size_t n;
    void operator()(const std::vector<pf::host::FrontierJob>& jobs, std::vector<double>& sum, std::vector<double>& sq) {
        for (size_t i = 0; i != jobs.size(); ++i) {
            JobNode* jn = tree->node(jobs[i].id);
            JobInfo job;
            job.state = JobInfo::s_has_proposal;
            job.theta = jobs[i].theta;
            job.random_position = jn ? jn->random_position() : 0;
            (void) sampler.load(job);
            sampler.prepare();
            double s = 0, q = 0;
            sampler.evaluate_many(0, n, s, q);
            sum[i] = s;
            sq[i] = q;
        }
    }
    // items [first, first + count) in each node's own visiting order (what a worker does between two
    // updates, worker.hh:148-155)
    void range(const std::vector<pf::host::FrontierJob>& jobs, size_t first, size_t count, std::vector<double>& sum,
               std::vector<double>& sq) {
        for (size_t i = 0; i != jobs.size(); ++i) {
            JobNode* jn = tree->node(jobs[i].id);
            JobInfo job;
            job.state = JobInfo::s_has_proposal;
            job.theta = jobs[i].theta;
            job.random_position = jn ? jn->random_position() : 0;
            (void) sampler.load(job);
            sampler.prepare();
            double s = 0, q = 0;
            sampler.evaluate_many(first, count, s, q);
            sum[i] = s;
            sq[i] = q;
        }
    }
};

template <typename T>
class TraceHook : public SamplerType {
  public:
    std::string unparse(const std::string& theta) const { return load_unparse_theta<T>(theta); }
    void notify(const JobTree&, size_t depth, const double metric_sum, const std::string& theta, bool accept) {
        printf("R\t%zu\t%d\t%.17g\t%s\n", depth, (int) accept, metric_sum, hex(theta).c_str());
    }
};

template <typename S, typename T>
int run_frontier(const S& sampler, size_t workers, size_t data_size, size_t levels, size_t policy, double reassign,
                 bool eager, size_t batch, int style) {
    TraceHook<T> hook;
    tree = new JobTree(workers + 1, levels, &hook, policy, 0.99, 1);
    tree->set_reassign_ratio(reassign);
    tree->set_require_comparison_theta(false);
    tree->set_threshold(false);
    RefEvaluator<S> evaluator{sampler, data_size};
    RefTrace trace;
    typedef pf::host::FrontierCommunicator<JobTree, S, RefTypes, RefEvaluator<S>, RefTrace> Comm;
    Comm comm(*tree, sampler, evaluator, (int) workers, data_size, &trace, eager, batch, style);
    tree->run_master<Comm, typename Comm::Request, typename Comm::Status>(comm, JobTree::run_behavior().quiet(true));
    printf("# levels %zu engine_calls %zu nodes_scored %zu max_batch %zu items_scored %zu abandoned %zu skipped %zu\n",
           tree->completed_depth(), comm.stats().engine_calls, comm.stats().nodes_scored, comm.stats().max_batch,
           comm.stats().items_scored, comm.stats().abandoned, comm.stats().skipped);
    return 0;
}

int cmd_frontier(int argc, char** argv) {
    const char* sampler_name = "normal";
    size_t workers = 8, data_size = 10000, levels = 50, policy = PREDICTIVE;
    double reassign = 1.1;
    bool eager = true;
    size_t batch = 0;
    int style = 1;
    for (int i = 0; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return i + 1 < argc ? argv[++i] : "0"; };
        if (a == "--deferred") eager = false;
        else if (a == "-b") batch = strtoull(next(), 0, 0);
        else if (a == "-y") style = (int) strtol(next(), 0, 0);
        else if (a == "-W") workers = strtoull(next(), 0, 0);
        else if (a == "-d") data_size = strtoull(next(), 0, 0);
        else if (a == "-l") levels = strtoull(next(), 0, 0);
        else if (a == "-s") policy = strtoull(next(), 0, 0);
        else if (a == "-r") reassign = strtod(next(), 0);
        else sampler_name = argv[i];
    }
    if (strcmp(sampler_name, "boltzmann") == 0)
        return run_frontier<BoltzmannSampler, double>(BoltzmannSampler(), workers, data_size, levels, policy, reassign, eager, batch, style);
    double* data = normal_dataset(data_size);
    NormalSampler<2> sampler(data, data_size);
    return run_frontier<NormalSampler<2>, NormalTheta<2> >(sampler, workers, data_size, levels, policy, reassign, eager, batch, style);
}

struct EvalHeader {
    uint64_t n, nodes, theta_len;
};

/* IN : EvalHeader, double data[n][2], double theta[nodes][5]
 * OUT: double (sum, sum_sq)[nodes]   -- reference visiting order (StrideIndex drawn in
 *      prepare(), normalsampler.hh:115-119) with the replay stream at position 0. */
int cmd_normal_eval(const char* in, const char* out) {
    FILE* f = fopen(in, "rb");
    if (!f) return 2;
    EvalHeader h;
    if (fread(&h, sizeof h, 1, f) != 1) return 2;
    std::vector<double> data(h.n * 2), theta(h.nodes * h.theta_len);
    if (fread(data.data(), 8, data.size(), f) != data.size()) return 2;
    if (fread(theta.data(), 8, theta.size(), f) != theta.size()) return 2;
    fclose(f);
    assert(h.theta_len == (uint64_t) NormalTheta<2>::size);
    NormalSampler<2> s(data.data(), h.n);
    std::vector<double> res(h.nodes * 2);
    for (size_t b = 0; b != h.nodes; ++b) {
        JobInfo job;
        job.state = JobInfo::s_has_proposal;
        job.random_position = 0;
        job.theta = save_theta(NormalTheta<2>(&theta[b * h.theta_len]));
        (void) s.load(job);
        s.prepare();
        double sum = 0, sq = 0;
        s.evaluate_many(0, h.n, sum, sq);
        res[2 * b] = sum;
        res[2 * b + 1] = sq;
    }
    f = fopen(out, "wb");
    fwrite(res.data(), 8, res.size(), f);
    fclose(f);
    return 0;
}

/* IN : EvalHeader (n = count, theta_len = 1), double theta[nodes]; OUT as above. */
int cmd_boltz_eval(const char* in, const char* out) {
    FILE* f = fopen(in, "rb");
    if (!f) return 2;
    EvalHeader h;
    if (fread(&h, sizeof h, 1, f) != 1) return 2;
    std::vector<double> theta(h.nodes);
    if (fread(theta.data(), 8, theta.size(), f) != theta.size()) return 2;
    fclose(f);
    BoltzmannSampler s;
    std::vector<double> res(h.nodes * 2);
    for (size_t b = 0; b != h.nodes; ++b) {
        JobInfo job;
        job.state = JobInfo::s_has_proposal;
        job.random_position = 0;
        job.theta = save_theta(theta[b]);
        (void) s.load(job);
        s.prepare();
        double sum = 0, sq = 0;
        s.evaluate_many(0, h.n, sum, sq);
        res[2 * b] = sum;
        res[2 * b + 1] = sq;
    }
    f = fopen(out, "wb");
    fwrite(res.data(), 8, res.size(), f);
    fclose(f);
    return 0;
}

/* CPU baseline, kind "reference": B frontier nodes, each scored over all n points by the
 * reference's own NormalSampler<2>::evaluate_many on its own thread -- the in-process
 * analogue of `mpirun -np B+1 ./fetching` where every worker holds the whole dataset
 * (normal.cc:43-48) and scores one node (worker.hh:148-155).  Prints one JSON line. */
int cmd_normal_bench(size_t n, size_t nodes, size_t reps, size_t threads) {
    double* data = normal_dataset(n);
    if (threads == 0) threads = std::thread::hardware_concurrency();
    if (threads > nodes) threads = nodes;
    std::vector<double> sums(nodes, 0.0);
    auto t0 = std::chrono::steady_clock::now();
    for (size_t r = 0; r != reps; ++r) {
        std::vector<std::thread> pool;
        for (size_t t = 0; t != threads; ++t)
            pool.emplace_back([&, t]() {
                NormalSampler<2> s(data, n);
                for (size_t b = t; b < nodes; b += threads) {
                    double th[5] = {0.01 * (double) b, -0.02 * (double) b, 1, 0, 1};
                    JobInfo job;
                    job.state = JobInfo::s_has_proposal;
                    job.random_position = 0;
                    job.theta = save_theta(NormalTheta<2>(th));
                    (void) s.load(job);
                    s.prepare();
                    double sum = 0, sq = 0;
                    s.evaluate_many(0, n, sum, sq);
                    sums[b] = sum;
                }
            });
        for (auto& th : pool) th.join();
    }
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("{\"kind\": \"reference\", \"what\": \"NormalSampler<2>::evaluate_many\", \"n\": %zu, "
           "\"nodes\": %zu, \"reps\": %zu, \"threads\": %zu, \"seconds\": %.6f, "
           "\"evals_per_sec\": %.6f, \"item_evals_per_sec\": %.1f, \"check\": %.17g}\n",
           n, nodes, reps, threads, dt, nodes * reps / dt, (double) n * nodes * reps / dt, sums[0]);
    return 0;
}

} // namespace

/* `pf_ref mpirun -np N [-S] SAMPLER [-l L] [-b B] [-y STYLE] [-s POLICY] [-r RATIO] [-c CORR] [-d DIM] [-t]`
 * The reference's parallel program with THREADS for ranks: rank 0 builds the JobTree and runs
 * JobTree::run_master(world, ...) (tree.cc:144-152, master.cc:551-556, masterloop.cc), ranks 1..N-1 run the registered
 * sampler's work() (samplers/normal.cc:38-49 -> worker<NormalSampler<2>>(), worker.hh:53-177), all unmodified and
 * compiled in place; only boost::mpi::communicator is the thread-mailbox shim (oracle/shim/boost/mpi/communicator.hpp).
 * stdout: the stock TSV of stype.cc:67-84; stderr: the stock `OK sample/s` line (tree.cc:160-167). */
int cmd_mpirun(int argc, char** argv) {
    const char* sampler_name = 0;
    int np = 2, chain_length = 100, batch_size = 1000, dimension = 1, update_style = INCREMENTAL,
        scheduling_policy = PREDICTIVE;
    double reassign_ratio = 1.1, correlation = 0.99;
    bool threshold = false;
    for (int i = 0; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return i + 1 < argc ? argv[++i] : "0"; };
        if (a == "-np") np = atoi(next());
        else if (a == "-S") sampler_name = next();
        else if (a == "-l") chain_length = atoi(next());
        else if (a == "-b") batch_size = atoi(next());
        else if (a == "-y") update_style = atoi(next());
        else if (a == "-s") scheduling_policy = atoi(next());
        else if (a == "-r") reassign_ratio = strtod(next(), 0);
        else if (a == "-c") correlation = strtod(next(), 0);
        else if (a == "-d") dimension = atoi(next());
        else if (a == "-t") threshold = true;
        else sampler_name = argv[i];
    }
    if (!sampler_name) sampler_name = "normal";
    if (np < 2) {
        fprintf(stderr, "mpirun: -np must be at least 2 (one master, one worker)\n");
        return 2;
    }
    SamplerType* sampler = SamplerType::find(sampler_name);
    if (!sampler) {
        std::cerr << "No such sampler \"" << sampler_name << "\"\n";
        SamplerType::print_list();
        return 1;
    }
    const double tic = timestamp();
    boost::mpi::shim::World threads_as_ranks(np);
    boost::mpi::shim::world() = &threads_as_ranks;
    boost::mpi::communicator world;
    std::vector<std::thread> ranks;
    for (int r = 1; r < np; ++r)
        ranks.emplace_back([&, r]() {
            boost::mpi::shim::my_rank() = r;
            world.barrier();
            sampler->work(world, (size_t) batch_size, (size_t) update_style, 0);  // tree.cc:153-154
        });
    boost::mpi::shim::my_rank() = 0;
    world.barrier();
    fprintf(stderr, "# pf_ref mpirun -np %d (threads for ranks) -S %s -l %d -b %d -d %d -r %g -c %g%s -y %d -s %d\n", np,
            sampler_name, chain_length, batch_size, dimension, reassign_ratio, correlation, threshold ? " -t" : "",
            update_style, scheduling_policy);
    {
        JobTree master_tree(np, chain_length, sampler, scheduling_policy, correlation, dimension);  // tree.cc:147-152
        master_tree.set_reassign_ratio(reassign_ratio);
        master_tree.set_require_comparison_theta(false);
        master_tree.set_threshold(threshold);
        sampler->notify(master_tree, 0, 0, std::string(), false);
        master_tree.run_master(world, JobTree::run_behavior().quiet(true));
    }
    for (auto& t : ranks) t.join();
    const double toc = timestamp();
    fflush(stdout);
    fprintf(stderr, "OK sample/s = %-7.3f samples = %-6d seconds = %-7.3f N = %d\n", chain_length / (toc - tic),
            chain_length, toc - tic, np);
    boost::mpi::shim::world() = nullptr;
    return 0;
}

#ifndef PF_REF_DRIVER_NO_MAIN  /* ref_gpu_driver.cc includes this file for its helpers and brings its own main() */
int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: pf_ref unit | chain [normal|boltzmann] [-n C] [-d N] [-b B] [-l L] [-s P] "
                        "[-D dim] [-r R] [-c C] [-t] | normal-eval IN OUT | boltz-eval IN OUT | "
                        "normal-bench N NODES REPS [THREADS] | data N | frontier [normal|boltzmann] [-W W] [-d N] [-l L] [-s P] [-r R]\n");
        return 1;
    }
    std::string cmd = argv[1];
    if (cmd == "unit") {
        test_random_sequence();
        test_proposals();
        test_proposals2();
        printf("reference unit tests: test_random_sequence test_proposals test_proposals2 OK\n");
        return 0;
    } else if (cmd == "chain")
        return cmd_chain(argc - 2, argv + 2);
    else if (cmd == "normal-eval" && argc == 4)
        return cmd_normal_eval(argv[2], argv[3]);
    else if (cmd == "boltz-eval" && argc == 4)
        return cmd_boltz_eval(argv[2], argv[3]);
    else if (cmd == "normal-bench" && argc >= 5)
        return cmd_normal_bench(strtoull(argv[2], 0, 0), strtoull(argv[3], 0, 0),
                                strtoull(argv[4], 0, 0), argc > 5 ? strtoull(argv[5], 0, 0) : 0);
    else if (cmd == "frontier")
        return cmd_frontier(argc - 2, argv + 2);
    else if (cmd == "mpirun")
        return cmd_mpirun(argc - 2, argv + 2);
    else if (cmd == "data" && argc == 3) {
        size_t n = strtoull(argv[2], 0, 0);
        double* d = normal_dataset(n);
        for (size_t i = 0; i != n * 2; ++i) printf("%.17g\n", d[i]);
        return 0;
    }
    return 1;
}
#endif
