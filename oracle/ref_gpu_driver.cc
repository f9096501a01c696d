/* oracle/ref_gpu_driver.cc -- TEST INFRASTRUCTURE ONLY.  Never linked into the product.
 *
 * The drop-in, compiled: the REFERENCE's own speculative tree and master loop (master.cc, masterloop.cc,
 * master.hh), its own NormalSampler<2> as the host half of the Executor concept (proposals, priors, replay:
 * normalsampler.hh:66-120), its own records (protocol.hh) -- all compiled in place from /root/reference against
 * oracle/shim, exactly like oracle/_ref/pf_ref -- with the evaluation path replaced by the product's C ABI:
 *
 *     JobTree::run_master<Comm, ...>            masterloop.cc:21-134 (reference)
 *       Comm = pf::host::FrontierCommunicator   fetching_b200/csrc/host/frontier_comm.hpp (product, header only)
 *         EngineEvaluator                       INTEGRATION.md section 2a, as printed there
 *           pf_eval_frontier(...)               include/pfetch.h -> libpfetch.so (CUDA, sm_100a)
 *
 * This is the binding INTEGRATION.md describes for tree.cc, minus MPI: what a maintainer of the reference would
 * build.  tests/test_gpu_dropin.py runs it on the GPU and compares the chain it prints with the reference's own
 * MasterTester chain (tests/golden/chain_normal_d10000.tsv): same accept / reject decisions, same theta bytes.
 *
 *   pf_ref_gpu [normal] [-W frontier] [-d ndata] [-l levels] [-s policy] [-r reassign_ratio] [--device N]
 *
 * Output: the golden-vector format of `pf_ref chain` (depth, accept, metric_sum %.17g, theta hex, unparse(theta)),
 * then a `# ...` line with the hand-off statistics.
 */
#define PF_REF_DRIVER_NO_MAIN
#include "ref_driver.cc"

#include <pfetch.h>

namespace {

// ---- INTEGRATION.md section 2a -------------------------------------------------------------------------------
struct EngineEvaluator {  // frontier -> one kernel pass
    pf_engine* engine;
    uint32_t theta_len;
    std::vector<double> th;
    int status = PF_OK;
    bool failed() const { return status != PF_OK; }
    void stage(const std::vector<pf::host::FrontierJob>& jobs) {
        th.resize(jobs.size() * theta_len);
        for (size_t i = 0; i != jobs.size(); ++i)  // theta bytes are raw doubles (protocol.hh:200-206)
            memcpy(&th[i * theta_len], jobs[i].theta.data(), theta_len * sizeof(double));
    }
    void operator()(const std::vector<pf::host::FrontierJob>& jobs, std::vector<double>& sum, std::vector<double>& sum_sq) {
        stage(jobs);
        status = pf_eval_frontier(engine, (uint32_t) jobs.size(), th.data(), sum.data(), sum_sq.data());
        if (status != PF_OK) fprintf(stderr, "pfetch: %s\n", pf_last_error());
    }
    // partial updates (-b): items in natural order (the reference sampler keeps its StrideIndex private)
    void range(const std::vector<pf::host::FrontierJob>& jobs, size_t first, size_t count, std::vector<double>& sum,
               std::vector<double>& sum_sq) {
        stage(jobs);
        status = pf_eval_frontier_range(engine, (uint32_t) jobs.size(), th.data(), first, count, nullptr, sum.data(),
                                        sum_sq.data());
        if (status != PF_OK) fprintf(stderr, "pfetch: %s\n", pf_last_error());
    }
};
struct NoTrace {
    void job(const JobInfo&, const ProposalInfo&) {}
};

template <typename T>
class ChainHook : public SamplerType {  // prints what the reference passes to notify() (master.cc:347-355)
  public:
    std::string unparse(const std::string& theta) const { return load_unparse_theta<T>(theta); }
    void notify(const JobTree&, size_t depth, const double metric_sum, const std::string& theta, bool accept) {
        printf("%zu\t%d\t%.17g\t%s\t%s\n", depth, (int) accept, metric_sum, hex(theta).c_str(), unparse(theta).c_str());
    }
};

}  // namespace

int main(int argc, char** argv) {
    size_t workers = 16, data_size = 10000, levels = 300, policy = PREDICTIVE;
    double reassign = 1.1;
    int device = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return i + 1 < argc ? argv[++i] : "0"; };
        if (a == "-W") workers = strtoull(next(), 0, 0);
        else if (a == "-d") data_size = strtoull(next(), 0, 0);
        else if (a == "-l") levels = strtoull(next(), 0, 0);
        else if (a == "-s") policy = strtoull(next(), 0, 0);
        else if (a == "-r") reassign = strtod(next(), 0);
        else if (a == "--device") device = (int) strtol(next(), 0, 0);
        else if (a != "normal") {
            fprintf(stderr, "usage: pf_ref_gpu [normal] [-W frontier] [-d ndata] [-l levels] [-s policy] [-r ratio] [--device N]\n");
            return 2;
        }
    }
    double* data = normal_dataset(data_size);  // normal.cc:41-46

    pf_model_desc d = {};
    d.model = PF_MODEL_NORMAL;
    d.device = device;
    d.n_rows = data_size;
    d.dim = 2;
    d.data = data;
    pf_engine* engine = nullptr;
    if (pf_engine_create(&d, &engine) != PF_OK) {
        fprintf(stderr, "pf_engine_create: %s\n", pf_last_error());
        return 1;
    }

    ChainHook<NormalTheta<2> > hook;
    /* same construction sequence as tree.cc:147-150 */
    tree = new JobTree((int) workers + 1, levels, &hook, policy, 0.99, 1);
    tree->set_reassign_ratio(reassign);
    tree->set_require_comparison_theta(false);
    tree->set_threshold(false);

    NormalSampler<2> executor(data, data_size);  // host half of the Executor concept, unchanged
    EngineEvaluator eval{engine, pf_engine_theta_len(engine), {}};
    NoTrace trace;
    typedef pf::host::FrontierCommunicator<JobTree, NormalSampler<2>, RefTypes, EngineEvaluator, NoTrace> Comm;
    Comm comm(*tree, executor, eval, (int) workers, pf_engine_data_size(engine), &trace);
    tree->run_master<Comm, Comm::Request, Comm::Status>(comm, JobTree::run_behavior().quiet(true));
    // one kernel pass per burst: a launch, or a command served by the resident frontier kernel (launched once)
    uint64_t res_calls = 0, res_launches = 0;
    pf_engine_resident_stats(engine, &res_calls, &res_launches);
    printf("# levels %zu engine_calls %zu nodes_scored %zu max_batch %zu kernel_launches %llu resident_calls %llu resident_launches %llu\n",
           tree->completed_depth(), comm.stats().engine_calls, comm.stats().nodes_scored, comm.stats().max_batch,
           (unsigned long long) pf_engine_launch_count(engine), (unsigned long long) res_calls, (unsigned long long) res_launches);
    fflush(stdout);
    const int rc = eval.status;
    pf_engine_destroy(engine);
    return rc == PF_OK ? 0 : 1;
}
