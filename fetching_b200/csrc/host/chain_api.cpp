// host/chain_api.cpp -- pf_chain_run: a whole prefetching-MH chain through the GPU engine.
//
// Wires together SpecTree (speculative tree + master loop), the sampler's host Executor
// (proposals, prior, random-stream bookkeeping) and FrontierCommunicator, whose evaluator is
// pf_eval_frontier: every scheduling burst of the master becomes one kernel pass.
// Equivalent of `mpirun -np W+1 ./fetching -S SAMPLER -l L ...` (tree.cc:43-170), with the MPI
// workers replaced by the C-ABI engine.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../../../include/pfetch.h"
#include "executors.hpp"
#include "frontier_comm.hpp"
#include "sampler_type.hpp"
#include "spec_tree.hpp"

namespace pf {
void set_error(const char* fmt, ...);  // pf_engine.cu: the text behind pf_last_error()
}

namespace pf { namespace host {

struct HostTypes {
    using JobInfo = pf::host::JobInfo;
    using ProposalInfo = pf::host::ProposalInfo;
    using MetricChange = pf::host::MetricChange;
    using AbandonInfo = pf::host::AbandonInfo;
};

static_assert((int) MSG_W2M_REGISTER == (int) FC_REGISTER && (int) MSG_W2M_WANTWORK == (int) FC_WANTWORK &&
                  (int) MSG_W2M_SETPROPOSAL == (int) FC_SETPROPOSAL && (int) MSG_W2M_UPDATE == (int) FC_UPDATE &&
                  (int) MSG_W2M_QUIT == (int) FC_QUIT,
              "tags");

// completion hook: forwards every chain row to the caller (and prints the TSV when asked to)
template <class Executor>
class ChainHook : public SamplerType {
  public:
    ChainHook(const Executor& ex, uint32_t theta_len, pf_chain_row_fn fn, void* user, bool print)
        : ex_(ex), theta_len_(theta_len), fn_(fn), user_(user), print_(print) {}
    std::string unparse(const std::string& theta) const override { return ex_.unparse(theta); }
    void notify(const SpecTree& tree, std::size_t depth, const double metric_sum, const std::string& theta,
                bool accept) override {
        if (print_) SamplerType::notify(tree, depth, metric_sum, theta, accept);
        if (fn_ && !(depth == 0 && !accept)) {
            buf_.assign(theta_len_, 0.0);
            load_doubles(theta, buf_.data(), theta_len_);
            fn_(user_, depth, accept ? 1 : 0, metric_sum, buf_.data(), theta_len_);
        }
        ++rows_;
        if (depth >= 1 && accept) ++accepted_;
    }
    std::size_t rows() const { return rows_; }
    std::size_t accepted() const { return accepted_; }

  private:
    const Executor& ex_;
    uint32_t theta_len_;
    pf_chain_row_fn fn_;
    void* user_;
    bool print_;
    std::vector<double> buf_;
    std::size_t rows_ = 0, accepted_ = 0;
};

struct JobTrace {
    pf_chain_job_fn fn;
    void* user;
    uint32_t theta_len;
    std::vector<double> buf;
    void job(const JobInfo& in, const ProposalInfo& out) {
        if (!fn) return;
        buf.assign(theta_len, 0.0);
        load_doubles(out.proposal, buf.data(), theta_len);
        fn(user, in.id, in.generation, in.state, in.path.c_str(), in.random_position, out.random_position,
           out.next_random_position, out.log_prior, buf.data(), theta_len);
    }
};

// the frontier evaluator: the GPU engine, or a caller-supplied function with the same contract
struct FrontierEvaluator {
    pf_engine* engine;
    pf_chain_eval_fn fn;
    void* user;
    uint32_t theta_len;
    int status = PF_OK;
    // Row-sharded engine group (pf_engine_peer_attach / pf_engine_comm_init): every rank holds `shards` times fewer
    // items than the chain sees.  The tree's item j is local item j / shards of rank j % shards, so a range of the
    // chain's items that starts and ends at multiples of `shards` is the same LOCAL range on every rank, and the
    // exchange inside the engine adds the ranks' partial sums.
    std::size_t shards = 1;
    bool failed() const { return status != PF_OK; }
    std::vector<double> thetas;
    // where the frontier is assembled: the engine's pinned staging buffer when it fits (DMA'd in place by the eval
    // calls, pf_engine_theta_buffer), else a vector of this evaluator's own
    double* staging(std::size_t nodes) {
        const std::size_t need = nodes * (std::size_t) theta_len;
        if (engine) {
            uint64_t cap = 0;
            double* pinned = pf_engine_theta_buffer(engine, &cap);
            if (pinned && need <= cap) return pinned;
        }
        thetas.assign(need, 0.0);
        return thetas.data();
    }
    void operator()(const std::vector<FrontierJob>& jobs, std::vector<double>& sum, std::vector<double>& sum_sq) {
        double* th = staging(jobs.size());
        for (std::size_t i = 0; i < jobs.size(); ++i)
            load_doubles(jobs[i].theta, &th[i * theta_len], theta_len);
        const int r = engine ? pf_eval_frontier(engine, (uint32_t) jobs.size(), th, sum.data(), sum_sq.data())
                             : fn(user, (uint32_t) jobs.size(), th, sum.data(), sum_sq.data());
        if (r != PF_OK && status == PF_OK) status = r;
    }
    // items [first, first + count) of every node; the normal target walks them in each node's own order
    void range(const std::vector<FrontierJob>& jobs, std::size_t first, std::size_t count, std::vector<double>& sum,
               std::vector<double>& sum_sq) {
        double* th = staging(jobs.size());
        order.assign(jobs.size() * 2, 0);
        for (std::size_t i = 0; i < jobs.size(); ++i) {
            load_doubles(jobs[i].theta, &th[i * theta_len], theta_len);
            order[2 * i] = jobs[i].order_stride;
            order[2 * i + 1] = jobs[i].order_start;
        }
        int r;
        if (engine)
            r = pf_eval_frontier_range(engine, (uint32_t) jobs.size(), th, first / shards, count / shards,
                                       ordered ? order.data() : nullptr, sum.data(), sum_sq.data());
        else
            r = range_fn ? range_fn(user, (uint32_t) jobs.size(), th, first, count, order.data(), sum.data(),
                                    sum_sq.data())
                         : PF_ERR_UNSUPPORTED;
        if (r != PF_OK && status == PF_OK) status = r;
    }
    pf_chain_range_fn range_fn = nullptr;
    bool ordered = false;
    std::vector<uint64_t> order;
};

struct ChainTarget {  // what is being sampled and who scores it
    int model;
    uint32_t dim, components, theta_len;
    std::size_t data_size;
    pf_engine* engine;
    pf_chain_eval_fn fn;
    void* eval_user;
    pf_chain_range_fn range_fn;
    std::size_t shards = 1;  // ranks of a row-sharded engine group (1: the engine holds the whole data set)
};

template <class Executor>
static int run_chain(const ChainTarget& tg, Executor& ex, const pf_chain_config& cfg, pf_chain_row_fn row,
                     pf_chain_job_fn job, void* user, pf_chain_stats* stats) {
    const int workers = cfg.frontier ? (int) cfg.frontier : 1;
    const std::size_t data_size = tg.data_size;
    const uint32_t theta_len = tg.theta_len;
    ChainHook<Executor> hook(ex, theta_len, row, user, cfg.print_tsv != 0);
    SpecTree tree(workers + 1, (std::size_t) cfg.chain_length, &hook, cfg.scheduling_policy,
                  cfg.correlation > 0 ? cfg.correlation : 0.99, cfg.dimension ? cfg.dimension : 1);
    tree.set_reassign_ratio(cfg.reassign_ratio);  // tree.cc:147-151
    tree.set_require_comparison_theta(false);
    tree.set_threshold(cfg.threshold != 0);
    if (cfg.print_tsv) hook.SamplerType::notify(tree, 0, 0, std::string(), false);  // the header line

    FrontierEvaluator eval{tg.engine, tg.fn, tg.eval_user, theta_len};
    eval.shards = tg.shards;
    eval.range_fn = tg.range_fn;
    eval.ordered = tg.model == PF_MODEL_NORMAL;
    JobTrace trace{job, user, theta_len, {}};
    FrontierCommunicator<SpecTree, Executor, HostTypes, FrontierEvaluator, JobTrace> comm(
        tree, ex, eval, workers, data_size, &trace, cfg.deferred_proposals == 0,
        // partial updates of a sharded group cover whole rounds of the shards (see FrontierEvaluator::shards)
        ((std::size_t) cfg.batch_items + tg.shards - 1) / tg.shards * tg.shards, (int) cfg.update_style);
    SpecTree::run_behavior rb;
    rb.quiet = cfg.verbose == 0;
    const auto t0 = std::chrono::steady_clock::now();
    tree.run_master(comm, rb);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stats) {
        stats->levels = tree.completed_depth();
        stats->engine_calls = comm.stats().engine_calls;
        stats->nodes_scored = comm.stats().nodes_scored;
        stats->max_frontier = comm.stats().max_batch;
        stats->seconds = dt;
        stats->eval_seconds = comm.stats().eval_seconds;
        stats->allocated_nodes = tree.allocated_nodes();
        stats->recycled_nodes = tree.recycled_nodes();
        stats->rows = hook.rows();
        stats->items_scored = comm.stats().items_scored;
        stats->abandoned = comm.stats().abandoned + comm.stats().skipped;
        stats->accepted = hook.accepted();
        stats->min_margin = tree.min_margin();
        stats->min_margin_rel = tree.min_margin_rel();
    }
    return eval.status;
}

} }  // namespace pf::host

static int dispatch_chain(const pf::host::ChainTarget& tg, const pf_chain_config* cfg, pf_chain_row_fn row,
                          pf_chain_job_fn job, void* user, pf_chain_stats* stats) {
    using namespace pf::host;
    switch (tg.model) {
        case PF_MODEL_NORMAL: {
            if (tg.dim != 2) return PF_ERR_UNSUPPORTED;  // the reference plugin is NormalSampler<2> (normal.cc:40)
            NormalExecutor<2> ex(tg.data_size);
            return run_chain(tg, ex, *cfg, row, job, user, stats);
        }
        case PF_MODEL_BOLTZMANN_SCALAR: {
            BoltzmannScalarExecutor ex(tg.data_size);
            return run_chain(tg, ex, *cfg, row, job, user, stats);
        }
        case PF_MODEL_BOLTZMANN_DENSE: {
            BoltzmannDenseExecutor ex((int) tg.dim);
            return run_chain(tg, ex, *cfg, row, job, user, stats);
        }
        case PF_MODEL_MIXTURE: {
            MixtureExecutor ex(tg.data_size, (int) tg.components, (int) tg.dim, cfg->seed);
            return run_chain(tg, ex, *cfg, row, job, user, stats);
        }
        case PF_MODEL_BLASSO: {
            BLassoExecutor ex(tg.data_size, (int) tg.dim, cfg->tau > 0 ? cfg->tau : 5.0, cfg->initial_theta);
            return run_chain(tg, ex, *cfg, row, job, user, stats);
        }
    }
    return PF_ERR_INVALID;
}

extern "C" PF_API int pf_chain_run(pf_engine* engine, const pf_chain_config* cfg, pf_chain_row_fn row,
                                   pf_chain_job_fn job, void* user, pf_chain_stats* stats) {
    if (!engine || !cfg) return PF_ERR_INVALID;
    pf::host::ChainTarget tg{pf_engine_model(engine), pf_engine_dim(engine), pf_engine_components(engine),
                             pf_engine_theta_len(engine), (std::size_t) pf_engine_data_size(engine), engine, nullptr,
                             nullptr, nullptr};
    int rank = 0, nranks = 1, node_split = 0;
    pf_engine_group(engine, &rank, &nranks, &node_split);
    if (nranks > 1 && !node_split) {
        // Row shards: the chain samples the posterior of the WHOLE data set, so it sees nranks x the local item count
        // (every rank must hold the same number of items -- pfetch.h -- or the ranks' trees would diverge).
        if (tg.model == PF_MODEL_NORMAL && cfg->batch_items) {
            pf::set_error("pf_chain_run: StrideIndex-ordered partial ranges address the whole data set; not available "
                                 "on a row shard (use batch_items = 0)");
            return PF_ERR_UNSUPPORTED;
        }
        tg.shards = (std::size_t) nranks;
        tg.data_size *= (std::size_t) nranks;
    }
    return dispatch_chain(tg, cfg, row, job, user, stats);
}

extern "C" PF_API int pf_chain_run_with_evaluator(int model, uint32_t dim, uint32_t components, uint64_t data_size,
                                                  const pf_chain_config* cfg, pf_chain_eval_fn eval,
                                                  pf_chain_range_fn eval_range, void* eval_user, pf_chain_row_fn row,
                                                  pf_chain_job_fn job, void* user, pf_chain_stats* stats) {
    if (!cfg || !eval || data_size == 0) return PF_ERR_INVALID;
    uint32_t theta_len = 0;
    switch (model) {
        case PF_MODEL_NORMAL: theta_len = dim + dim * (dim + 1) / 2; break;
        case PF_MODEL_BOLTZMANN_SCALAR: theta_len = 1; break;
        case PF_MODEL_BOLTZMANN_DENSE: theta_len = dim; break;
        case PF_MODEL_MIXTURE: theta_len = dim * components; break;
        case PF_MODEL_BLASSO: theta_len = dim + 1; break;
        default: return PF_ERR_INVALID;
    }
    if (cfg->batch_items && !eval_range) return PF_ERR_INVALID;
    pf::host::ChainTarget tg{model, dim, components, theta_len, (std::size_t) data_size, nullptr, eval, eval_user,
                             eval_range};
    return dispatch_chain(tg, cfg, row, job, user, stats);
}
