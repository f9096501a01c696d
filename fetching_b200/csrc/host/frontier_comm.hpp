// host/frontier_comm.hpp -- the Communicator that turns a scheduling burst into ONE engine call.
//
// It satisfies the Communicator concept of the master loop (masterloop.cc:21-22; the reference's
// other in-process instance is MasterTester, testmaster.cc:57-293) for W virtual workers:
//   * isend(HAVEWORK, JobInfo): materialise the proposal on the host right away (setup_job,
//     worker.hh:28-49), tell the tree (set_proposal) and queue the SETPROPOSAL message, and put
//     the node on the current FRONTIER BATCH -- nothing is evaluated yet;
//   * iprobe(): when no message is left to deliver, the burst is over: score the whole batch
//     with one evaluator call (the GPU engine: pf_eval_frontier) and queue, per node, one
//     whole-node UPDATE  MetricChange{id, 0, data_size, sum, sum_sq, ..., abandon = true}
//     (accepted by the tree: master.cc:161-163,169) followed by WANTWORK.
// Deterministic: per-rank FIFO inboxes, batch scored in hand-out order.
//
// BATCHED MODE (batch_items > 0): the predictive part of predictive prefetching.  Like a reference
// worker (worker.hh:130-175) every virtual worker walks its node in batches of items and reports the
// CUMULATIVE partial sums after each batch, so the tree can estimate branch probabilities from partial
// evidence (master.cc:97-142), stop speculating down improbable branches, and pull workers off
// nodes that became useless (ABANDON, masterloop.cc:70-82).  Each flush advances every active
// worker by one batch: one evaluator.range() call per distinct item range.  Update styles as in
// worker.hh:141-147 (FAST_SLOW: one probe batch, then the rest; INCREMENTAL: fixed batches; LOGARITHMIC: like
// FAST_SLOW, and the worker's probe batch doubles with every job it starts from item 0).
//
// FAILURE.  An evaluator that exposes `bool failed() const` (the GPU engine's does) is asked after every call; on
// the first failure the batch's UPDATEs are dropped -- their sums are not results -- every inbox is emptied and
// each virtual worker answers QUIT, so run_master returns at once instead of accepting every proposal on metric 0.
//
// Generic over the tree / record types on purpose: the oracle build instantiates this same
// header with the REFERENCE's JobTree and records and a CPU evaluator to produce the golden
// traces that the product (SpecTree + GPU engine) must reproduce bit for bit.
#pragma once
#include <chrono>
#include <cstddef>
#include <deque>
#include <optional>
#include <string>
#include <utility>
#include <vector>

#include "setup_job.hpp"

namespace pf { namespace host {

struct FrontierJob {  // one node of the batch handed to the evaluator
    int rank;
    std::size_t id, generation;
    std::string theta;  // raw doubles
    std::size_t order_stride = 1, order_start = 0;  // the node's item visiting order, when the sampler has one
};

struct FrontierStats {
    std::size_t engine_calls = 0, nodes_scored = 0, max_batch = 0, jobs = 0;
    std::size_t items_scored = 0;  // sum over calls of nodes x items: the work actually paid for
    std::size_t abandoned = 0, skipped = 0;
    double eval_seconds = 0.0;
};

// message tags shared with the reference (protocol.hh:32-41)
enum { FC_REGISTER = 2, FC_WANTWORK = 3, FC_HAVEWORK = 4, FC_SETPROPOSAL = 5, FC_UPDATE = 6, FC_ABANDON = 7, FC_QUIT = 8 };

template <class Tree, class Executor, class Types, class Evaluator, class Trace>
class FrontierCommunicator {
  public:
    using JobInfoT = typename Types::JobInfo;
    using ProposalInfoT = typename Types::ProposalInfo;
    using MetricChangeT = typename Types::MetricChange;
    using AbandonInfoT = typename Types::AbandonInfo;

    class Status {
      public:
        Status(int source, int tag) : source_(source), tag_(tag) {}
        int source() const { return source_; }
        int tag() const { return tag_; }

      private:
        int source_, tag_;
    };
    struct Request {
        bool test() const { return true; }
    };

    // `eager_proposals`: tell the tree about a proposal the moment it is made (like MasterTester,
    // testmaster.cc:284-287), so later hand-outs of the same burst already see their ancestors'
    // thetas.  With false the tree only learns it when the master reads the SETPROPOSAL message
    // (like real workers), and jobs carry accept/reject paths to be replayed (worker.hh:38-47).
    // Every virtual worker owns a copy of the executor, as every reference worker does.
    FrontierCommunicator(Tree& tree, const Executor& executor, Evaluator& evaluator, int workers,
                         std::size_t data_size, Trace* trace, bool eager_proposals = true,
                         std::size_t batch_items = 0, int update_style = 1)
        : tree_(tree), executors_(workers + 1, executor), evaluator_(evaluator), trace_(trace), data_size_(data_size),
          eager_(eager_proposals), batch_items_(batch_items >= data_size ? 0 : batch_items),
          update_style_(update_style), inbox_(workers + 1), proposals_(workers + 1), updates_(workers + 1),
          work_(workers + 1) {
        for (Work& w : work_) w.batch = batch_items_;
        for (int r = 1; r <= workers; ++r) {  // worker.hh:66,74
            push(r, FC_REGISTER);
            push(r, FC_WANTWORK);
        }
    }

    int size() const { return (int) inbox_.size(); }
    const FrontierStats& stats() const { return stats_; }
    bool failed() const { return failed_; }

    std::optional<Status> iprobe(int rank, int) {
        if (queued_ == 0 && !failed_) {
            if (batch_items_ == 0) {
                if (!batch_.empty()) flush();
            } else
                advance();
        }
        if (inbox_[rank].empty()) return std::nullopt;
        return Status(rank, inbox_[rank].front());
    }

    void recv(int rank, int) { pop(rank); }
    void recv(int rank, int, std::size_t& value) {
        value = data_size_;
        pop(rank);
    }
    void recv(int rank, int, ProposalInfoT& value) {
        value = std::move(proposals_[rank].front());
        proposals_[rank].pop_front();
        pop(rank);
    }
    void recv(int rank, int, MetricChangeT& value) {
        value = updates_[rank].front();
        updates_[rank].pop_front();
        pop(rank);
    }

    Request isend(int rank, int, const JobInfoT& job_in) {
        if (job_in.id == 0) {  // the stop job (masterloop.cc:46-49); a worker answers QUIT (worker.hh:79-80)
            if (!failed_) push(rank, FC_QUIT);
            return Request();
        }
        if (failed_) return Request();  // the QUITs are already queued: hand-outs of the dying burst go nowhere
        JobInfoT job(job_in);
        ProposalInfoT proposal;
        const auto t0 = std::chrono::steady_clock::now();
        setup_job(executors_[rank], job, proposal);
        proposal.proposal_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (trace_) trace_->job(job_in, proposal);
        FrontierJob fj{rank, job.id, job.generation, proposal.proposal};
        order_of(executors_[rank], fj, 0);
        if (job_in.state != JobInfoT::s_has_proposal) {
            if (eager_) tree_.set_proposal(proposal, rank);
            proposals_[rank].push_back(std::move(proposal));  // (theta is kilobytes for the big models: move, not copy)
            push(rank, FC_SETPROPOSAL);
        }
        ++stats_.jobs;
        if (batch_items_ == 0)
            batch_.push_back(std::move(fj));
        else {  // worker.hh:130-135: cumulative sums continue from what the tree already has
            Work& w = work_[rank];
            w.active = true;
            w.job = std::move(fj);
            w.position = job_in.first_position;
            w.sum = job_in.metric_sum;
            w.sum_sq = job_in.metric_sum_sq;
        }
        return Request();
    }

    // masterloop.cc:76-82 -> worker.hh:169-175: drop the node if it is still the one being worked on
    Request isend(int rank, int, const AbandonInfoT& ab) {
        if (batch_items_ != 0) {
            Work& w = work_[rank];
            if (w.active && w.job.id == ab.id && w.job.generation == ab.generation) {
                w.active = false;
                ++stats_.abandoned;
                push(rank, FC_WANTWORK);
            }
        }
        return Request();
    }

  private:
    struct Work {  // what a virtual worker is in the middle of (batched mode)
        bool active = false;
        FrontierJob job;
        std::size_t position = 0, batch = 0;
        double sum = 0.0, sum_sq = 0.0;
    };

    // executors that define a visiting order expose order(); the others are walked in natural order
    template <class E>
    static auto order_of(const E& ex, FrontierJob& fj, int) -> decltype(ex.order().stride(), void()) {
        fj.order_stride = ex.order().stride();
        fj.order_start = ex.order().start();
    }
    template <class E>
    static void order_of(const E&, FrontierJob&, long) {}

    // One step of every active worker (batched mode): its next item range, grouped so that workers at the
    // same range are scored by one evaluator call.
    void advance() {
        struct Group {
            std::size_t first, count;
            std::vector<int> ranks;
        };
        std::vector<Group> groups;
        for (int r = 1; r < (int) work_.size(); ++r) {
            Work& w = work_[r];
            if (!w.active) continue;
            // a node the tree has already deleted or decided against is not worth more work
            // (MasterTester makes the same check, testmaster.cc:170); the worker asks for new work
            auto* n = tree_.node(w.job.id);
            if (!n || n->is_dead() || n->complete()) {
                w.active = false;
                ++stats_.skipped;
                push(r, FC_WANTWORK);
                continue;
            }
            std::size_t stop;
            if (update_style_ == 1 || w.position == 0) {  // worker.hh:141-147
                stop = w.position + w.batch;
                if (update_style_ == 2) w.batch *= 2;  // LOGARITHMIC: this WORKER's probe batch doubles per job
            } else
                stop = data_size_;
            if (stop > data_size_) stop = data_size_;
            const std::size_t count = stop - w.position;
            Group* g = nullptr;
            for (Group& c : groups)
                if (c.first == w.position && c.count == count) g = &c;
            if (!g) {
                groups.push_back(Group{w.position, count, {}});
                g = &groups.back();
            }
            g->ranks.push_back(r);
        }
        for (Group& g : groups) {
            batch_.clear();
            for (int r : g.ranks) batch_.push_back(work_[r].job);
            const std::size_t n = batch_.size();
            sum_.assign(n, 0.0);
            sum_sq_.assign(n, 0.0);
            const auto t0 = std::chrono::steady_clock::now();
            evaluator_.range(batch_, g.first, g.count, sum_, sum_sq_);
            if (evaluator_failed(evaluator_, 0)) return fail();
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            stats_.eval_seconds += dt;
            ++stats_.engine_calls;
            stats_.nodes_scored += n;
            stats_.items_scored += n * g.count;
            if (n > stats_.max_batch) stats_.max_batch = n;
            for (std::size_t i = 0; i < n; ++i) {
                const int r = g.ranks[i];
                Work& w = work_[r];
                w.sum += sum_[i];
                w.sum_sq += sum_sq_[i];
                w.position = g.first + g.count;
                MetricChangeT u;
                u.id = w.job.id;
                u.first_position = 0;  // cumulative since item 0 (worker.hh:131)
                u.last_position = w.position;
                u.metric_sum = w.sum;
                u.metric_sum_sq = w.sum_sq;
                u.step_time = dt / (double) n;
                u.step_count = g.count;
                u.abandon = w.position == data_size_;
                u.worker_wait_time = 0.0;
                updates_[r].push_back(u);
                push(r, FC_UPDATE);
                if (u.abandon) {
                    w.active = false;
                    push(r, FC_WANTWORK);
                }
            }
        }
        batch_.clear();
    }

    template <class E>
    static auto evaluator_failed(const E& ev, int) -> decltype(ev.failed()) { return ev.failed(); }
    template <class E>
    static bool evaluator_failed(const E&, long) { return false; }

    // the evaluator reported an error: nothing it returned is a result.  Drop everything in flight and let every
    // virtual worker QUIT, which ends run_master (masterloop.cc:89-98) without another level being decided.
    void fail() {
        failed_ = true;
        batch_.clear();
        for (Work& w : work_) w.active = false;
        queued_ = 0;
        for (int r = 1; r < (int) inbox_.size(); ++r) {
            inbox_[r].clear();
            proposals_[r].clear();
            updates_[r].clear();
            push(r, FC_QUIT);
        }
    }

    void push(int rank, int tag) {
        inbox_[rank].push_back(tag);
        ++queued_;
    }
    void pop(int rank) {
        inbox_[rank].pop_front();
        --queued_;
    }

    void flush() {
        const std::size_t n = batch_.size();
        sum_.assign(n, 0.0);
        sum_sq_.assign(n, 0.0);
        const auto t0 = std::chrono::steady_clock::now();
        evaluator_(batch_, sum_, sum_sq_);
        if (evaluator_failed(evaluator_, 0)) return fail();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        stats_.eval_seconds += dt;
        ++stats_.engine_calls;
        stats_.nodes_scored += n;
        stats_.items_scored += n * data_size_;
        if (n > stats_.max_batch) stats_.max_batch = n;
        for (std::size_t i = 0; i < n; ++i) {
            MetricChangeT u;
            u.id = batch_[i].id;
            u.first_position = 0;
            u.last_position = data_size_;
            u.metric_sum = sum_[i];
            u.metric_sum_sq = sum_sq_[i];
            u.step_time = dt / (double) n;
            u.step_count = data_size_;
            u.abandon = true;
            u.worker_wait_time = 0.0;
            updates_[batch_[i].rank].push_back(u);
            push(batch_[i].rank, FC_UPDATE);
            push(batch_[i].rank, FC_WANTWORK);
        }
        batch_.clear();
    }

    Tree& tree_;
    std::vector<Executor> executors_;
    Evaluator& evaluator_;
    Trace* trace_;
    std::size_t data_size_;
    bool eager_;
    std::size_t batch_items_;
    int update_style_;
    std::vector<std::deque<int>> inbox_;
    std::vector<std::deque<ProposalInfoT>> proposals_;
    std::vector<std::deque<MetricChangeT>> updates_;
    std::vector<FrontierJob> batch_;
    std::vector<Work> work_;
    std::vector<double> sum_, sum_sq_;
    std::size_t queued_ = 0;
    bool failed_ = false;
    FrontierStats stats_;
};

} }  // namespace pf::host
