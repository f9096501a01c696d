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
};

struct FrontierStats {
    std::size_t engine_calls = 0, nodes_scored = 0, max_batch = 0, jobs = 0;
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
                         std::size_t data_size, Trace* trace, bool eager_proposals = true)
        : tree_(tree), executors_(workers + 1, executor), evaluator_(evaluator), trace_(trace), data_size_(data_size),
          eager_(eager_proposals), inbox_(workers + 1), proposals_(workers + 1), updates_(workers + 1) {
        for (int r = 1; r <= workers; ++r) {  // worker.hh:66,74
            push(r, FC_REGISTER);
            push(r, FC_WANTWORK);
        }
    }

    int size() const { return (int) inbox_.size(); }
    const FrontierStats& stats() const { return stats_; }

    std::optional<Status> iprobe(int rank, int) {
        if (queued_ == 0 && !batch_.empty()) flush();
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
            push(rank, FC_QUIT);
            return Request();
        }
        JobInfoT job(job_in);
        ProposalInfoT proposal;
        const auto t0 = std::chrono::steady_clock::now();
        setup_job(executors_[rank], job, proposal);
        proposal.proposal_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (trace_) trace_->job(job_in, proposal);
        if (job_in.state != JobInfoT::s_has_proposal) {
            if (eager_) tree_.set_proposal(proposal, rank);
            proposals_[rank].push_back(proposal);
            push(rank, FC_SETPROPOSAL);
        }
        batch_.push_back(FrontierJob{rank, job.id, job.generation, proposal.proposal});
        ++stats_.jobs;
        return Request();
    }

    // Nodes are scored whole, so there is never work in progress to abandon.
    Request isend(int, int, const AbandonInfoT&) { return Request(); }

  private:
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
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        stats_.eval_seconds += dt;
        ++stats_.engine_calls;
        stats_.nodes_scored += n;
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
    std::vector<std::deque<int>> inbox_;
    std::vector<std::deque<ProposalInfoT>> proposals_;
    std::vector<std::deque<MetricChangeT>> updates_;
    std::vector<FrontierJob> batch_;
    std::vector<double> sum_, sum_sq_;
    std::size_t queued_ = 0;
    FrontierStats stats_;
};

} }  // namespace pf::host
