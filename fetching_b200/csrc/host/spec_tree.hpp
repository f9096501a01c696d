/* Derived from Fetching (elaine84/fetching): master.hh (JobNode / JobTree declarations).
 *
 * Fetching
 * Elaine Angelino, Eddie Kohler
 * Copyright (c) 2012-2014 President and Fellows of Harvard College
 *
 * Permission is hereby granted, free of charge, to any person obtaining a
 * copy of this software and associated documentation files (the "Software"),
 * to deal in the Software without restriction, subject to the conditions
 * listed in the Fetching LICENSE file. These conditions include: you must
 * preserve this copyright notice, and you cannot mention the copyright
 * holders in advertising related to the Software without their permission.
 * The Software is provided WITHOUT ANY WARRANTY, EXPRESS OR IMPLIED. This
 * notice is a summary of the Fetching LICENSE file; the license in that file
 * is legally binding.  (The licence text is reproduced in LICENSE.fetching
 * next to this file.)
 */
// host/spec_tree.hpp -- the speculative (prefetching) Metropolis-Hastings tree.
//
// JobTree / JobNode of the reference (master.hh, master.cc) under other names -- a derived work, see the
// notice above: same node semantics, same id allocation, same random draws in the same order, same
// accept test, same trimming -- so that, driven by the same message sequence, it hands out the
// same JobInfos and prints the same chain.  (Parity is tested against traces produced by the
// reference's own tree: tests/golden/frontier_trace_*.tsv.)
//
// A node is "the proposal made at MH step `depth`, given the accept/reject path from the root"
// (master.hh:170-249).  child[1] is the ACCEPT child, child[0] the REJECT child; the comparison
// parent is the node holding the state the proposal is compared with (master.cc:41).  One
// log_u per level, drawn lazily (master.cc:193-200).  Node ids are handed out in creation order.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <deque>
#include <iosfwd>
#include <limits>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "protocol.hpp"
#include "rng.hpp"

namespace pf { namespace host {

class SamplerType;

enum SchedulingPolicy { NAIVE = 0, CONSTANT = 1, PREDICTIVE = 2 };  // master.hh:29-33

class SpecTree;

class SpecNode {
  public:
    bool is_root() const { return parent_ == nullptr; }
    bool is_accept() const { return parent_ && this == parent_->child_[1]; }
    std::size_t id() const { return id_; }
    std::size_t generation() const { return generation_; }
    std::size_t depth() const { return depth_; }
    std::size_t true_depth() const { return true_depth_; }
    bool has_theta() const { return has_theta_; }
    const std::string& theta() const { return theta_; }
    std::size_t position() const { return position_; }
    std::size_t data_size() const { return ndata_; }
    std::size_t random_position() const { return random_position_; }
    std::size_t next_random_position() const { return next_random_position_; }
    double metric_sum() const { return metric_sum_; }
    double metric_sum_sq() const { return metric_sum_sq_; }
    double scale() const { return scale_; }
    double log_u() const { return log_u_; }
    double branch_probability() const { return branch_probability_; }
    bool pending() const { return executor_ == 0; }
    bool complete() const { return position_ == ndata_; }
    bool has_final_branch_probability() const { return branch_probability_ == 0.0 || branch_probability_ == 1.0; }
    bool is_dead() const { return utility() == 0.0; }
    SpecNode* child(bool accept) const { return child_[accept]; }
    double utility() const;

  private:
    friend class SpecTree;
    SpecNode(std::size_t ndata, std::size_t policy, double correlation, std::size_t dimension);  // the true root
    SpecNode(std::size_t id, SpecNode* parent, bool accept_side, double log_u, std::size_t ndata,
             double local_accept_rate, std::size_t policy, double correlation);

    double mean() const { return metric_sum_ / position_; }
    double variance() const;
    double trace_mean(std::size_t i) const { return trace_sum_[i] / trace_position_[i]; }
    double trace_variance(std::size_t i) const;
    std::size_t closest_trace(std::size_t pos) const;
    void update_branch_probability(bool threshold, double local_accept_rate);
    void absorb(const MetricChange& u, bool threshold, double local_accept_rate);

    std::size_t id_ = 1, generation_ = 0, depth_ = 0, true_depth_ = 0;
    SpecNode* child_[2] = {nullptr, nullptr};
    SpecNode* parent_ = nullptr;
    SpecNode* comparison_parent_ = nullptr;
    bool has_theta_ = false;
    std::string theta_;
    std::size_t random_position_ = 0, next_random_position_ = 0;
    std::size_t ndata_ = 0, position_ = 0;
    double scale_ = 0.0, log_prior_ = 0.0, metric_sum_ = 0.0, metric_sum_sq_ = 0.0;
    std::deque<std::size_t> trace_position_;
    std::deque<double> trace_sum_, trace_sum_sq_;
    double log_u_ = 0.0, branch_probability_ = 1.0;
    int executor_ = -1, last_executor_ = 0;
    double proposal_time_ = 0.0, step_time_ = 0.0;
    std::size_t step_count_ = 0;
    int flip_count_ = 0, threshold_position_ = 0;
    std::size_t policy_ = PREDICTIVE;
    double correlation_ = 0.99;
};

class SpecTree {
  public:
    using node_type = SpecNode;

    struct run_behavior {
        bool print_sending_work = false;
        bool quiet = true;
    };

    SpecTree(int world_size, std::size_t max_depth = std::numeric_limits<std::size_t>::max(),
             SamplerType* completion_hook = nullptr, std::size_t scheduling_policy = PREDICTIVE,
             double correlation = 0.99, std::size_t dimension = 1);
    ~SpecTree();
    SpecTree(const SpecTree&) = delete;
    SpecTree& operator=(const SpecTree&) = delete;

    void set_data_size(std::size_t data_size);
    void set_require_comparison_theta(bool x);
    void set_reassign_ratio(double r) { reassign_ratio_ = r; }
    void set_threshold(bool x) { threshold_ = x; }

    const SpecNode* root() const { return root_; }
    SpecNode* node(std::size_t id);
    const SpecNode* node(std::size_t id) const { return const_cast<SpecTree*>(this)->node(id); }
    SamplerType* sampler() const { return hook_; }
    std::size_t completed_depth() const { return first_depth_; }
    std::size_t max_depth() const { return max_depth_; }
    std::size_t allocated_nodes() const { return allocated_nodes_; }
    std::size_t recycled_nodes() const { return recycled_nodes_; }
    double local_accept_rate() const { return accept_count_ / 100.; }
    double reassign_ratio() const { return reassign_ratio_; }
    // smallest |delta_log_prior + metric_sum - parent.metric_sum - log_u| over the levels decided so far, and the
    // smallest such margin relative to max(|metric_sum|, |parent.metric_sum|): how far the chain's accept / reject
    // decisions were from flipping under an error in the sums (infinity before the first decision)
    double min_margin() const { return min_margin_; }
    double min_margin_rel() const { return min_margin_rel_; }

    std::pair<std::size_t, double> high_utility_pending(std::size_t stopper = 0);
    JobInfo execute(std::size_t id, int executor);
    void set_proposal(const ProposalInfo& proposal, int executor);
    void create_children(SpecNode* parent) {
        create_child(parent, false);
        create_child(parent, true);
    }
    void set_metric(const MetricChange& update, const run_behavior& rb, double master_wait_time,
                    double master_work_time, int rank, std::size_t iprobe_status, std::size_t iprobe_no_status);
    void abandon(std::size_t id, int executor);

    // The master event loop (masterloop.cc:21-134), generic over the Communicator concept.
    template <typename Communicator>
    void run_master(Communicator& world, const run_behavior& rb);

    void print(std::ostream& s) const;

  private:
    std::size_t next_id() const { return first_id_ + by_id_.size(); }
    void create_child(SpecNode* parent, bool accept_side);
    void delete_subtree(SpecNode* n);
    void delete_loser_child(SpecNode* n);
    void trim(const run_behavior& rb, double master_wait_time, double master_work_time, std::size_t iprobe_status,
              std::size_t iprobe_no_status);

    SpecNode* root_ = nullptr;
    Mt11213b utility_generator_;  // work selection walk (master.hh:129)
    Mt19937 u_generator_;         // one log(u) per level (master.hh:131)
    std::deque<double> level_log_u_;
    std::size_t first_depth_ = 0;
    std::deque<SpecNode*> by_id_;
    std::size_t first_id_ = 1;
    std::size_t ndata_ = 0;
    bool require_comparison_theta_ = true, threshold_ = true;
    std::size_t allocated_nodes_ = 0, recycled_nodes_ = 0, max_depth_;
    double reassign_ratio_ = 0.0;
    SamplerType* hook_;
    bool own_hook_ = false;
    std::size_t policy_;
    double correlation_;
    double accept_count_ = 50.0;
    std::deque<bool> accept_window_;
    std::size_t dimension_;
    double wasted_proposal_time_ = 0.0, wasted_step_time_ = 0.0, useful_proposal_time_ = 0.0,
           useful_step_time_ = 0.0;
    std::vector<double> wait_time_;
    double min_margin_ = std::numeric_limits<double>::infinity(),
           min_margin_rel_ = std::numeric_limits<double>::infinity();
};

std::ostream& operator<<(std::ostream& s, const SpecTree& tree);

} }  // namespace pf::host

#include "master_loop.hpp"
