/* Derived from Fetching (elaine84/fetching): master.cc (JobNode / JobTree), function for function, with the names changed.
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
// host/spec_tree.cpp -- see spec_tree.hpp.  A DERIVED WORK of the reference's master.cc (notice above): the
// same functions in the same order under other names, kept in-tree so that pf_chain_run works where
// /root/reference does not exist (the GPU box).  The reference's own JobTree drives the same engine through the
// same FrontierCommunicator in oracle/ref_driver.cc (`pf_ref frontier --engine gpu`), which is the drop-in proof.
#include "spec_tree.hpp"

#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <sstream>

#include "sampler_type.hpp"

namespace pf { namespace host {

// ------------------------------------------------------------------------------------------- SpecNode
// the true root: master.cc:19-33.  Initial proposal scale log(2.38^2 / dimension).
SpecNode::SpecNode(std::size_t ndata, std::size_t policy, double correlation, std::size_t dimension)
    : ndata_(ndata), scale_(std::log(2.38 * 2.38 / (double) dimension)), policy_(policy), correlation_(correlation) {
    assert(ndata_ != 0);
}

// a child: master.cc:35-53.  The adaptive scale moves with the branch taken:
// scale = parent.scale + (accept - 0.234) / sqrt(parent.depth + 2).
SpecNode::SpecNode(std::size_t id, SpecNode* parent, bool accept_side, double log_u, std::size_t ndata,
                   double local_accept_rate, std::size_t policy, double correlation)
    : id_(id), depth_(parent->depth_ + 1), true_depth_(parent->depth_ + 1), parent_(parent),
      comparison_parent_(accept_side ? parent : parent->comparison_parent_), theta_(comparison_parent_->theta_),
      ndata_(ndata), scale_(parent->scale_ + 1.0 / std::sqrt(parent->depth_ + 2.0) * ((double) accept_side - 0.234)),
      log_u_(log_u), branch_probability_(local_accept_rate), policy_(policy), correlation_(correlation) {
    assert(ndata_ == parent->ndata_);
    parent->child_[accept_side] = this;
}

// probability that this node lies on the true chain, given what is known (master.cc:55-75)
double SpecNode::utility() const {
    double u = 1.0;
    for (const SpecNode* n = this; n->parent_; n = n->parent_) {
        double p = n->parent_->branch_probability_;
        if (!n->is_accept()) p = 1.0 - p;
        if (n->parent_->complete()) {
            if (p < 0.5) return 0.0;  // decided against us: dead
        } else {
            u *= p;
            if (u <= DBL_MIN) u = DBL_MIN;  // 0 is reserved for "truly dead"
        }
    }
    return u;
}

static double clamp_variance(double v) {
    assert(v > -0.1);
    return v < 0.0 ? 0.0 : v;
}

double SpecNode::variance() const {  // master.hh:309-316
    return clamp_variance((metric_sum_sq_ - (metric_sum_ * metric_sum_) / position_) / position_);
}

double SpecNode::trace_variance(std::size_t i) const {  // master.hh:352-359
    return clamp_variance((trace_sum_sq_[i] - (trace_sum_[i] * trace_sum_[i]) / trace_position_[i]) /
                          trace_position_[i]);
}

// index of the recorded partial sum whose position is nearest to `pos` (master.hh:327-338)
std::size_t SpecNode::closest_trace(std::size_t pos) const {
    auto dist = [&](std::size_t i) {
        return std::abs((int) (trace_position_[i] - pos));  // the reference narrows to int, too
    };
    int prev = dist(0);
    for (std::size_t i = 1; i < trace_position_.size(); ++i) {
        const int d = dist(i);
        if (d >= prev) return i - 1;
        prev = d;
    }
    return trace_position_.size() - 1;
}

// master.cc:77-153.  Once this node and its comparison parent are both complete the MH accept
// test is final:  delta_log_prior + metric_sum - parent.metric_sum > log_u  (master.cc:89).
void SpecNode::update_branch_probability(bool threshold, double local_accept_rate) {
    const double before = branch_probability_;
    const SpecNode* cp = comparison_parent_;
    const double delta_prior = cp ? log_prior_ - cp->log_prior_ : 0.0;

    if (!cp)
        branch_probability_ = 1.0;  // the root has nothing to be compared with
    else if (complete() && cp->complete())
        branch_probability_ = (delta_prior + metric_sum() - cp->metric_sum() > log_u_) ? 1.0 : 0.0;
    else if (policy_ == NAIVE)
        branch_probability_ = 0.5;
    else if (policy_ == CONSTANT)
        branch_probability_ = local_accept_rate;
    else if (policy_ == PREDICTIVE) {
        if (threshold) {
            if (position_ == 0 || cp->position_ == 0)
                branch_probability_ = local_accept_rate;
            else if (delta_prior + ndata_ * (mean() - cp->mean()) > log_u_)
                branch_probability_ = 0.999999;
            else
                branch_probability_ = 0.000001;
        } else if (position_ < 2 || cp->position_ < 2) {
            branch_probability_ = local_accept_rate;  // no variance estimate yet
        } else {
            // normal approximation of the final log ratio from the partial sums (master.cc:111-141)
            double mm, mv, cmm, cmv;
            if (cp->position_ == position_) {
                mm = mean(), mv = variance(), cmm = cp->mean(), cmv = cp->variance();
            } else if (cp->position_ > position_) {
                const std::size_t i = cp->closest_trace(position_);
                mm = mean(), mv = variance(), cmm = cp->trace_mean(i), cmv = cp->trace_variance(i);
            } else {
                const std::size_t i = closest_trace(cp->position_);
                mm = trace_mean(i), mv = trace_variance(i), cmm = cp->mean(), cmv = cp->variance();
            }
            const double estimate = delta_prior + ndata_ * (mm - cmm);
            const double error = (ndata_ - position_) *
                                 (1. / std::sqrt(position_) + 1. / std::sqrt(ndata_ - position_)) *
                                 std::sqrt(mv + cmv - 2 * correlation_ * std::sqrt(mv * cmv));
            double p = 0.5 + 0.5 * std::erf((estimate - log_u_) / std::sqrt(2) / error);
            if (p < 0.000001) p = 0.000001;  // final values 0 / 1 are reserved for complete nodes
            if (p > 0.999999) p = 0.999999;
            branch_probability_ = p;
        }
    }

    if ((before < 0.5 && branch_probability_ > 0.5) || (before > 0.5 && branch_probability_ < 0.5)) ++flip_count_;
    const double delta = 0.1;
    if ((before < 1.0 - delta && branch_probability_ > 1.0 - delta) || (before > delta && branch_probability_ < delta))
        threshold_position_ = (int) position_;
}

// master.cc:155-182: take a (cumulative) partial result, then refresh the branch probability of
// this node and of every node compared against it (its accept child and that child's reject chain).
void SpecNode::absorb(const MetricChange& u, bool threshold, double local_accept_rate) {
    assert(!complete() || u.last_position <= position_);
    step_time_ += u.step_time;
    step_count_ += u.step_count;
    if (u.last_position <= position_) return;
    if (u.first_position == 0) {
        metric_sum_ = u.metric_sum;
        metric_sum_sq_ = u.metric_sum_sq;
    } else {
        metric_sum_ += u.metric_sum;
        metric_sum_sq_ += u.metric_sum_sq;
    }
    position_ = u.last_position;
    trace_sum_.push_back(u.metric_sum);
    trace_sum_sq_.push_back(u.metric_sum_sq);
    trace_position_.push_back(u.last_position);
    update_branch_probability(threshold, local_accept_rate);
    for (SpecNode* n = child_[1]; n; n = n->child_[0]) n->update_branch_probability(threshold, local_accept_rate);
}

// ------------------------------------------------------------------------------------------- SpecTree
// master.cc:213-232.  The first level's log(u) is drawn at construction; the 100-step acceptance
// window starts half full.
SpecTree::SpecTree(int world_size, std::size_t max_depth, SamplerType* hook, std::size_t policy, double correlation,
                   std::size_t dimension)
    : max_depth_(max_depth), hook_(hook), policy_(policy), correlation_(correlation), dimension_(dimension),
      wait_time_(world_size > 0 ? world_size : 0, 0.0) {
    level_log_u_.push_back(std::log(uniform01(u_generator_)));
    if (!hook_) {
        hook_ = new EmptySamplerType;
        own_hook_ = true;
    }
    while (accept_window_.size() < 100) {
        accept_window_.push_back(true);
        accept_window_.push_back(false);
    }
}

SpecTree::~SpecTree() {
    for (SpecNode* n : by_id_) delete n;
    if (own_hook_) delete hook_;
}

SpecNode* SpecTree::node(std::size_t id) {  // master.hh:411-418: id 1 always names the current root
    if (id - first_id_ < by_id_.size()) return by_id_[id - first_id_];
    return id == 1 ? root_ : nullptr;
}

void SpecTree::set_data_size(std::size_t ndata) {  // master.cc:241-253
    assert(ndata != 0 && (ndata_ == 0 || ndata_ == ndata));
    if (ndata_ != 0) return;
    ndata_ = ndata;
    root_ = new SpecNode(ndata_, policy_, correlation_, dimension_);
    by_id_.push_back(root_);
    ++allocated_nodes_;
    root_->executor_ = 0;
}

void SpecTree::set_require_comparison_theta(bool x) {
    assert(!root_ && first_id_ == 1 && by_id_.empty());
    require_comparison_theta_ = x;
}

// master.cc:185-211
void SpecTree::create_child(SpecNode* parent, bool accept_side) {
    if (parent->id() == JobInfo::root_id && !accept_side) return;  // the true root has no reject child
    if (parent->child_[accept_side]) return;
    if (!accept_side && !parent->comparison_parent_) return;       // (a re-parented root: nothing to compare with)
    // one log(u) per level, drawn the first time the level is touched
    if (level_log_u_.size() + first_depth_ <= parent->depth_ + 1) {
        assert(level_log_u_.size() + first_depth_ == parent->depth_ + 1);
        level_log_u_.push_back(std::log(uniform01(u_generator_)));
    }
    const double log_u = level_log_u_[parent->depth_ - first_depth_];
    SpecNode* n = new SpecNode(next_id(), parent, accept_side, log_u, ndata_, local_accept_rate(), policy_,
                               correlation_);
    by_id_.push_back(n);
    ++allocated_nodes_;
    n->executor_ = 0;
    while (next_id() == 0 || next_id() == 1) by_id_.push_back(nullptr);  // ids 0 and 1 are never reused
}

void SpecTree::set_proposal(const ProposalInfo& p, int) {  // master.cc:260-272
    SpecNode* n = node(p.id);
    if (!n) return;
    if (n->has_theta_ && n->theta_ != p.proposal)
        std::cerr << "node " << n->id_ << ": proposal conflict, " << hook_->unparse(n->theta_) << " vs. "
                  << hook_->unparse(p.proposal) << "\n";
    assert(!n->has_theta_ || n->theta_ == p.proposal);
    n->has_theta_ = true;
    n->theta_ = p.proposal;
    n->log_prior_ = p.log_prior;
    n->random_position_ = p.random_position;
    n->next_random_position_ = p.next_random_position;
    n->proposal_time_ = p.proposal_time;
}

void SpecTree::delete_subtree(SpecNode* n) {  // master.cc:282-292
    for (int side = 0; side != 2; ++side)
        if (n->child_[side]) delete_subtree(n->child_[side]);
    by_id_[n->id_ - first_id_] = nullptr;
    wasted_proposal_time_ += n->proposal_time_;
    wasted_step_time_ += n->step_time_;
    delete n;
    --allocated_nodes_;
    ++recycled_nodes_;
}

void SpecTree::delete_loser_child(SpecNode* n) {  // master.cc:294-302
    assert(n && n->complete() && n->comparison_parent_ && n->comparison_parent_->complete() &&
           n->has_final_branch_probability());
    const bool loser = n->branch_probability_ < 0.5;  // rejected -> its accept subtree is dead, and vice versa
    if (n->child_[loser]) delete_subtree(n->child_[loser]);
    n->child_[loser] = nullptr;
}

// One level of the chain is decided: emit it, advance the root (master.cc:304-399).
void SpecTree::trim(const run_behavior& rb, double master_wait_time, double master_work_time,
                    std::size_t iprobe_status, std::size_t iprobe_no_status) {
    assert(root_->depth_ == first_depth_ && root_->complete());
    assert(!root_->child_[0] && root_->child_[1]);
    SpecNode* child = root_->child_[1];
    assert(child->complete() && child->depth_ == child->true_depth_ && child->has_final_branch_probability());
    const bool accept = child->branch_probability_ > 0.5;
    assert(child->child_[accept] && !child->child_[!accept]);
    if (const SpecNode* cp = child->comparison_parent_) {  // how close this level's test (master.cc:89) was to flipping
        const double lhs = (child->log_prior_ - cp->log_prior_) + child->metric_sum() - cp->metric_sum();
        const double margin = std::fabs(lhs - child->log_u_);
        const double mag = std::max(std::fabs(child->metric_sum()), std::fabs(cp->metric_sum()));
        if (std::isfinite(margin)) {
            min_margin_ = std::min(min_margin_, margin);
            if (mag > 0.0) min_margin_rel_ = std::min(min_margin_rel_, margin / mag);
        }
    }

    if (!rb.quiet) {  // the -p table (master.cc:324-345)
        if (child->depth_ == 1)
            std::cerr << "#depth last_executor allocated_nodes useful_proposal_time useful_step_time "
                         "wasted_proposal_time wasted_step_time master_wait_time master_work_time worker_wait_time "
                         "iprobe_status iprobe_no_status flip_count threshold_position local_accept_rate scale\n";
        std::stringstream sbuf;
        sbuf << child->depth_ << " " << child->last_executor_ << " " << allocated_nodes() << " "
             << useful_proposal_time_ << " " << useful_step_time_ << " " << wasted_proposal_time_ << " "
             << wasted_step_time_ << " " << master_wait_time << " " << master_work_time << " "
             << std::accumulate(wait_time_.begin(), wait_time_.end(), 0.0) << " " << iprobe_status << " "
             << iprobe_no_status << " " << child->flip_count_ << " " << child->threshold_position_ << " "
             << local_accept_rate() << " " << child->scale_ << "\n";
        std::cerr << sbuf.str();
    }

    if (root_->depth_ == 0) hook_->notify(*this, root_->depth_, root_->metric_sum(), root_->theta(), true);
    if (child->depth_ <= max_depth_) {
        const SpecNode* state = accept ? child : child->comparison_parent_;
        hook_->notify(*this, child->depth_, state->metric_sum(), state->theta(), accept);
    }

    by_id_[root_->id_ - first_id_] = nullptr;
    if (accept) {  // the child becomes the state
        useful_proposal_time_ += root_->proposal_time_;
        useful_step_time_ += root_->step_time_;
        delete root_;
        root_ = child;
        child->parent_ = child->comparison_parent_ = nullptr;
    } else {  // the state survives one more step: adopt the reject grandchild, take over the child's id
        SpecNode* grandchild = child->child_[0];
        root_->id_ = child->id_;
        root_->child_[1] = grandchild;
        ++root_->depth_;
        root_->random_position_ = child->random_position_;
        root_->next_random_position_ = child->next_random_position_;
        by_id_[root_->id_ - first_id_] = root_;
        useful_proposal_time_ += child->proposal_time_;
        useful_step_time_ += child->step_time_;
        delete child;
        grandchild->parent_ = root_;
    }
    --allocated_nodes_;
    ++recycled_nodes_;
    ++first_depth_;
    level_log_u_.pop_front();
    while (by_id_.front() == nullptr) {
        ++first_id_;
        by_id_.pop_front();
    }

    accept_count_ += (int) accept;  // 100-step acceptance window (master.cc:392-398)
    if (accept_window_.size() == 100) {
        accept_count_ -= (int) accept_window_.front();
        accept_window_.pop_front();
    }
    accept_window_.push_back(accept);
}

// master.cc:401-426
void SpecTree::set_metric(const MetricChange& u, const run_behavior& rb, double master_wait_time,
                          double master_work_time, int rank, std::size_t iprobe_status,
                          std::size_t iprobe_no_status) {
    if (rank >= 0 && (std::size_t) rank < wait_time_.size()) wait_time_[rank] = u.worker_wait_time;
    SpecNode* n = node(u.id);
    if (!n) {
        wasted_step_time_ += u.step_time;
        return;
    }
    n->absorb(u, threshold_, local_accept_rate());
    if (!n->complete()) return;
    n->last_executor_ = n->executor_;
    n->executor_ = -1;
    if (n->comparison_parent_ && n->comparison_parent_->complete()) delete_loser_child(n);
    for (SpecNode* c = n->child_[1]; c; c = c->child_[0])
        if (c->complete()) delete_loser_child(c);
    while (root_->complete() && root_->child_[1]->complete())  // maybe a level (or several) just completed
        trim(rb, master_wait_time, master_work_time, iprobe_status, iprobe_no_status);
}

void SpecTree::abandon(std::size_t id, int executor) {  // master.cc:428-433
    SpecNode* n = node(id);
    if (n && !n->complete() && n->executor_ == executor) n->executor_ = 0;
}

// Stochastic root-to-leaf walk (master.cc:435-454): follow decided branches, flip a coin weighted by
// the branch probability at undecided ones, stop at the first pending node (or fall off the tree).
std::pair<std::size_t, double> SpecTree::high_utility_pending(std::size_t stopper) {
    const SpecNode* n = root_;
    double u = 1.0;
    while (n && !n->pending()) {
        if (n->id() == stopper) return {0, 0.0};
        if (n->complete())
            n = n->child_[n->branch_probability_ > 0.5];
        else {
            const bool side = n->branch_probability_ >= uniform01(utility_generator_);
            u *= side ? n->branch_probability_ : 1 - n->branch_probability_;
            n = n->child_[side];
        }
    }
    if (n) return {n->id(), u};
    return {0, 0.0};
}

// Hand node `id` to `executor` (master.cc:456-509).  If the node's theta is not known yet the job
// carries the nearest known ancestor state plus the accept/reject path from it, to be replayed by
// setup_job (worker.hh:28-49).
JobInfo SpecTree::execute(std::size_t id, int executor) {
    JobInfo info;
    SpecNode* n = node(id);
    if (!n) {
        info.id = 0;
        return info;
    }
    assert(n->pending());
    n->executor_ = executor;
    ++n->generation_;
    if (!require_comparison_theta_ || n->has_theta()) create_children(n);

    info.id = n->id_;
    info.generation = n->generation_;
    info.utility = n->utility();
    if (n->has_theta()) {
        info.state = JobInfo::s_has_proposal;
        info.theta = n->theta();
        info.random_position = n->random_position_;
    } else {
        SpecNode* cp = n->comparison_parent_;
        while (cp && !cp->has_theta() && cp->comparison_parent_) cp = cp->comparison_parent_;
        SpecNode* walk = n;
        SpecNode* p;
        while ((p = walk->parent_) && p != cp &&
               (!p->has_theta() || p->comparison_parent_ != cp || !cp->has_theta())) {
            info.path += walk->is_accept() ? "a" : "r";
            walk = p;
        }
        if (cp && cp->has_theta()) {
            info.state = JobInfo::s_need_proposal;
            info.theta = cp->theta();
        } else {
            info.state = JobInfo::s_need_initial;
            if (n->parent_) info.path += "a";  // "accept" the initial proposal
        }
        info.random_position = p ? p->next_random_position_ : 0;
    }
    info.log_u = level_log_u_[n->depth_ - first_depth_];
    info.first_position = n->position_;
    info.last_position = n->ndata_;
    info.metric_sum = n->metric_sum_;
    info.metric_sum_sq = n->metric_sum_sq_;
    info.scale = n->scale_;
    return info;
}

static void print_node(std::ostream& s, const SpecNode* n, const SamplerType* sampler, int indent) {
    s << std::setw(indent) << "" << "#" << n->id() << " D" << n->true_depth();
    if (n->complete())
        s << (n->branch_probability() > 0.5 ? " (accept): " : " (reject): ");
    else
        s << " (" << n->branch_probability() << "): ";
    s << (n->has_theta() ? sampler->unparse(n->theta()) : std::string("<pending>"));
    s << " (u " << n->utility();
    if (n->position()) {
        const std::streamsize prec = s.precision(3);
        s << ", " << (100 * double(n->position()) / n->data_size()) << "%";
        s.precision(prec);
    }
    if (n->pending()) s << ", pending";
    s << ")\n";
    for (int side = 0; side != 2; ++side)
        if (n->child(side)) print_node(s, n->child(side), sampler, indent + 2);
}

void SpecTree::print(std::ostream& s) const {
    if (root_)
        print_node(s, root_, hook_, 0);
    else
        s << "EMPTY\n";
}

std::ostream& operator<<(std::ostream& s, const SpecTree& tree) {
    tree.print(s);
    return s;
}

// ------------------------------------------------------------------------------------------- SamplerType
namespace {
struct Named {
    std::string name;
    SamplerType* sampler;
};
std::vector<Named>& registry() {
    static std::vector<Named>* r = new std::vector<Named>;  // never destroyed: plugins sign up from static ctors
    return *r;
}
}  // namespace

int SamplerType::signup(const char* name, SamplerType* s) {
    if (find(name)) return -1;
    registry().push_back({name, s});
    return 0;
}

SamplerType* SamplerType::find(const char* name) {
    for (const Named& n : registry())
        if (n.name == name) return n.sampler;
    return nullptr;
}

void SamplerType::print_list(std::ostream& o) {
    for (const Named& n : registry()) o << n.name << "\n";
}

// stype.cc:67-84: depth 0 with accept == false prints the header and starts the clock
void SamplerType::notify(const SpecTree&, std::size_t depth, const double metric_sum, const std::string& theta,
                         bool accept) {
    std::ostream& o = *out;
    if (depth == 0 && !accept) {
        t0_ = timestamp();
        o << "time\tdepth\taccept\tmetric_sum\ttheta\n";
        return;
    }
    char tbuf[100];
    snprintf(tbuf, sizeof tbuf, "%.06f", timestamp() - t0_);
    o.precision(5);
    o << tbuf << "\t" << depth << "\t" << accept << "\t" << std::fixed;
    o << metric_sum << "\t" << unparse(theta) << "\n";
}

} }  // namespace pf::host
