// host/setup_job.hpp -- materialise a node's proposal on the host before it is scored.
//
// Semantics of Worker::setup_job (worker.hh:28-49): call the Executor in the fixed order
// load -> log_prior -> random_position -> prepare -> random_position, and, when the job carries
// an accept/reject `path` (its ancestors' thetas were not known when it was handed out), replay
// the proposals along that path -- last path character first -- moving to the proposed state on
// 'a' and staying on 'r'.
#pragma once
#include <cstddef>

namespace pf { namespace host {

template <typename Executor, typename JobInfoT, typename ProposalInfoT>
void setup_job(Executor& executor, JobInfoT& info, ProposalInfoT& proposal) {
    for (std::size_t replayed = 0;; ++replayed) {
        proposal.proposal = executor.load(info);
        proposal.log_prior = executor.log_prior();
        proposal.random_position = executor.random_position();
        executor.prepare();
        proposal.next_random_position = executor.random_position();
        if (info.has_proposal() || replayed == info.path.length()) {
            proposal.id = info.id;
            return;
        }
        info.random_position = proposal.next_random_position;
        info.state = JobInfoT::s_need_proposal;
        if (info.path[info.path.length() - 1 - replayed] == 'a') info.theta = proposal.proposal;
    }
}

} }  // namespace pf::host
