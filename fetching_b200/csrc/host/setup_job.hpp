/* Derived from Fetching (elaine84/fetching): worker.hh:28-49 (Worker::setup_job).
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
