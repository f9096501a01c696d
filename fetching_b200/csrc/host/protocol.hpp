/* Derived from Fetching (elaine84/fetching): protocol.hh (message tags and the JobInfo / ProposalInfo / MetricChange / AbandonInfo records).
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
// host/protocol.hpp -- the data model that crosses the master <-> evaluator seam.
//
// Same records, same field names and same meaning as protocol.hh:26-191 of the reference, so a
// maintainer can map one onto the other field by field; there is no serialisation here because
// nothing crosses a process boundary any more (the frontier is handed over in-process).
#pragma once
#include <cstddef>
#include <cstring>
#include <string>

namespace pf { namespace host {

enum UpdateStyle { FAST_SLOW = 0, INCREMENTAL = 1, LOGARITHMIC = 2 };  // protocol.hh:26-30

enum MessageTag {  // protocol.hh:32-41
    MSG_PAUSE = 1,
    MSG_W2M_REGISTER,
    MSG_W2M_WANTWORK,
    MSG_M2W_HAVEWORK,
    MSG_W2M_SETPROPOSAL,
    MSG_W2M_UPDATE,
    MSG_M2W_ABANDON,
    MSG_W2M_QUIT
};
constexpr int ANY_TAG = -1;

struct MetricChange {  // protocol.hh:43-53
    std::size_t id = 0;
    std::size_t first_position = 0;
    std::size_t last_position = 0;
    double metric_sum = 0.0;
    double metric_sum_sq = 0.0;
    double step_time = 0.0;
    std::size_t step_count = 0;
    bool abandon = false;
    double worker_wait_time = 0.0;
};

struct JobInfo {  // protocol.hh:90-116
    enum { root_id = 1 };
    enum { s_need_initial = 0, s_need_proposal = 1, s_has_proposal = 2 };
    std::size_t id = 0;
    std::size_t generation = 0;
    double log_u = 0.0;
    std::size_t random_position = 0;
    std::size_t first_position = 0;
    std::size_t last_position = 0;
    double utility = 0.0;
    double metric_sum = 0.0;
    double metric_sum_sq = 0.0;
    double scale = 0.0;
    int state = s_need_initial;
    std::string theta;
    std::string path;
    bool need_initial() const { return state == s_need_initial; }
    bool need_proposal() const { return state == s_need_proposal; }
    bool has_proposal() const { return state == s_has_proposal; }
};

struct ProposalInfo {  // protocol.hh:140-159
    std::size_t id = 0;
    std::string proposal;
    double log_prior = 0.0;
    std::size_t random_position = 0;
    std::size_t next_random_position = 0;
    double proposal_time = 0.0;
    ProposalInfo() {}
    ProposalInfo(std::size_t node_id, const std::string& p, double lp, std::size_t rp, std::size_t nrp,
                 double t = 0.0)
        : id(node_id), proposal(p), log_prior(lp), random_position(rp), next_random_position(nrp),
          proposal_time(t) {}
};

struct AbandonInfo {  // protocol.hh:174-183
    std::size_t id = 0;
    std::size_t generation = 0;
    AbandonInfo() {}
    AbandonInfo(std::size_t node_id, std::size_t gen) : id(node_id), generation(gen) {}
};

// theta codec: the master treats theta as an opaque byte string (master.hh:49); for every native
// sampler it is the raw bytes of its doubles, which is also what boost's binary archive with
// no_header produces for `double` and NormalTheta<D> (protocol.hh:200-215, normaltheta.hh:41-46).
inline std::string save_doubles(const double* v, std::size_t n) {
    return std::string(reinterpret_cast<const char*>(v), n * sizeof(double));
}
inline std::size_t load_doubles(const std::string& s, double* out, std::size_t max_n) {
    std::size_t n = s.size() / sizeof(double);
    if (n > max_n) n = max_n;
    std::memcpy(out, s.data(), n * sizeof(double));
    return n;
}

} }  // namespace pf::host
