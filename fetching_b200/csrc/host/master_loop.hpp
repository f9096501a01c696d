/* Derived from Fetching (elaine84/fetching): masterloop.cc (JobTree::run_master), statement for statement, with the locals renamed.
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
// host/master_loop.hpp -- the master's event loop over the Communicator concept.
//
// Same protocol as JobTree::run_master (masterloop.cc:21-134): poll the workers round-robin,
// dispatch on the message tag, then hand every waiting worker the pending node of highest
// utility.  That hand-out burst IS the speculative frontier: with FrontierCommunicator
// (frontier_comm.hpp) every JobInfo of a burst is scored by ONE call into the GPU engine.
//
// Communicator concept (masterloop.cc:21-22; second instance: MasterTester, testmaster.cc:57-102):
//   int size();  optional<Status> iprobe(int rank, int tag)   with Status::tag(), Status::source();
//   recv(rank, tag), recv(rank, tag, size_t&), recv(rank, tag, ProposalInfo&), recv(rank, tag, MetricChange&);
//   Request isend(rank, tag, const JobInfo&), isend(rank, tag, const AbandonInfo&);  Request::test().
#pragma once
#include <sys/time.h>

#include <cassert>
#include <cstdio>
#include <deque>
#include <iostream>
#include <sstream>
#include <vector>

namespace pf { namespace host {

inline double timestamp() {  // time.hh:21-25
    struct timeval now;
    gettimeofday(&now, nullptr);
    return now.tv_sec + now.tv_usec * 0.000001;
}

template <typename Communicator>
void SpecTree::run_master(Communicator& world, const run_behavior& rb) {
    using Request = decltype(world.isend(0, 0, std::declval<const JobInfo&>()));
    int quits = 0;
    std::size_t probes_hit = 0, probes_miss = 0;
    bool want_to_quit = false;
    double wait_time = 0.0, work_time = 0.0;
    std::deque<int> idle;                       // ranks that asked for work and have none yet
    std::deque<Request> in_flight;
    std::vector<std::size_t> assigned(world.size(), 0);

    int rank = 1;
    double t0 = timestamp();
    for (;;) {
        auto status = world.iprobe(rank, ANY_TAG);
        if (status) {
            const double t1 = timestamp();
            wait_time += t1 - t0;
            ++probes_hit;
            const int tag = (*status).tag();

            if (tag == MSG_W2M_WANTWORK) {
                if (assigned[rank]) abandon(assigned[rank], rank);
                world.recv(rank, tag);
                if (want_to_quit) {
                    JobInfo stop;
                    stop.id = 0;
                    in_flight.push_back(world.isend(rank, MSG_M2W_HAVEWORK, stop));
                } else {
                    if (rb.print_sending_work) fprintf(stderr, "%f sending work to: %d\n", timestamp(), rank);
                    idle.push_back(rank);
                }
            } else if (tag == MSG_W2M_SETPROPOSAL) {
                ProposalInfo proposal;
                world.recv(rank, tag, proposal);
                set_proposal(proposal, rank);
                if (SpecNode* n = node(proposal.id)) create_children(n);
            } else if (tag == MSG_W2M_UPDATE) {
                MetricChange update;
                world.recv(rank, tag, update);
                set_metric(update, rb, wait_time, work_time, rank, probes_hit, probes_miss);
                const SpecNode* n = node(update.id);
                // take the worker off a node whose further evaluation is pointless ...
                if (update.abandon || !n || n->is_dead() || n->complete())
                    abandon(update.id, rank);
                // ... or much less useful than something that is still pending
                else if (reassign_ratio_ > 0 &&
                         n->utility() * reassign_ratio_ <= high_utility_pending(update.id).second) {
                    AbandonInfo ab(update.id, n ? n->generation() : 0);
                    in_flight.push_back(world.isend(rank, MSG_M2W_ABANDON, ab));
                }
            } else if (tag == MSG_W2M_REGISTER) {
                std::size_t data_size = 0;
                world.recv(rank, tag, data_size);
                set_data_size(data_size);
            } else if (tag == MSG_W2M_QUIT) {
                world.recv(rank, tag);
                if (++quits == world.size() - 1) {
                    if (!rb.quiet) {
                        std::stringstream sbuf;
                        sbuf << "master wait time: " << wait_time << "\n"
                             << "master work time: " << work_time << "\n";
                        std::cerr << sbuf.str();
                    }
                    return;
                }
            } else if (tag == MSG_PAUSE) {
                world.recv(rank, tag);
            } else {
                std::cerr << "Unknown message tag " << tag << "\n";
                assert(0);
            }

            // the scheduling burst: every idle worker gets the best pending node
            std::pair<std::size_t, double> best;
            while (!idle.empty() && (best = high_utility_pending()).first != 0) {
                const int to = idle.front();
                JobInfo info = execute(best.first, to);
                assert(info.id != 0);
                assigned[to] = info.id;
                in_flight.push_back(world.isend(to, MSG_M2W_HAVEWORK, info));
                idle.pop_front();
            }
            while (!in_flight.empty() && in_flight.front().test()) in_flight.pop_front();
            if (max_depth_ && completed_depth() > max_depth_) want_to_quit = true;
            t0 = timestamp();
            work_time += t0 - t1;
        } else
            ++probes_miss;

        if (++rank == world.size()) rank = 1;
    }
}

} }  // namespace pf::host
