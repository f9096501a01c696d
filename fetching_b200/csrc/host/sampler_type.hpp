/* Derived from Fetching (elaine84/fetching): stype.hh / stype.cc (SamplerType plugin interface, registry, default chain printer).
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
// host/sampler_type.hpp -- the sampler plugin interface and registry, as in stype.hh:23-69.
//
// Same virtuals, same registry contract (signup returns -1 for a taken name, find returns 0 for an
// unknown one) and the same default chain printer (stype.cc:67-84): a TSV header
// `time depth accept metric_sum theta`, then one row per level with 5 fixed decimals.
// `work()` takes the chain configuration and a CUDA device instead of an MPI communicator: there are no
// worker processes any more -- a plugin's work() builds its data set (what every reference worker does for
// itself, samplers/normal.cc:38-48), puts it on the GPU behind the C ABI (pf_engine_create) and runs the chain
// through it (pf_chain_run).  Plugins register from static constructors (samplers.cpp; samplers/normal.cc:59-62)
// and take their arguments as `key=value` strings (pymixture.py:193-204, pyblasso.py:254-267).
#pragma once
#include <iostream>
#include <string>
#include <vector>

#include "../../../include/pfetch.h"
#include "protocol.hpp"

namespace pf { namespace host {

class SpecTree;

class SamplerType {
  public:
    SamplerType() {}
    virtual ~SamplerType() {}

    static int signup(const char* name, SamplerType* s);   // stype.cc:29-34
    static SamplerType* find(const char* name);            // stype.cc:36-43
    static void print_list(std::ostream& out = std::cout); // stype.cc:45-49

    virtual int set_arguments(std::vector<const char*>& argv, bool complain) {
        (void) argv;
        (void) complain;
        return 0;
    }
    // 0 ok, -1 "this plugin cannot work" (stype.cc:55-61), > 0: the chain failed (text in pf_last_error()).
    // `cfg` carries tree.cc's flags; the plugin fills in what only it knows (seed, initial theta).
    virtual int work(int device, const pf_chain_config& cfg, pf_chain_stats* stats) {
        (void) device;
        (void) cfg;
        (void) stats;
        return -1;
    }
    virtual std::string unparse(const std::string& theta) const {
        (void) theta;
        return "<proposal>";
    }
    virtual void notify(const SpecTree& tree, std::size_t depth, const double metric_sum, const std::string& theta,
                        bool accept);

    double wait_time = 0.0;
    std::ostream* out = &std::cout;  // where notify() prints the chain

  protected:
    double t0_ = 0.0;
};

class EmptySamplerType : public SamplerType {
  public:
    void notify(const SpecTree&, std::size_t, const double, const std::string&, bool) override {}
};

} }  // namespace pf::host
