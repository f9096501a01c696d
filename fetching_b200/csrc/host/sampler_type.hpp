// host/sampler_type.hpp -- the sampler plugin interface and registry, as in stype.hh:23-69.
//
// Same virtuals, same registry contract (signup returns -1 for a taken name, find returns 0 for an
// unknown one) and the same default chain printer (stype.cc:67-84): a TSV header
// `time depth accept metric_sum theta`, then one row per level with 5 fixed decimals.
// `work()` takes the engine-backed frontier runner instead of an MPI communicator: there are no
// worker processes any more -- the evaluation side is a C-ABI call (include/pfetch.h).
#pragma once
#include <iostream>
#include <string>
#include <vector>

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
