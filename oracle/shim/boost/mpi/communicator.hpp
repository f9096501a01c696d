/* shim: a boost::mpi::communicator that exists only so that the declaration
 * JobTree::run_master(boost::mpi::communicator&, ...) (master.hh:121, master.cc:551-556)
 * and worker<>() (worker.hh:53-177) type-check.  There is no MPI in this image; every
 * data-moving member aborts.  The oracle drives the reference tree through the
 * Communicator *concept* (masterloop.cc:21-22) exactly like testmaster.cc does. */
#ifndef PF_SHIM_BOOST_MPI_COMMUNICATOR_HPP
#define PF_SHIM_BOOST_MPI_COMMUNICATOR_HPP
#include <stdlib.h>
#include "../optional.hpp"
#include "datatype.hpp"
#ifndef MPI_ANY_TAG
#define MPI_ANY_TAG (-1)
#endif
#ifndef MPI_ANY_SOURCE
#define MPI_ANY_SOURCE (-2)
#endif
namespace boost { namespace mpi {
class status {
  public:
    int source() const { return 0; }
    int tag() const { return 0; }
};
class request {
  public:
    bool test() { return true; }
};
class communicator {
  public:
    int rank() const { return 0; }
    int size() const { return 1; }
    void barrier() const {}
    boost::optional<status> iprobe(int, int) const { abort(); }
    status recv(int, int) const { abort(); }
    template <typename T> status recv(int, int, T&) const { abort(); }
    void send(int, int) const { abort(); }
    template <typename T> void send(int, int, const T&) const { abort(); }
    template <typename T> request isend(int, int, const T&) const { abort(); }
};
} }
#endif
