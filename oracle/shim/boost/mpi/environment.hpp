#ifndef PF_SHIM_BOOST_MPI_ENVIRONMENT_HPP
#define PF_SHIM_BOOST_MPI_ENVIRONMENT_HPP
namespace boost { namespace mpi {
class environment { public: environment(int&, char**&) {} };
} }
#endif
