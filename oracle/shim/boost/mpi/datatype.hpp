/* shim: only the trait the reference specialises (protocol.hh:194-197, normaltheta.hh:148-150). */
#ifndef PF_SHIM_BOOST_MPI_DATATYPE_HPP
#define PF_SHIM_BOOST_MPI_DATATYPE_HPP
namespace boost { namespace mpl {
struct true_ { static const bool value = true; };
struct false_ { static const bool value = false; };
} namespace mpi {
namespace mpl = boost::mpl;
template <typename T> struct is_mpi_datatype : public mpl::false_ {};
} }
#endif
