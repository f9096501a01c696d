/* shim: inverse of binary_oarchive.hpp (protocol.hh:208-215). */
#ifndef PF_SHIM_BOOST_ARCHIVE_BINARY_IARCHIVE_HPP
#define PF_SHIM_BOOST_ARCHIVE_BINARY_IARCHIVE_HPP
#include <streambuf>
#include <type_traits>
#include "binary_oarchive.hpp"
namespace boost { namespace archive {
class binary_iarchive {
  public:
    binary_iarchive(std::streambuf& sb, unsigned = 0) : sb_(sb) {}
    template <typename T>
    typename std::enable_if<std::is_arithmetic<T>::value, binary_iarchive&>::type
    operator&(T& x) {
        sb_.sgetn(reinterpret_cast<char*>(&x), sizeof(T));
        return *this;
    }
    template <typename T>
    typename std::enable_if<!std::is_arithmetic<T>::value, binary_iarchive&>::type
    operator&(T& x) {
        boost::serialization::access::serialize(*this, x, 0u);
        return *this;
    }
  private:
    std::streambuf& sb_;
};
} }
#endif
