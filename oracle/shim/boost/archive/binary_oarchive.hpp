/* shim: boost::archive::binary_oarchive with no_header -- raw native bytes of each
 * arithmetic member in declaration order, which is what the real archive emits for
 * the theta types the C++ samplers use (double, NormalTheta<D>: protocol.hh:200-206). */
#ifndef PF_SHIM_BOOST_ARCHIVE_BINARY_OARCHIVE_HPP
#define PF_SHIM_BOOST_ARCHIVE_BINARY_OARCHIVE_HPP
#include <streambuf>
#include <type_traits>
#include "../serialization/access.hpp"
namespace boost { namespace archive {
enum archive_flags { no_header = 1 };
class binary_oarchive {
  public:
    binary_oarchive(std::streambuf& sb, unsigned = 0) : sb_(sb) {}
    template <typename T>
    typename std::enable_if<std::is_arithmetic<T>::value, binary_oarchive&>::type
    operator&(const T& x) {
        sb_.sputn(reinterpret_cast<const char*>(&x), sizeof(T));
        return *this;
    }
    template <typename T>
    typename std::enable_if<!std::is_arithmetic<T>::value, binary_oarchive&>::type
    operator&(const T& x) {
        boost::serialization::access::serialize(*this, const_cast<T&>(x), 0u);
        return *this;
    }
  private:
    std::streambuf& sb_;
};
} }
#endif
