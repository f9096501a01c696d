/* shim: boost::optional on top of std::optional (test infrastructure only). */
#ifndef PF_SHIM_BOOST_OPTIONAL_HPP
#define PF_SHIM_BOOST_OPTIONAL_HPP
#include <optional>
namespace boost {
template <typename T> class optional : public std::optional<T> {
  public:
    optional() {}
    optional(const T& v) : std::optional<T>(v) {}
    optional(const std::optional<T>& o) : std::optional<T>(o) {}  /* lets a std::optional-returning Communicator drive run_master */
    T& get() { return **this; }
    const T& get() const { return **this; }
};
}
#endif
