/* shim: boost::math::gcd (strideindex.hh:42,54). */
#ifndef PF_SHIM_BOOST_MATH_COMMON_FACTOR_RT_HPP
#define PF_SHIM_BOOST_MATH_COMMON_FACTOR_RT_HPP
namespace boost { namespace math {
template <typename T> inline T gcd(T a, T b) {
    while (b != 0) { T t = a % b; a = b; b = t; }
    return a;
}
} }
#endif
