/* shim: master.cc:135 calls unqualified erf(), which is libm's. Nothing to provide. */
#ifndef PF_SHIM_BOOST_MATH_ERF_HPP
#define PF_SHIM_BOOST_MATH_ERF_HPP
#include <math.h>
#endif
