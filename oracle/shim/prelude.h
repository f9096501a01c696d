/* oracle/shim/prelude.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Force-included (-include) ahead of every reference translation unit that the
 * oracle build compiles *in place* from /root/reference.  It supplies the libc
 * headers the reference used to get transitively from Boost, and one overload
 * that modern libstdc++ made ambiguous.
 *
 * master.hh:330-332 calls abs(size_t - size_t); under gcc 4.7 that resolved to
 * ::abs(int).  Reproduce exactly that narrowing. */
#ifndef PF_ORACLE_SHIM_PRELUDE_H
#define PF_ORACLE_SHIM_PRELUDE_H
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef __cplusplus
#include <cmath>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <sstream>
#include <string>
inline int abs(unsigned long x) { return ::abs(static_cast<int>(x)); }
#endif
#endif
