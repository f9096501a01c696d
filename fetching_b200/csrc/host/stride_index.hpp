/* Derived from Fetching (elaine84/fetching): strideindex.hh (StrideIndex).
 *
 * Fetching
 * Elaine Angelino, Eddie Kohler
 * Copyright (c) 2012-2014 President and Fellows of Harvard College
 *
 * Permission is hereby granted, free of charge, to any person obtaining a
 * copy of this software and associated documentation files (the "Software"),
 * to deal in the Software without restriction, subject to the conditions
 * listed in the Fetching LICENSE file. These conditions include: you must
 * preserve this copyright notice, and you cannot mention the copyright
 * holders in advertising related to the Software without their permission.
 * The Software is provided WITHOUT ANY WARRANTY, EXPRESS OR IMPLIED. This
 * notice is a summary of the Fetching LICENSE file; the license in that file
 * is legally binding.  (The licence text is reproduced in LICENSE.fetching
 * next to this file.)
 */
// host/stride_index.hpp -- per-node pseudo-random visiting order of the items
// ((start + i*stride) mod n with gcd(n, stride) = 1); strideindex.hh:19-57 of the reference.
// Only PARTIAL sums depend on it; it is kept so that prepare() consumes the same draws.
#pragma once
#include <cstddef>
#include <numeric>

namespace pf { namespace host {

class StrideIndex {
  public:
    StrideIndex() {}
    StrideIndex(std::size_t n, std::size_t stride, std::size_t start = 0) : n_(n), stride_(stride), start_(start) {}
    std::size_t operator()(std::size_t i) const { return (start_ + i * stride_) % n_; }
    std::size_t stride() const { return stride_; }
    std::size_t start() const { return start_; }

    // bump `stride` until it is coprime to n (strideindex.hh:50-57)
    static std::size_t valid_stride(std::size_t n, std::size_t stride) {
        if (stride <= 1) return 1;
        for (std::size_t g; (g = std::gcd(n, stride)) > 1;) stride = (stride + g - 1) % n;
        return stride <= 1 ? 1 : stride;
    }

  private:
    std::size_t n_ = 0, stride_ = 1, start_ = 0;
};

} }  // namespace pf::host
