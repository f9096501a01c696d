// pf_common.cuh -- device-side building blocks shared by the frontier kernels (sm_100a).
//
// * mbarrier + cp.async.bulk (TMA) wrappers: data tiles are staged into shared memory by the
//   copy engine (SASS: UBLKCP for the 1-D form, UTMALDG for the tensor-map form) and signalled
//   through transaction barriers, so no thread spends issue slots or registers on loads.
// * fixed-order reductions: no floating-point atomics anywhere, so results are run-to-run
//   bit-stable for a given launch geometry.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pf_engine.h"

namespace pf {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make mbarrier.init visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 1-D bulk copy global -> shared, completion counted in bytes on `bar`.
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// 2-D tiled TMA load: box at (c0 = innermost coordinate, c1 = row) of `map` -> shared.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// L2 eviction policies for streamed data: a matrix that is read once per call and is larger than L2 (lasso: 400 / 800 MB
// through 126 MB) otherwise pushes out everything the call still needs from L2 -- the coefficient panel every CTA
// re-reads, and the per-CTA records the finalizing CTA folds (which then came back from HBM: 1.3 us per round trip).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::
            "r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// named barrier over `count` threads (the consumer warps only; the producer warp never joins)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// non-blocking arrival on a named barrier (the matching waiters use named_bar_sync)
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ double shfl_down_f64(double v, int delta) {
    const int hi = __shfl_down_sync(0xffffffffu, __double2hiint(v), delta);
    const int lo = __shfl_down_sync(0xffffffffu, __double2loint(v), delta);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_up_f64(double v, int delta) {
    const int hi = __shfl_up_sync(0xffffffffu, __double2hiint(v), delta);
    const int lo = __shfl_up_sync(0xffffffffu, __double2loint(v), delta);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor_f64(double v, int mask) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, mask);
    return __hiloint2double(hi, lo);
}

// sum over the lanes that differ from this one in bits [lo_bit .. 16]: butterfly, fixed order
__device__ __forceinline__ double warp_sum_from(double v, int first_mask) {
    for (int m = 16; m >= first_mask; m >>= 1) v += shfl_xor_f64(v, m);
    return v;
}

__device__ __forceinline__ double ld_cg_f64(const double* p) { return __ldcg(p); }

// ---------------------------------------------------------------------------------------------
// Item assembly.  metric_sum_sq is a sum of squares of ITEM totals (worker.hh:151-152), so row
// contributions must be summed per item before squaring.  A CTA scores a contiguous row range
// [r0, r1) in tiles that never straddle an item.  One owner thread per node folds tile totals
// into item totals; an item cut by the CTA boundary is exported as a `head` (its beginning is
// in an earlier CTA) or `tail` (its end is in a later CTA) partial, and finalize_items() stitches
// those in CTA order.  Everything is a fixed-order sum: bit-stable for a given geometry.
// Record layout: [S | Q | head | tail | stitch][kMaxNodesPerPass][kMaxGrid CTAs] (transposed: see ItemOwner::store).
// ---------------------------------------------------------------------------------------------
struct ItemOwner {
    double S, Q, run, head, tail;
    double stitch;  // index of the CTA in which the cut item that ENDS in this CTA began, or -1
    unsigned long long next_close;  // the next item boundary after the rows folded so far (no modulo per tile)
    bool from_start, open;
    __device__ __forceinline__ void init(unsigned long long r0, unsigned long long item_rows) {
        S = Q = run = head = tail = 0.0;
        stitch = -1.0;
        from_start = (item_rows <= 1) || (r0 % item_rows == 0);
        next_close = item_rows <= 1 ? r0 + 1 : (r0 / item_rows + 1) * item_rows;
        open = false;
    }
    // `tot` = this node's total over tile [.., t1); tiles arrive in row order and never straddle an item
    __device__ __forceinline__ void add_tile(double tot, unsigned long long t1, unsigned long long item_rows,
                                             unsigned long long row_end) {
        run += tot;
        const bool at_boundary = t1 >= next_close;
        if (at_boundary) next_close += item_rows;
        const bool closes = at_boundary || (t1 == row_end);
        if (closes) {
            if (from_start) {
                S += run;
                Q = fma(run, run, Q);
            } else
                head = run;
            run = 0.0;
            from_start = true;
        }
        open = !closes;
    }
    // Called once per CTA with its row range [c0, c1): the 64-bit divisions that locate the beginning of a cut
    // item are paid here, by every CTA in parallel, instead of in the finalizing CTA's serial tail.
    __device__ __forceinline__ void finish(unsigned long long c0, unsigned long long c1, unsigned long long row_begin,
                                           unsigned long long row_end, unsigned long long rows_per_cta,
                                           unsigned long long item_rows) {
        if (open) {
            if (from_start)
                tail = run;
            else
                head = run;  // the whole CTA lies inside one item
        }
        if (item_rows > 1 && rows_per_cta && c1 > c0) {
            const unsigned long long q0 = c0 / item_rows;
            if (c0 != q0 * item_rows) {  // an item begun earlier reaches into this CTA
                unsigned long long ie = (q0 + 1) * item_rows;
                if (ie > row_end) ie = row_end;
                if (ie <= c1) stitch = (double) ((q0 * item_rows - row_begin) / rows_per_cta);  // ... and ends here
            }
        }
    }
    // `rec` = part + (index of this CTA): the records are stored TRANSPOSED, [slot][node][CTA], so that the finalizing CTA
    // reads every (slot, node) row with coalesced loads (see finalize_items)
    __device__ __forceinline__ void store(double* rec, unsigned tid) const {
        rec[(size_t) (0 * kMaxNodesPerPass + tid) * kMaxGrid] = S;
        rec[(size_t) (1 * kMaxNodesPerPass + tid) * kMaxGrid] = Q;
        rec[(size_t) (2 * kMaxNodesPerPass + tid) * kMaxGrid] = head;
        rec[(size_t) (3 * kMaxNodesPerPass + tid) * kMaxGrid] = tail;
        rec[(size_t) (4 * kMaxNodesPerPass + tid) * kMaxGrid] = stitch;
    }
};

// Finalize, run by the consumer threads of the LAST CTA to finish.  Records are laid out [slot][node][CTA]
// (ItemOwner::store).  Warp w folds nodes w, w + nwarps, ...; lane l of it takes the CTAs l, l + 32, ... -- plain S / Q
// sums plus, for every CTA in which a cut item ENDS (slot 4 names the CTA it began in), that item's stitched total
//     tail(first CTA of the item) + head(next CTA) + ... + head(this CTA)
// -- and a butterfly over the lanes adds the 32 partials up.  Fixed order => bit-stable.  Every load of a round (10 CTAs
// per lane: 2 x 148 CTAs in one round) is issued before the first use and is coalesced over the warp: the tail of every
// small call is the latency of these rounds (and, with one 8-byte request per thread, was their number: 7 us at B = 1).
// `nthreads` consumer threads take part (tid < nthreads, whole warps); `scratch` holds 2 * kMaxNodesPerPass doubles.
// Returns S/Q to threads tid < B.
// BATCH x 32 = CTAs folded per round (10: the 2 x 148 CTAs of the pointwise kernels in one round); NI = nodes a warp folds
// TOGETHER (their loads in flight at the same time: the fold of a node is two or three dependent L2 / shared-memory
// round trips, and at B = 64 a warp of the matvec kernels has eight nodes to fold -- <true, 5, 4> there).
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <bool MAY_STITCH = true, int BATCH = 10, int NI = 1>
__device__ __forceinline__ void finalize_items(const double* part, unsigned grid, unsigned long long row_begin,
                                               unsigned long long row_end, unsigned long long rows_per_cta,
                                               unsigned long long item_rows, unsigned tid, unsigned B,
                                               double* scratch, int bar_id, unsigned nthreads, double& S_out,
                                               double& Q_out, double* big = nullptr, size_t big_doubles = 0,
                                               unsigned long long* tr = nullptr) {
    unsigned gmax = grid;  // rows_per_cta == 0: every launched CTA wrote a plain (S, Q) record
    if (rows_per_cta) {
        const unsigned long long n_cta = (row_end - row_begin + rows_per_cta - 1) / rows_per_cta;
        if (n_cta < gmax) gmax = (unsigned) n_cta;
    }
    const bool stitching = MAY_STITCH && item_rows > 1 && rows_per_cta;
    const unsigned nwarps = nthreads / 32, warp = tid / 32, lane = tid & 31;
    // Items much longer than a CTA's row range (lasso: 18000-row items over ~700-row ranges) make every stitch a walk over
    // the `head` partials of all the CTAs inside the item -- a chain of L2 round trips run by a single lane (25 CTAs per
    // item at C4).  With shared memory to spare (`big`: the idle data ring of the matvec kernels) the heads (and tails) are
    // fetched once, coalesced, by all the threads, and turned into prefix sums (below).
    constexpr unsigned PITCH = 32 * BATCH;  // row pitch of the staged partials (a compile-time divisor)
    const bool pre = stitching && big != nullptr && item_rows > 2 * rows_per_cta && gmax <= PITCH && (size_t) B * PITCH <= big_doubles;
    const bool pre_tail = pre && 2 * (size_t) B * PITCH <= big_doubles;  // ... and the tails, when there is room
    if (pre) {
        // flat index space [heads | tails][node][PITCH], sixteen loads per thread in flight (a load-store loop would be one
        // L2 round trip per iteration: 40 of them per warp at B = 64)
        const unsigned per = B * PITCH, total = pre_tail ? 2 * per : per;
        constexpr int SU = 16;  // loads in flight per thread
        for (unsigned base = tid; base < total; base += nthreads * SU) {
            double h[SU];
#pragma unroll
            for (int u = 0; u < SU; ++u) {
                const unsigned idx = base + (unsigned) u * nthreads;
                const unsigned slot = idx >= per ? 3u : 2u, r = idx >= per ? idx - per : idx;
                const unsigned b = r / PITCH, c = r % PITCH;
                h[u] = (idx < total && c < gmax)
                           ? ld_cg_f64(part + (size_t) b * kMaxGrid + (size_t) slot * kMaxNodesPerPass * kMaxGrid + c)
                           : 0.0;
            }
#pragma unroll
            for (int u = 0; u < SU; ++u) {
                const unsigned idx = base + (unsigned) u * nthreads;
                if (idx < total) big[idx] = h[u];
            }
        }
        named_bar_sync(bar_id, (int) nthreads);
    }
    if (tr != nullptr && tid == 0) tr[16 + 3] = globaltimer_ns();  // diagnostics: heads / tails staged
    constexpr size_t ROW = (size_t) kMaxGrid, SLOT = (size_t) kMaxNodesPerPass * kMaxGrid;
    for (unsigned b0 = warp * NI; b0 < B; b0 += nwarps * NI) {
        const double* pS[NI];
        bool on[NI];  // (uniform per warp) node b0 + i exists
        double S[NI], Q[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            on[i] = b0 + i < B;
            pS[i] = part + (size_t) (on[i] ? b0 + i : b0) * ROW;
            S[i] = Q[i] = 0.0;
        }
        if (pre) {
            // a node's heads -> INCLUSIVE PREFIX SUMS over the CTAs, in place (the warp that folds the node owns its row):
            // a cut item's middle is then  P[c] - P[cs]  instead of a walk over the CTAs in between.  The difference of
            // two prefixes carries an error of ~1e-16 of the node's total.
            const unsigned per = (gmax + 31) / 32, k0 = lane * per;
            double* hb[NI];
            double run[NI], acc[NI];
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                hb[i] = big + (size_t) (on[i] ? b0 + i : b0) * PITCH;  // (a missing node repeats node b0: same values rewritten)
                run[i] = 0.0;
            }
            for (unsigned k = k0; k < k0 + per && k < gmax; ++k)  // the NI nodes side by side: their chains overlap
#pragma unroll
                for (int i = 0; i < NI; ++i) run[i] += hb[i][k];
#pragma unroll
            for (int i = 0; i < NI; ++i) acc[i] = run[i];
            for (int m = 1; m < 32; m <<= 1)
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const double up = shfl_up_f64(acc[i], m);
                    if ((int) lane >= m) acc[i] += up;
                }
#pragma unroll
            for (int i = 0; i < NI; ++i) acc[i] -= run[i];  // exclusive offset of this lane's chunk
            for (unsigned k = k0; k < k0 + per && k < gmax; ++k)
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    acc[i] += hb[i][k];
                    if (on[i]) hb[i][k] = acc[i];
                }
            __syncwarp();
        }
        if (tr != nullptr && tid == 0 && b0 == 0) tr[16 + 6] = globaltimer_ns();  // diagnostics: prefix sums done
        for (unsigned c0 = 0; c0 < gmax; c0 += 32 * BATCH) {
            double s[NI][BATCH], q[NI][BATCH], st[NI][BATCH];
#pragma unroll
            for (int i = 0; i < NI; ++i)
#pragma unroll
                for (int u = 0; u < BATCH; ++u) {
                    const unsigned c = c0 + u * 32 + lane;
                    const bool live = on[i] && c < gmax;
                    s[i][u] = live ? ld_cg_f64(pS[i] + c) : 0.0;
                    q[i][u] = live ? ld_cg_f64(pS[i] + SLOT + c) : 0.0;
                    st[i][u] = (stitching && live) ? ld_cg_f64(pS[i] + 4 * SLOT + c) : -1.0;
                }
#pragma unroll
            for (int i = 0; i < NI; ++i)
#pragma unroll
                for (int u = 0; u < BATCH; ++u) {
                    S[i] += s[i][u];
                    Q[i] += q[i][u];
                }
            if (tr != nullptr && tid == 0 && b0 == 0) tr[16 + 7] = globaltimer_ns();  // records loaded and summed
            if (stitching) {
                // a cut item ends in CTA c when st >= 0: its total is tail(cs) + head(cs + 1) + ... + head(c).  The two
                // loads every stitch needs are issued for the whole round before any use; CTAs in between exist only
                // when an item spans more than two CTAs.
                double tl[NI][BATCH], hd[NI][BATCH];
#pragma unroll
                for (int i = 0; i < NI; ++i)
#pragma unroll
                    for (int u = 0; u < BATCH; ++u) {
                        const bool stitch = st[i][u] >= 0.0;
                        const unsigned c = c0 + u * 32 + lane, cs = stitch ? (unsigned) st[i][u] : 0u;
                        if (pre_tail) {
                            tl[i][u] = stitch ? big[(size_t) (B + b0 + i) * PITCH + cs] : 0.0;
                            hd[i][u] = 0.0;
                        } else {
                            tl[i][u] = stitch ? ld_cg_f64(pS[i] + 3 * SLOT + cs) : 0.0;
                            hd[i][u] = (stitch && !pre) ? ld_cg_f64(pS[i] + 2 * SLOT + c) : 0.0;
                        }
                    }
                if (pre) {  // branch free: every shared-memory read of the round is issued before the first use
                    double pc[NI][BATCH], ps[NI][BATCH];
#pragma unroll
                    for (int i = 0; i < NI; ++i)
#pragma unroll
                        for (int u = 0; u < BATCH; ++u) {
                            const bool stitch = st[i][u] >= 0.0;
                            const unsigned c = c0 + u * 32 + lane;  // (< PITCH: inside the node's staged row)
                            const unsigned cs = stitch ? (unsigned) (int) st[i][u] : c;
                            const double* const hb = big + (size_t) (on[i] ? b0 + i : b0) * PITCH;
                            pc[i][u] = hb[c];
                            ps[i][u] = hb[cs];
                        }
#pragma unroll
                    for (int i = 0; i < NI; ++i)
#pragma unroll
                        for (int u = 0; u < BATCH; ++u) {
                            const double tot = st[i][u] >= 0.0 ? tl[i][u] + (pc[i][u] - ps[i][u]) : 0.0;  // heads of the CTAs cs + 1 .. c
                            S[i] += tot;
                            Q[i] = fma(tot, tot, Q[i]);
                        }
                } else {
#pragma unroll
                    for (int i = 0; i < NI; ++i)
#pragma unroll
                        for (int u = 0; u < BATCH; ++u) {
                            if (st[i][u] >= 0.0) {
                                const unsigned c = c0 + u * 32 + lane, cs = (unsigned) st[i][u];
                                double tot = tl[i][u];
                                for (unsigned k = cs + 1; k < c; ++k) tot += ld_cg_f64(pS[i] + 2 * SLOT + k);
                                tot += hd[i][u];
                                S[i] += tot;
                                Q[i] = fma(tot, tot, Q[i]);
                            }
                        }
                }
            }
        }
        if (tr != nullptr && tid == 0 && b0 == 0) tr[16 + 8] = globaltimer_ns();  // cut items stitched
        for (int m = 16; m >= 1; m >>= 1)
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                S[i] += shfl_xor_f64(S[i], m);
                Q[i] += shfl_xor_f64(Q[i], m);
            }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < NI; ++i)
                if (on[i]) {
                    scratch[b0 + i] = S[i];
                    scratch[kMaxNodesPerPass + b0 + i] = Q[i];
                }
        }
        if (tr != nullptr && tid == 0 && b0 == 0) tr[16 + 4] = globaltimer_ns();  // first round of warp 0
    }
    if (tr != nullptr && tid == 0) tr[16 + 5] = globaltimer_ns();  // warp 0 done
    named_bar_sync(bar_id, (int) nthreads);
    double S = 0.0, Q = 0.0;
    if (tid < B) {
        S = scratch[tid];
        Q = scratch[kMaxNodesPerPass + tid];
    }
    S_out = S;
    Q_out = Q;
}

// ---------------------------------------------------------------------------------------------
// Fused all-reduce over NVLink peer memory (see PeerXchg in pf_engine.h).  Called by the consumer
// threads of the finalizing CTA with this rank's totals in (S, Q) of threads tid < B; returns the
// totals over all ranks, summed in rank order (identical bits on every rank) -- or, in gather mode (node split),
// every node's sums from the one rank that scored it, to threads tid < gather_total.
//   1. every rank stores its [S | Q] record into slot (seq & 1), row `rank`, of EVERY rank's buffer
//      (plain peer stores, spread over the CTA), fences at system scope, then releases a flag = seq
//      next to each record;
//   2. it acquires the G flags of its OWN buffer and folds the G records.
// Slot reuse is safe with two slots: a rank can only start call seq + 2 after it has seen every
// peer's flag for seq + 1, which a peer publishes after it finished reading the records of seq.
// A peer that never shows up (died, or made a different sequence of calls) trips a 20 s timeout:
// *status = 1 in mapped host memory and the caller reports an error instead of hanging the GPU.
// ---------------------------------------------------------------------------------------------

// Result delivery of the host path: value i of n as two self-validating 8-byte words (see PointwiseArgs::done_flag).
__device__ __forceinline__ void store_result_words(double* out, unsigned i, double v, unsigned long long seq) {
    volatile unsigned long long* w = reinterpret_cast<volatile unsigned long long*>(out) + 2 * (size_t) i;
    const unsigned long long tag = (seq & 0xFFFFFFFFull) << 32;
    w[0] = tag | (unsigned) __double2loint(v);
    w[1] = tag | (unsigned) __double2hiint(v);
}

// diagnostics: time mark `i` of CTA 0 (row 0) or of the finalizing CTA (row 1); `tr` is nullptr in normal operation
__device__ __forceinline__ void trace_mark(unsigned long long* tr, int row, int i, bool who) {
    if (tr != nullptr && who) tr[row * 16 + i] = globaltimer_ns();
}
// ... and per CTA: [32 + 2 * cta] = kernel entry, [33 + 2 * cta] = record published
__device__ __forceinline__ void trace_cta(unsigned long long* tr, int which, bool who) {
#ifdef PF_TRACE_SMID  // diagnostics: the low byte of the ENTRY time carries the SM the CTA runs on
    if (tr != nullptr && who) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        const unsigned long long t = globaltimer_ns();
        tr[32 + 2 * blockIdx.x + which] = which == 0 ? ((t & ~0xFFull) | (smid & 0xFFu)) : t;
    }
#else
    if (tr != nullptr && who) tr[32 + 2 * blockIdx.x + which] = globaltimer_ns();
#endif
}

static __device__ __noinline__ double2 peer_allreduce(const PeerXchg x, unsigned tid, unsigned B, unsigned nthreads,
                                                      int bar_id, double* scratch, double S, double Q) {
    const unsigned G = x.nranks;
    const size_t slot_base = (size_t) (x.seq & 1) * kXchgMaxRanks * kXchgRec;
    const size_t my_rec = slot_base + (size_t) x.rank * kXchgRec;
    named_bar_sync(bar_id, (int) nthreads);  // everybody is done with `scratch`
    if (tid < B) {
        scratch[tid] = S;
        scratch[kMaxNodesPerPass + tid] = Q;
    }
    named_bar_sync(bar_id, (int) nthreads);
    for (unsigned i = tid; i < G * 2 * B; i += nthreads) {
        const unsigned r = i / (2 * B), j = i - r * 2 * B;
        const unsigned col = j < B ? j : kMaxNodesPerPass + (j - B);
        x.peers[r][my_rec + col] = scratch[col];
    }
    __threadfence_system();
    named_bar_sync(bar_id, (int) nthreads);
    if (tid < G) {
        unsigned long long* f = reinterpret_cast<unsigned long long*>(x.peers[tid] + my_rec + 2 * kMaxNodesPerPass);
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(x.seq) : "memory");
        // ... and wait for rank `tid`'s record in my own buffer
        const unsigned long long* w = reinterpret_cast<const unsigned long long*>(
            x.peers[x.rank] + slot_base + (size_t) tid * kXchgRec + 2 * kMaxNodesPerPass);
        const unsigned long long t0 = globaltimer_ns();
        for (unsigned spins = 1;; ++spins) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(w) : "memory");
            if (v == x.seq) break;
            if ((spins & 1023u) == 0 && globaltimer_ns() - t0 > 20000000000ull) {
                *reinterpret_cast<volatile unsigned int*>(x.status) = 1u;
                __threadfence_system();
                break;
            }
        }
    }
    named_bar_sync(bar_id, (int) nthreads);
    const double* mine = x.peers[x.rank] + slot_base;
    if (x.gather_per) {  // node split: output node tid was scored by rank tid / per as its column tid % per
        if (tid < x.gather_total) {
            const unsigned r = tid / x.gather_per, c = tid - r * x.gather_per;
            S = ld_cg_f64(mine + (size_t) r * kXchgRec + c);
            Q = ld_cg_f64(mine + (size_t) r * kXchgRec + kMaxNodesPerPass + c);
        }
    } else if (tid < B) {
        double s = 0.0, q = 0.0;
        for (unsigned r = 0; r < G; ++r) {
            s += ld_cg_f64(mine + (size_t) r * kXchgRec + tid);
            q += ld_cg_f64(mine + (size_t) r * kXchgRec + kMaxNodesPerPass + tid);
        }
        S = s;
        Q = q;
    }
    return make_double2(S, Q);
}

}  // namespace pf
