// pf_pointwise.cu -- frontier kernels for the models whose item log-probability is a function
// of one data row and the node's theta:
//
//   normal target   normalsampler.hh:128-154   lp_i = -1/2 (mu-x_i)' Winv (mu-x_i) + scale
//   mixture         pymixture.py:86-92         lp_item = sum_rows log sum_k exp(-|theta_k - x_r|^2 / 2)
//   scalar Boltzmann boltzmannsampler.hh:71-82 closed form, no data
//
// and per node the two numbers a reference worker accumulates (worker.hh:148-155):
//   metric_sum = sum_items lp_item,   metric_sum_sq = sum_items lp_item^2.
//
// DESIGN (B200).  The data set is read from HBM exactly once per launch, whatever the
// frontier size B.  Each CTA owns a contiguous row range, streamed through a 4-stage shared-memory
// ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier tx counts) issued by one elected lane
// (PF_PW_FOLD; a dedicated producer warp otherwise), and eight consumer warps score every staged row
// against every node.  Lanes are split
// NODES x POINT-LANES: lane l scores node (l mod NB) on the rows of point-lane (l / NB), so theta
// lives in registers, row reads are shared-memory broadcasts, and per-node accumulators never
// leave the lane.  For B > 32 each lane carries two nodes.  Reductions are fixed-order (warp
// shuffle -> per-warp slots -> per-CTA record -> last-CTA finalize), so results are bit-stable
// run to run; there are no floating-point atomics.
//
// ITEMS.  For the mixture an item is a batch of rows and metric_sum_sq needs the item totals
// before squaring.  Tiles never straddle an item, a per-tile block reduction gives the tile
// total, and one owner thread per node assembles item totals; items cut by a CTA boundary are
// handed to the finalize step as head/tail partials and stitched there in CTA order.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <type_traits>

#include "pf_common.cuh"
#include "pf_engine.h"
#include "pf_math.cuh"

namespace pf {

#ifndef PF_PW_WARPS
#define PF_PW_WARPS 8
#endif
// Tuning knobs of the FP64 mixture loops, with the measurements behind the defaults (N = 1e8, K = 8, d = 2, one B200;
// ms per step at B = 8 / 1 / 32).  The loop is bound by REGISTERS, not by instruction count: 16 warps per SM leave 128
// registers per thread, the theta-derived constants of the factored form take 48 of them, and every row in flight ~20
// more, so fewer rows per iteration and fewer resident constants schedule better (ncu: short-scoreboard and fixed-latency
// stalls at the consumers of the table reads fall from 2.9 + 2.8 to ~1 warp per issue slot):
//   rows per iteration 4, A_k and B_k in registers, two partial sums      5.90 / 0.995 / 21.9
//   rows per iteration 4, A_k re-read from shared memory per row          5.44 / 0.965
//   ... and a single accumulation chain                                    5.35 / 0.950
//   rows per iteration 2, A_k re-read, single chain  (the default)         5.24 / 0.927 / 19.7
//   rows per iteration 3                                                   5.21 / 0.941 / 19.6
//   rows per iteration 1                                                   6.67 / 0.989 / 26.0
//   A_k and B_k both re-read (rows 2)                                      5.24 / 0.927 / 19.7
#ifndef PF_MIX_RB
#define PF_MIX_RB 2       // rows per iteration of the mixture loops
#endif
#ifndef PF_MIX_RELOAD
#define PF_MIX_RELOAD 1   // factored form: 0 = theta-derived constants wherever the compiler puts them (registers),
#endif                    // 1 = A_k re-read from shared memory for every row, 2 = A_k and B_k
#ifndef PF_MIX_TWO_ACC
#define PF_MIX_TWO_ACC 0  // factored mixture form: two partial sums per row (1) or a single accumulation chain (0)
#endif
constexpr int PW_CONSUMER_WARPS = PF_PW_WARPS;  // 8 warps = 256 threads: two CTAs per SM fit at up to 128 registers
constexpr int PW_CONSUMERS = PW_CONSUMER_WARPS * 32;
// PF_PW_FOLD: no dedicated producer warp -- lane 0 of consumer warp 0 issues the TMA copies between its rows.
// The register file is per SM sub-partition (16K registers each, warps dealt round-robin), so 2 CTAs x (7 + 1)
// warps leave two sub-partitions with 4 consumer warps and two with 3 + an idle producer: the loaded ones
// bound the kernel at 7/8 of its rate.  2 CTAs x 8 consumer warps put 4 on every sub-partition at 128 registers.
// Measured (mixture N = 1e8): 7 + 1 warps 7.28 ms (B = 8) / 55.8 ms (B = 64); 8 folded 6.47 / 49.1 ms.
#ifndef PF_PW_FOLD
#define PF_PW_FOLD 1
#endif
constexpr bool PW_FOLD = PF_PW_FOLD != 0;
constexpr int PW_THREADS = PW_CONSUMERS + (PW_FOLD ? 0 : 32);  // + one producer warp
#ifndef PF_PW_STAGES
#define PF_PW_STAGES 3  // 16 KB tiles in flight per CTA (a tile is ~20 us of mixture work, its copy ~2 us)
#endif
constexpr int PW_STAGES = PF_PW_STAGES;
constexpr int PW_TILE_BYTES = 16384;
constexpr int PW_STAGE_BYTES = PW_TILE_BYTES + 32;  // room for the 16-byte aligned copy window
constexpr int PW_RED_BUFS = 4;  // >= PW_STAGES: a warp can never lead the owner warp by more than PW_STAGES - 1 tiles
static_assert(PW_RED_BUFS >= PW_STAGES, "tile-total buffers rotate at least as slowly as the stages");
constexpr int PW_BAR_TILE0 = 1;  // named barriers 1..4: per-tile partials ready; 5: CTA-wide among consumers
constexpr int PW_BAR_ALL = 5;
constexpr double kLog2Pi = 1.837877066409345483560659472810;  // normalsampler.hh:24-26

template <typename T> struct ComputeOf { using type = double; };
template <> struct ComputeOf<float> { using type = float; };


// ---------------------------------------------------------------------------------------------
// row loader: DIM values of T from shared memory (alignment = gcd(16, DIM*sizeof(T)))
// ---------------------------------------------------------------------------------------------
template <typename T, int DIM, typename CT>
__device__ __forceinline__ void load_row(const unsigned char* p, CT (&x)[DIM]) {
    constexpr int RB = DIM * (int) sizeof(T);
    T tmp[DIM];
    if constexpr (RB % 16 == 0) {
#pragma unroll
        for (int i = 0; i < RB / 16; ++i) reinterpret_cast<int4*>(tmp)[i] = reinterpret_cast<const int4*>(p)[i];
    } else if constexpr (RB % 8 == 0) {
#pragma unroll
        for (int i = 0; i < RB / 8; ++i) reinterpret_cast<int2*>(tmp)[i] = reinterpret_cast<const int2*>(p)[i];
    } else {
#pragma unroll
        for (int i = 0; i < RB / 4; ++i) reinterpret_cast<int*>(tmp)[i] = reinterpret_cast<const int*>(p)[i];
    }
#pragma unroll
    for (int i = 0; i < DIM; ++i) x[i] = (CT) tmp[i];
}

// ---------------------------------------------------------------------------------------------
// Model policies
// ---------------------------------------------------------------------------------------------
// Per-lane state a model carries through the rows of one tile besides its sums (see pw_main)
struct NoRowState {
    __device__ __forceinline__ void reset() {}
};

// Gaussian target.  theta = D means + packed lower-triangular covariance (normaltheta.hh:81-87).
// prepare() of the reference (normalsampler.hh:87-113): identity covariance -> scale = -D log(2pi)/2;
// otherwise LU-invert the covariance, scale = -(D log 2pi + log det)/2, and evaluate() uses the
// UPPER triangle of the inverse as a symmetric matrix (dsymv CblasUpper, :145).
template <int D, typename T, bool GENERAL>
struct NormalModel {
    static constexpr int DIM = D;
    static constexpr bool ITEM1 = true;
    static constexpr bool HAS_REDO = false;
    static constexpr bool HAS_FIXUP = false;
    static constexpr bool THETA_IN_SMEM = false;
    using CT = typename ComputeOf<T>::type;
    using RowState = NoRowState;
    struct Node {
        CT mu[D];
        CT w[GENERAL ? D * (D + 1) / 2 : 1];  // upper triangle of Winv, row-major, off-diagonals doubled
        CT scale;
    };
    static __device__ void load(Node& nd, const double* th, unsigned) {
#pragma unroll
        for (int i = 0; i < D; ++i) nd.mu[i] = (CT) th[i];
        if constexpr (!GENERAL) {
            nd.w[0] = 0;
            nd.scale = (CT) (-D * kLog2Pi / 2);
        } else {
            double A[D][D], inv[D][D];
            int perm[D];
            for (int i = 0, idx = 0; i < D; ++i)
                for (int j = 0; j <= i; ++j, ++idx) A[i][j] = A[j][i] = th[D + idx];
            int signum = 1;
            for (int i = 0; i < D; ++i) perm[i] = i;
            for (int j = 0; j + 1 < D; ++j) {  // partial pivoting, as gsl_linalg_LU_decomp
                double mx = fabs(A[j][j]);
                int piv = j;
                for (int i = j + 1; i < D; ++i)
                    if (fabs(A[i][j]) > mx) { mx = fabs(A[i][j]); piv = i; }
                if (piv != j) {
                    for (int k = 0; k < D; ++k) { double t = A[j][k]; A[j][k] = A[piv][k]; A[piv][k] = t; }
                    int t = perm[j]; perm[j] = perm[piv]; perm[piv] = t;
                    signum = -signum;
                }
                double ajj = A[j][j];
                if (ajj != 0.0)
                    for (int i = j + 1; i < D; ++i) {
                        double aij = A[i][j] / ajj;
                        A[i][j] = aij;
                        for (int k = j + 1; k < D; ++k) A[i][k] -= aij * A[j][k];
                    }
            }
            for (int c = 0; c < D; ++c) {
                double x[D];
                for (int i = 0; i < D; ++i) x[i] = (perm[i] == c) ? 1.0 : 0.0;
                for (int i = 0; i < D; ++i)
                    for (int k = 0; k < i; ++k) x[i] -= A[i][k] * x[k];
                for (int i = D - 1; i >= 0; --i) {
                    for (int k = i + 1; k < D; ++k) x[i] -= A[i][k] * x[k];
                    x[i] /= A[i][i];
                }
                for (int i = 0; i < D; ++i) inv[i][c] = x[i];
            }
            double det = (double) signum;
            for (int i = 0; i < D; ++i) det *= A[i][i];
            for (int i = 0, idx = 0; i < D; ++i)
                for (int j = i; j < D; ++j, ++idx) nd.w[idx] = (CT) (i == j ? inv[i][j] : 2.0 * inv[i][j]);
            nd.scale = (CT) (-(D * kLog2Pi + log(det)) / 2);
        }
    }
    static __device__ __forceinline__ CT rowlp(const Node& nd, const CT (&x)[D]) {
        CT t[D];
#pragma unroll
        for (int i = 0; i < D; ++i) t[i] = nd.mu[i] - x[i];
        CT dot = 0;
        if constexpr (!GENERAL) {
#pragma unroll
            for (int i = 0; i < D; ++i) dot = fma(t[i], t[i], dot);
        } else {
            int idx = 0;
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = i; j < D; ++j, ++idx) dot = fma(nd.w[idx] * t[i], t[j], dot);
        }
        return fma((CT) -0.5, dot, nd.scale);
    }
};

// Gaussian mixture with unit-variance spherical components, unnormalised (pymixture.py:86-92).
// The reference multiplies d one-dimensional exps per component; exp of the summed exponent is
// the same number to rounding (<= d ulp) and costs one exp instead of d.
//
// theta lives in a padded shared-memory copy ([node][K*D + pad], pad chosen so that the node rows
// start 16 bytes apart modulo 128: conflict-free LDS.128 for up to 8 nodes per warp-load), NOT in
// registers: the registers go to G interleaved exp chains per lane, which is what keeps the FP64
// pipe fed (a single dependent DFMA chain only issues once per pipe latency).
template <typename CT>
__host__ __device__ constexpr unsigned theta_smem_stride(unsigned theta_len) {
    // elements; rows start (stride * sizeof(CT)) apart: make that == 16 (mod 128) bytes
    const unsigned per16 = 16 / (unsigned) sizeof(CT);
    unsigned s = (theta_len + per16 - 1) / per16 * per16;
    while ((s * (unsigned) sizeof(CT)) % 128 != 16) s += per16;
    return s;
}

// FACT = false: the clamped squared-distance form (any theta, any data; slow path for rows out of range).
// FACT = true (FP64 only): the factored form  ln sum_k 2^(A_k + B_k.x) - |x|^2/2  of pf_math.cuh, chosen by the kernel
// itself when every node's theta passes the range test against the data set's bounding box (pw_stage_theta).
template <int K, int D, typename T, bool FACT = false>
struct MixtureModel {
    static constexpr int DIM = D;
    static constexpr bool ITEM1 = false;
    static constexpr bool THETA_IN_SMEM = true;
    using CT = typename ComputeOf<T>::type;
    static constexpr bool BIG_EXP_TABLE = sizeof(CT) == 8;  // FP64: 2048-entry table + mix_exp_accumulate_*
    static constexpr bool FACTORED = FACT && BIG_EXP_TABLE;
    static constexpr bool HAS_REDO = !FACTORED;
    static constexpr bool HAS_FIXUP = BIG_EXP_TABLE;       // n ln 2 (and -|x|^2/2) added once per tile
    static constexpr int NCOMP = K;
    static constexpr unsigned RB = BIG_EXP_TABLE ? PF_MIX_RB : 4;  // rows per loop iteration (FP32: MUFU-bound, 4 is best)
    using Factored = MixtureModel<K, D, T, true>;           // the variant pw_frontier_body may switch to
    // doubles staged per node: theta as it is, or [K][D] B then [K] A
    static __host__ __device__ constexpr unsigned staged_len(unsigned theta_len) {
        return FACTORED ? theta_len + K : theta_len;
    }
    struct RowState {
        unsigned redo;  // slow_path_accumulate() key, tested once per tile
        int nacc;       // sum of the rows' binary exponents (log_pos_acc)
        double xsq;     // sum of the rows' |x|^2 (factored form)
        __device__ __forceinline__ void reset() {
            redo = 0u;
            nacc = 0;
            xsq = 0.0;
        }
    };
    struct Node {
        const CT* th;
        unsigned tab_s;       // shared-memory ADDRESS of the exp table (FP64: 2048 adjusted entries; float: unused)
        unsigned ltab_s;      // ... and of the log table
        const double2* ltab;  // {1/c_j, log c_j} in shared memory
    };
    static __device__ void load(Node& nd, const CT* th_smem, unsigned, const double* tab, const double2* ltab) {
        nd.th = th_smem;
        // FP64: this lane's copy of the replicated table (bank pair = lane mod 16: every lookup is conflict free)
        nd.tab_s = smem_u32(tab) + (BIG_EXP_TABLE ? (threadIdx.x & (kExpRepCopies - 1)) * 8u : 0u);
        nd.ltab_s = smem_u32(ltab);
        nd.ltab = ltab;
    }
    // library exp/log: subnormal terms, log(0) = -inf, nan propagation exactly like numpy's
    static __device__ __noinline__ CT slow(const CT* th, const CT (&x)[D]) {
        double lp = 0;
        for (int k = 0; k < K; ++k) {
            double e = 0;
            for (int j = 0; j < D; ++j) {
                const double t = (double) th[k * D + j] - (double) x[j];
                e = fma(t, t, e);
            }
            lp += exp(-0.5 * e);
        }
        return (CT) log(lp);
    }
    // exp chains kept in flight per lane (bounded by the registers the row operands need)
#ifdef PF_MIX_G
    static constexpr int GMAX = PF_MIX_G;
#else
    static constexpr int GMAX = D <= 2 ? 8 : (D <= 4 ? 4 : 2);
#endif
    static constexpr int G = (K % GMAX == 0) ? GMAX : ((K % 2 == 0) ? 2 : 1);

    static __device__ __forceinline__ CT mix_sum(const Node& nd, const CT (&x)[D]) {
        if constexpr (FACTORED) {
            double lp0 = 0.0, lp1 = 0.0;
#pragma unroll
            for (int k0 = 0; k0 < K; k0 += G) {
                double z[G];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    double b[D], a;
#if PF_MIX_RELOAD >= 2
                    static_assert(D % 2 == 0, "reload variant: even d");
#pragma unroll
                    for (int j = 0; j < D; j += 2)
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(b[j]), "=d"(b[j + 1])
                                     : "r"(smem_u32(nd.th) + (unsigned) (((k0 + g) * D + j) * 8)));
#else
                    load_row<double, D, double>(reinterpret_cast<const unsigned char*>(nd.th + (k0 + g) * D), b);
#endif
#if PF_MIX_RELOAD >= 1
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(smem_u32(nd.th) + (unsigned) ((K * D + k0 + g) * 8)));
#else
                    a = nd.th[K * D + k0 + g];
#endif
                    z[g] = fma(b[0], x[0], a);
#pragma unroll
                    for (int j = 1; j < D; ++j) z[g] = fma(b[j], x[j], z[g]);
                }
                mix_exp_accumulate_z<G, PF_MIX_TWO_ACC != 0>(z, lp0, lp1, nd.tab_s);
            }
            return PF_MIX_TWO_ACC ? lp0 + lp1 : lp0;
        } else if constexpr (BIG_EXP_TABLE) {
            double lp0 = 0.0, lp1 = 0.0;
#pragma unroll
            for (int k0 = 0; k0 < K; k0 += G) {
                double e[G];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    double th[D];
                    load_row<double, D, double>(reinterpret_cast<const unsigned char*>(nd.th + (k0 + g) * D), th);
                    const double t0 = th[0] - x[0];
                    e[g] = t0 * t0;
#pragma unroll
                    for (int j = 1; j < D; ++j) {
                        const double t = th[j] - x[j];
                        e[g] = fma(t, t, e[g]);
                    }
                }
                mix_exp_accumulate_e<G>(e, lp0, lp1, nd.tab_s);
            }
            return (CT) (lp0 + lp1);
        } else {
            CT lp = 0;
#pragma unroll
            for (int k0 = 0; k0 < K; k0 += G) {
                CT e[G], v[G];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    CT th[D];
                    load_row<CT, D, CT>(reinterpret_cast<const unsigned char*>(nd.th + (k0 + g) * D), th);
                    e[g] = 0;
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        const CT t = th[j] - x[j];
                        e[g] = fma(t, t, e[g]);
                    }
                }
                exp_neg_half_fast_batch_s<G>(e, v, nd.tab_s);
#pragma unroll
                for (int w = 1; w < G; w *= 2)  // pairwise: depth log2(G) instead of G dependent adds
#pragma unroll
                    for (int g = 0; g + w < G; g += 2 * w) v[g] += v[g + w];
                lp += v[0];
            }
            return lp;
        }
    }
    // Hot loop: branch free.  Clamped form: a row whose sum left the fast path's range (every component ~40 sigma away,
    // or non-finite input) contributes garbage here and raises `st.redo`; the caller then DISCARDS this lane's sums of
    // the tile and re-walks its rows with rowlp_redo (fast value, or the library for the rows out of range), so
    // the library call -- whose ABI clobbers the uniform registers holding the polynomial constants -- and the
    // masking selects stay out of the loop the FP64 pipe lives in.  Factored form: nothing can leave the range.
    // FP64: the value returned lacks n ln 2 (and -|x|^2/2): tile_fixup() adds those once per tile.
    static __device__ __forceinline__ CT rowlp_fast(const Node& nd, const CT (&x)[D], RowState& st) {
        const CT lp = mix_sum(nd, x);
        if constexpr (FACTORED) {
#pragma unroll
            for (int j = 0; j < D; ++j) st.xsq = fma(x[j], x[j], st.xsq);
            return log_pos_acc(lp, nd.ltab_s, st.nacc);
        } else {
            slow_path_accumulate(st.redo, lp);  // the value is NOT masked (see above)
            if constexpr (BIG_EXP_TABLE)
                return log_pos_acc(lp, nd.ltab_s, st.nacc);
            else
                return log_pos_fast(lp, nd.ltab);
        }
    }
    static __device__ __forceinline__ bool needs_redo(const RowState& st) {
        if constexpr (HAS_REDO) return acc_needs_slow_path(st.redo);
        return false;
    }
    static __device__ __forceinline__ void tile_fixup(const RowState& st, double& acc) {
        if constexpr (FACTORED)
            acc = log_n_times_ln2(st.nacc, fma(-0.5, st.xsq, acc));
        else if constexpr (BIG_EXP_TABLE)
            acc = log_n_times_ln2(st.nacc, acc);
    }
    // complete value of one row (cold path)
    static __device__ CT rowlp_redo(const Node& nd, const CT (&x)[D]) {
        const CT lp = mix_sum(nd, x);
        if (needs_slow_path(lp)) return slow(nd.th, x);
        if constexpr (BIG_EXP_TABLE) {
            int n = 0;
            const double rest = log_pos_acc(lp, nd.ltab_s, n);
            return log_n_times_ln2(n, rest);
        } else
            return log_pos_fast(lp, nd.ltab);
    }
};

// Any K (<= 32): same, with a run-time component loop.  Fallback for shapes without a K-unrolled
// instantiation.
template <int D, typename T>
struct MixtureModelRT {
    static constexpr int DIM = D;
    static constexpr bool ITEM1 = false;
    static constexpr bool THETA_IN_SMEM = true;
    static constexpr bool HAS_REDO = true;
    static constexpr bool HAS_FIXUP = false;
    static constexpr bool BIG_EXP_TABLE = false;
    static constexpr bool FACTORED = false;
    static constexpr unsigned RB = 4;
    using CT = typename ComputeOf<T>::type;
    static __host__ __device__ constexpr unsigned staged_len(unsigned theta_len) { return theta_len; }
    struct RowState {
        unsigned redo;
        __device__ __forceinline__ void reset() { redo = 0u; }
    };
    static __device__ __forceinline__ bool needs_redo(const RowState& st) { return acc_needs_slow_path(st.redo); }
    struct Node {
        const CT* th;
        const double* tab;
        const double2* ltab;
        unsigned K;
    };
    static __device__ void load(Node& nd, const CT* th_smem, unsigned K, const double* tab, const double2* ltab) {
        nd.th = th_smem;
        nd.tab = tab;
        nd.ltab = ltab;
        nd.K = K;
    }
    static __device__ __forceinline__ CT mix_sum(const Node& nd, const CT (&x)[D]) {
        CT lp = 0;
        for (unsigned k = 0; k < nd.K; ++k) {
            CT e = 0;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const CT t = nd.th[k * D + j] - x[j];
                e = fma(t, t, e);
            }
            lp += exp_neg_half_fast(e, nd.tab);
        }
        return lp;
    }
    static __device__ __forceinline__ CT rowlp_fast(const Node& nd, const CT (&x)[D], RowState& st) {
        const CT lp = mix_sum(nd, x);
        slow_path_accumulate(st.redo, lp);
        return log_pos_fast(lp, nd.ltab);
    }
    static __device__ CT rowlp_redo(const Node& nd, const CT (&x)[D]) {
        const CT lp = mix_sum(nd, x);
        if (!needs_slow_path(lp)) return log_pos_fast(lp, nd.ltab);
        double l2 = 0;
        for (unsigned k = 0; k < nd.K; ++k) {
            double e = 0;
            for (int j = 0; j < D; ++j) {
                const double t = (double) nd.th[k * D + j] - (double) x[j];
                e = fma(t, t, e);
            }
            l2 += exp(-0.5 * e);
        }
        return (CT) log(l2);
    }
};

// ---------------------------------------------------------------------------------------------
// tile walk shared by producer and consumers: contiguous rows, cut at item boundaries
// ---------------------------------------------------------------------------------------------
struct TileIter {
    unsigned long long cur, end_all, item_rows, row_end;
    unsigned tile_rows;
    unsigned long long item_end;  // end of the item `cur` lies in (0: not computed yet) -- one division per CTA,
                                  // not one per tile: at B = 1 a tile is only four row iterations per warp
    __device__ __forceinline__ bool next(unsigned long long& t0, unsigned long long& t1) {
        if (cur >= end_all) return false;
        t0 = cur;
        unsigned long long e = cur + tile_rows;
        if (e > end_all) e = end_all;
        if (item_rows > 1) {
            if (item_end == 0) item_end = (cur / item_rows + 1) * item_rows;
            const unsigned long long ie = item_end < row_end ? item_end : row_end;
            if (e >= ie) {
                e = ie;
                item_end += item_rows;
            }
        }
        t1 = e;
        cur = e;
        return true;
    }
};

// Register budget: 256 threads per CTA, so two CTAs per SM fit at up to 128 registers per thread
// (registers are granted per pair of warps: 2 x 8 x 32 x 128 = 64K).  Models whose per-lane
// state is large get the whole register file of one CTA instead.
// per-item models rotate 4 tile buffers; one-row-per-item models reduce once at the end (2 x warps x nodes)
template <class M, class = void>
struct BigExpTable { static constexpr bool value = false; };
template <class M>
struct BigExpTable<M, std::enable_if_t<M::BIG_EXP_TABLE>> { static constexpr bool value = true; };
template <class M>
__host__ __device__ constexpr bool big_exp_table() { return BigExpTable<M>::value; }

// per-warp slots of the block reductions: [buffer][warp][32 * NPL nodes]; the finalize step reuses the area as scratch
// (finalize_items: 2 doubles per thread; peer_allreduce: 2 * kMaxNodesPerPass)
template <class M, int NPL>
__host__ __device__ constexpr int pw_red_doubles() {
    constexpr int n = (M::ITEM1 ? 2 : PW_RED_BUFS) * PW_CONSUMER_WARPS * 32 * NPL;
    return n > 2 * PW_CONSUMERS ? n : 2 * PW_CONSUMERS;
}

template <class M, int NPL>
constexpr bool pw_two_ctas() {
    return sizeof(typename M::Node) * NPL <= 160;
}

template <class M, typename T, int NPL>
__device__ __forceinline__ void pw_frontier_body(const PointwiseArgs& a);

#ifndef PF_PW_BLOCKS
#define PF_PW_BLOCKS 2
#endif
#ifndef PF_PW_MAXOCC
#define PF_PW_MAXOCC 2  // resident CTAs per SM actually used (a third one was measured slower: 80 vs 68 us, normal B=64)
#endif
template <class M, typename T, int NPL>
__global__ void __launch_bounds__(PW_THREADS, PF_PW_BLOCKS) pw_frontier_kernel2(const __grid_constant__ PointwiseArgs a) {
    pw_frontier_body<M, T, NPL>(a);
}

template <class M, typename T, int NPL>
__global__ void __launch_bounds__(PW_THREADS, 1) pw_frontier_kernel1(const __grid_constant__ PointwiseArgs a) {
    pw_frontier_body<M, T, NPL>(a);
}

// what pw_frontier_body carves out of shared memory and hands to pw_main
struct PwShared {
    unsigned char* stage_base;
    double* red;
    uint64_t *full, *empty;
    int* flag;
    void* th_smem;
    const double* exp2_tab;
    const double2* log_tab;
    unsigned long long r0, r1;
    const unsigned long long* theta_nan;
};

template <class M, class = void>
struct FactoredOf { using type = void; };
template <class M>
struct FactoredOf<M, std::enable_if_t<M::BIG_EXP_TABLE && !M::FACTORED>> { using type = typename M::Factored; };

// doubles of shared memory per staged node (the larger of the two layouts a model may use)
template <class M>
__host__ __device__ constexpr unsigned pw_staged_len(unsigned theta_len) {
    if constexpr (M::THETA_IN_SMEM) {
        if constexpr (!std::is_void<typename FactoredOf<M>::type>::value)
            return FactoredOf<M>::type::staged_len(theta_len);
    }
    return theta_len;
}

template <class M, typename T, int NPL>
__device__ __forceinline__ void pw_main(const PointwiseArgs& a, const PwShared& sh);

template <class M, typename T, int NPL>
__device__ __forceinline__ void pw_frontier_body(const PointwiseArgs& a) {
    using CT = typename M::CT;
    constexpr int DIM = M::DIM;
    constexpr int RB_BYTES = DIM * (int) sizeof(T);
    constexpr unsigned TILE_ROWS = PW_TILE_BYTES / RB_BYTES;

    extern __shared__ __align__(128) unsigned char smem[];
    PwShared sh;
    sh.stage_base = smem;
    sh.red = reinterpret_cast<double*>(smem + PW_STAGES * PW_STAGE_BYTES);
    sh.full = reinterpret_cast<uint64_t*>(sh.red + pw_red_doubles<M, NPL>());
    sh.empty = sh.full + PW_STAGES;
    sh.flag = reinterpret_cast<int*>(sh.empty + PW_STAGES);
    CT* th_smem = reinterpret_cast<CT*>(reinterpret_cast<unsigned char*>(sh.flag) + 16);  // 16-byte aligned
    sh.th_smem = th_smem;
    sh.exp2_tab = nullptr;
    sh.log_tab = nullptr;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned long long r0 = a.row_begin + (unsigned long long) blockIdx.x * a.rows_per_cta;
    if (r0 > a.row_end) r0 = a.row_end;
    unsigned long long r1 = r0 + a.rows_per_cta;
    if (r1 > a.row_end) r1 = a.row_end;
    sh.r0 = r0;
    sh.r1 = r1;

    __shared__ unsigned long long theta_nan;  // bit b: node b has a NaN in its theta (see mix_exp_accumulate)
    __shared__ uint64_t tabbar;               // completion of the exp / log table copies (mixture, FP64)
    __shared__ int not_factorable;            // some node's theta fails the range test of the factored form
    sh.theta_nan = &theta_nan;
    const bool tr0 = tid == 0 && blockIdx.x == 0;
    trace_mark(a.trace, 0, 0, tr0);
    trace_cta(a.trace, 0, tid == 0);
    if (tid == 0) {
        for (int s = 0; s < PW_STAGES; ++s) {
            mbar_init(&sh.full[s], 1);
            mbar_init(&sh.empty[s], PW_CONSUMER_WARPS);
        }
        mbar_init(&tabbar, 1);
        fence_barrier_init();
        theta_nan = 0ull;
        not_factorable = 0;
    }
    __syncthreads();

    if (!PW_FOLD && warp == PW_CONSUMER_WARPS) {
        // ------------------------------------------------------------------ producer warp
        if (lane == 0) {
            TileIter it{r0, r1, a.item_rows, a.row_end, TILE_ROWS, 0ull};
            unsigned long long t0, t1;
            unsigned t = 0;
            while (it.next(t0, t1)) {
                const unsigned s = t % PW_STAGES;
                if (t >= PW_STAGES) mbar_wait(&sh.empty[s], ((t / PW_STAGES) - 1) & 1);
                const unsigned long long b0 = t0 * RB_BYTES, b1 = t1 * RB_BYTES;
                const unsigned long long w0 = b0 & ~15ull, w1 = (b1 + 15ull) & ~15ull;
                const unsigned bytes = (unsigned) (w1 - w0);
                mbar_arrive_expect_tx(&sh.full[s], bytes);
                bulk_g2s(sh.stage_base + s * PW_STAGE_BYTES, static_cast<const unsigned char*>(a.data) + w0, bytes,
                         &sh.full[s]);
                ++t;
            }
        }
        return;
    }

    if constexpr (M::THETA_IN_SMEM) {
        // thetas: device memory, or the kernel parameters themselves (__grid_constant__: their address may be taken)
        const double* const theta = a.theta_inline_n ? a.theta_inline : a.theta;
        const unsigned B = a.B;
        // The 2^(j/256) table sits at a shared-memory ADDRESS aligned to its own size (exp_neg_half_fast_batch_s ORs
        // the entry offset into it).  __align__ only aligns the offset behind the 1 KB the hardware reserves at
        // the start of the window, so the aligned 2 KB are picked out of a 4 KB array by hand.
        constexpr bool BIG = big_exp_table<M>();
        __shared__ __align__(128) double exp2_raw[BIG ? kExpRepSize * kExpRepCopies : 2 * kExp2TabSize];
        __shared__ __align__(16) double2 log_tab[256];
        constexpr unsigned kTabBytes = kExp2TabSize * 8;
        const unsigned raw_s = smem_u32(exp2_raw);
        double* const exp2_tab = BIG ? exp2_raw : exp2_raw + ((((raw_s + kTabBytes - 1) & ~(kTabBytes - 1)) - raw_s) >> 3);
        sh.exp2_tab = exp2_tab;
        sh.log_tab = log_tab;
        if constexpr (BIG) {
            // 32 KB (256 entries x 16 copies) + 4 KB by two TMA bulk copies
            if (tid == 0) {
                constexpr unsigned kRepBytes = (unsigned) (sizeof(double) * kExpRepSize * kExpRepCopies);
                mbar_arrive_expect_tx(&tabbar, kRepBytes + (unsigned) (sizeof(double2) * 256));
                bulk_g2s(exp2_tab, g_exp2_rep, kRepBytes, &tabbar);
                bulk_g2s(log_tab, g_log_mid, (unsigned) (sizeof(double2) * 256), &tabbar);
            }
        } else {
            for (int i = tid; i < kExp2TabSize; i += PW_CONSUMERS) exp2_tab[i] = kExp2Tab[i];
            for (int i = tid; i < 256; i += PW_CONSUMERS) log_tab[i] = make_double2(kLogTab[2 * i], kLogTab[2 * i + 1]);
        }
        using MF = typename FactoredOf<M>::type;
        [[maybe_unused]] bool factored = false;
        if constexpr (!std::is_void<MF>::value && NPL == 1) {
            // Factored layout per node: [K][D] B = log2(e) theta, then [K] A = -log2(e)/2 |theta_k|^2 -- and the range
            // test that licenses it: |A_k| + sum_j |B_kj| max|x_j| <= kFactoredLimit for every component of every node
            // (a NaN / inf anywhere fails the comparison).  Every CTA reaches the same verdict from the same inputs.
            constexpr unsigned KC = MF::NCOMP;
            const unsigned stride = theta_smem_stride<double>(MF::staged_len(a.theta_len));
            bool bad = false;
            for (unsigned i = tid; i < B * KC; i += PW_CONSUMERS) {
                const unsigned node = i / KC, k = i - node * KC;
                const double* th = theta + (size_t) node * a.theta_len + k * DIM;
                double* dst = reinterpret_cast<double*>(th_smem) + node * stride;
                double nrm = 0.0, bound = 0.0;
#pragma unroll
                for (int j = 0; j < DIM; ++j) {
                    const double v = th[j], bj = v * kLog2e;
                    dst[k * DIM + j] = bj;
                    nrm = fma(v, v, nrm);
                    bound = fma(fabs(bj), a.xmax[j], bound);
                }
                const double A = nrm * kNegHalfLog2e;
                dst[KC * DIM + k] = A;
                bad |= !(fabs(A) + bound <= kFactoredLimit);
            }
            if (bad) atomicOr(&not_factorable, 1);
            named_bar_sync(PW_BAR_ALL, PW_CONSUMERS);
            factored = not_factorable == 0;
        }
        if (!factored) {
            const unsigned stride = theta_smem_stride<CT>(a.theta_len);
            for (unsigned i = tid; i < B * a.theta_len; i += PW_CONSUMERS) {
                const unsigned node = i / a.theta_len, j = i - node * a.theta_len;
                const double v = theta[i];
                th_smem[node * stride + j] = (CT) v;
                if constexpr (BIG)
                    if (v != v) atomicOr(&theta_nan, 1ull << node);
            }
            named_bar_sync(PW_BAR_ALL, PW_CONSUMERS);
        }
        if constexpr (BIG) mbar_wait(&tabbar, 0);
        if (a.trace != nullptr && tr0) a.trace[6] = factored ? 1ull : 2ull;  // diagnostics: which form scores this call
        if constexpr (!std::is_void<MF>::value && NPL == 1) {
            if (factored) {
                pw_main<MF, T, NPL>(a, sh);
                return;
            }
        }
    }
    pw_main<M, T, NPL>(a, sh);
}

template <class M, typename T, int NPL>
__device__ __forceinline__ void pw_main(const PointwiseArgs& a, const PwShared& sh) {
    using CT = typename M::CT;
    constexpr int DIM = M::DIM;
    constexpr int RB_BYTES = DIM * (int) sizeof(T);
    constexpr unsigned TILE_ROWS = PW_TILE_BYTES / RB_BYTES;
    unsigned char* const stage_base = sh.stage_base;
    double* const red = sh.red;
    uint64_t* const full = sh.full;
    uint64_t* const empty = sh.empty;
    int* const flag = sh.flag;
    const unsigned long long r0 = sh.r0, r1 = sh.r1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool tr0 = tid == 0 && blockIdx.x == 0;

    // one tile's copy: rows [t0, t1) -> stage s, completion counted on full[s]
    auto issue_tile = [&](unsigned long long t0, unsigned long long t1, unsigned s) {
        const unsigned long long b0 = t0 * RB_BYTES, b1 = t1 * RB_BYTES;
        const unsigned long long w0 = b0 & ~15ull, w1 = (b1 + 15ull) & ~15ull;
        const unsigned bytes = (unsigned) (w1 - w0);
        mbar_arrive_expect_tx(&full[s], bytes);
        bulk_g2s(stage_base + s * PW_STAGE_BYTES, static_cast<const unsigned char*>(a.data) + w0, bytes, &full[s]);
    };

    // ---------------------------------------------------------------------- consumer warps
    const unsigned B = a.B;
    unsigned NB = 32;  // node lanes (power of two >= min(B, 32))
    if (B < 32) {
        NB = 1;
        while (NB < B) NB <<= 1;
    }
    const unsigned PL = 32 / NB;
    const unsigned nb = lane & (NB - 1), pl = lane / NB;
    const unsigned slot = warp * PL + pl, nslots = PW_CONSUMER_WARPS * PL;

    typename M::Node nd[NPL];
    if constexpr (M::THETA_IN_SMEM) {
        const unsigned stride = theta_smem_stride<CT>(M::staged_len(a.theta_len));
#pragma unroll
        for (int n = 0; n < NPL; ++n) {
            unsigned node = nb + 32 * n;
            if (node >= B) node = B - 1;  // idle lanes shadow the last node; their sums are dropped
            M::load(nd[n], reinterpret_cast<const CT*>(sh.th_smem) + node * stride, a.K, sh.exp2_tab, sh.log_tab);
        }
    } else {
        const double* const theta = a.theta_inline_n ? a.theta_inline : a.theta;
#pragma unroll
        for (int n = 0; n < NPL; ++n) {
            unsigned node = nb + 32 * n;
            if (node >= B) node = B - 1;
            M::load(nd[n], theta + (size_t) node * a.theta_len, a.K);
        }
    }

    double accS[NPL], accQ[NPL];
#pragma unroll
    for (int n = 0; n < NPL; ++n) accS[n] = accQ[n] = 0.0;
    trace_mark(a.trace, 0, 1, tr0);  // prologue done (tables, thetas)

    // owner-thread state (thread b < B assembles node b's item totals)
    ItemOwner own;
    own.init(r0, a.item_rows);

    // folded producer (PW_FOLD): lane 0 of warp 0 fills all PW_STAGES stages up front; when tile t (>= 1) starts it
    // sends tile t + PW_STAGES - 1 into the stage tile t - 1 used.  If the other warps have not released that stage
    // yet, the copy is issued after this warp's rows of tile t instead (a tile is ~20 us of work, a copy ~2 us:
    // no stage ever runs dry, and the issuing warp never waits at the top of a tile).
    const bool issuer = PW_FOLD && tid == 0;
    TileIter pit{r0, r1, a.item_rows, a.row_end, TILE_ROWS, 0ull};
    unsigned pu = 0;  // next tile to issue
    if (issuer) {
        unsigned long long p0, p1;
        while (pu < PW_STAGES && pit.next(p0, p1)) {
            issue_tile(p0, p1, pu % PW_STAGES);
            ++pu;
        }
    }
    auto try_issue = [&](unsigned t, bool block) {
        // tile pu = t + PW_STAGES - 1 reuses the stage of tile t - 1 (its (pu / PW_STAGES - 1)-th release)
        const unsigned s = pu % PW_STAGES;
        const unsigned phase = ((pu / PW_STAGES) - 1) & 1;
        if (block)
            mbar_wait(&empty[s], phase);
        else if (!mbar_try_wait(&empty[s], phase))
            return false;
        unsigned long long p0, p1;
        if (pit.next(p0, p1)) issue_tile(p0, p1, s);
        ++pu;
        return true;
    };

    TileIter it{r0, r1, a.item_rows, a.row_end, TILE_ROWS, 0ull};
    unsigned long long t0, t1;
    unsigned t = 0;
    while (it.next(t0, t1)) {
        const unsigned s = t % PW_STAGES;
        bool issue_pending = false;
        if (issuer && t >= 1 && pu == t + PW_STAGES - 1 && pit.cur < pit.end_all) issue_pending = !try_issue(t, false);
        mbar_wait(&full[s], (t / PW_STAGES) & 1);
        if (t == 0) trace_mark(a.trace, 0, 2, tr0);  // first tile landed
        const unsigned char* base = stage_base + s * PW_STAGE_BYTES + (unsigned) ((t0 * RB_BYTES) & 15ull);
        const unsigned nrows = (unsigned) (t1 - t0);
        // rows in flight per lane: batches of RB rows with every shared-memory load issued before the first use
        // and no exit inside a batch (the one-row-per-item models are short dependent chains behind a 30-cycle LDS)
        // (mixture: several rows per iteration amortise the loop's re-materialised constants and overlap one row's
        // serial log chain with the other rows' exps.  N = 1e8, B = 8: 6.51 ms with one row, 6.14 with two, 5.93 with
        // four; with one node and 32 point-lanes (B = 1) a lane has only four rows of a tile and two at a time are
        // best; with two nodes per lane batching only spills.)
        unsigned j = slot;
        typename M::RowState st[NPL];  // per-tile lane state besides the sums (slow-path key, exponent sum, |x|^2 sum)
#pragma unroll
        for (int n = 0; n < NPL; ++n) st[n].reset();
        auto batch_rows = [&](auto rb_tag) {
            constexpr unsigned RB = decltype(rb_tag)::value;
            for (; j + (RB - 1) * nslots < nrows; j += RB * nslots) {
                CT x[RB][DIM];
#pragma unroll
                for (unsigned u = 0; u < RB; ++u) load_row<T, DIM, CT>(base + (size_t) (j + u * nslots) * RB_BYTES, x[u]);
#pragma unroll
                for (unsigned u = 0; u < RB; ++u)
#pragma unroll
                    for (int n = 0; n < NPL; ++n) {
                        double v;
                        if constexpr (M::HAS_REDO || M::HAS_FIXUP)
                            v = (double) M::rowlp_fast(nd[n], x[u], st[n]);
                        else
                            v = (double) M::rowlp(nd[n], x[u]);
                        accS[n] += v;
                        if constexpr (M::ITEM1) accQ[n] = fma(v, v, accQ[n]);
                    }
            }
        };
        if constexpr (M::ITEM1)
            batch_rows(std::integral_constant<unsigned, 4>{});
        else if constexpr (NPL == 1) {
            if constexpr (M::RB >= 4)
                if (PL <= 8) batch_rows(std::integral_constant<unsigned, 4>{});
            if constexpr (M::RB == 3) batch_rows(std::integral_constant<unsigned, 3>{});
            if constexpr (M::RB >= 2) batch_rows(std::integral_constant<unsigned, 2>{});
        }
        for (; j < nrows; j += nslots) {
            CT x[DIM];
            load_row<T, DIM, CT>(base + (size_t) j * RB_BYTES, x);
#pragma unroll
            for (int n = 0; n < NPL; ++n) {
                double v;
                if constexpr (M::HAS_REDO || M::HAS_FIXUP)
                    v = (double) M::rowlp_fast(nd[n], x, st[n]);
                else
                    v = (double) M::rowlp(nd[n], x);
                accS[n] += v;
                if constexpr (M::ITEM1) accQ[n] = fma(v, v, accQ[n]);
            }
        }
        if (issue_pending) try_issue(t, true);
        if constexpr (M::HAS_REDO) {
            bool again = false;
#pragma unroll
            for (int n = 0; n < NPL; ++n) again |= M::needs_redo(st[n]);
            if (again) {  // cold: discard this lane's sums of the tile (they start at 0 per tile) and walk its rows again
#pragma unroll
                for (int n = 0; n < NPL; ++n) accS[n] = 0.0;
                for (unsigned j2 = slot; j2 < nrows; j2 += nslots) {
                    CT x[DIM];
                    load_row<T, DIM, CT>(base + (size_t) j2 * RB_BYTES, x);
#pragma unroll
                    for (int n = 0; n < NPL; ++n) accS[n] += (double) M::rowlp_redo(nd[n], x);
                }
#pragma unroll
                for (int n = 0; n < NPL; ++n) st[n].reset();  // rowlp_redo returns complete values: nothing left for the fix-up
            }
        }
        if constexpr (M::HAS_FIXUP) {
#pragma unroll
            for (int n = 0; n < NPL; ++n) M::tile_fixup(st[n], accS[n]);
        }
        if constexpr (M::ITEM1) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        } else {
            // Tile total per node: lanes -> warp (shuffles) -> per-warp slot in one of 4 rotating
            // buffers.  Every warp ARRIVES on the tile's named barrier and moves on; only the owner
            // warp(s) -- the ones holding threads tid < B -- WAIT on it, then fold the per-warp slots.  The
            // smem stage is released after the arrival, which bounds any warp's lead to 3 tiles, so a
            // buffer is never rewritten before the owners have read it.
            constexpr int RS = 32 * NPL;  // slot row stride
            double* r = red + (t & (PW_RED_BUFS - 1)) * (PW_CONSUMER_WARPS * RS);
#pragma unroll
            for (int n = 0; n < NPL; ++n) {
                double v = accS[n];
                accS[n] = 0.0;
                for (unsigned m = 16; m >= NB; m >>= 1) v += shfl_xor_f64(v, m);
                if (pl == 0) r[warp * RS + nb + 32 * n] = v;
            }
            const int bar = PW_BAR_TILE0 + (int) (t & (PW_RED_BUFS - 1));
            const bool owner_warp = (unsigned) (warp * 32) < B;
            if (owner_warp)
                named_bar_sync(bar, PW_CONSUMERS);
            else
                named_bar_arrive(bar, PW_CONSUMERS);
            if ((unsigned) tid < B) {
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < PW_CONSUMER_WARPS; ++w) tot += r[w * RS + tid];
                own.add_tile(tot, t1, a.item_rows, a.row_end);
            }
            __syncwarp();  // owners have read the slots before this warp frees the stage
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        ++t;
    }

    trace_mark(a.trace, 0, 3, tr0);  // all tiles scored
    double* rec = a.part + (size_t) blockIdx.x * kPartSlots * kMaxNodesPerPass;
    if constexpr (M::ITEM1) {
        // one block reduction at the very end: every row is its own item
#pragma unroll
        for (int n = 0; n < NPL; ++n) {
            double v = accS[n], q = accQ[n];
            for (unsigned m = 16; m >= NB; m >>= 1) {
                v += shfl_xor_f64(v, m);
                q += shfl_xor_f64(q, m);
            }
            if (pl == 0) {
                red[warp * (32 * NPL) + nb + 32 * n] = v;
                red[PW_CONSUMER_WARPS * (32 * NPL) + warp * (32 * NPL) + nb + 32 * n] = q;
            }
        }
        named_bar_sync(PW_BAR_ALL, PW_CONSUMERS);
        if ((unsigned) tid < B) {
#pragma unroll
            for (int w = 0; w < PW_CONSUMER_WARPS; ++w) {
                own.S += red[w * (32 * NPL) + tid];
                own.Q += red[PW_CONSUMER_WARPS * (32 * NPL) + w * (32 * NPL) + tid];
            }
        }
    } else if ((unsigned) tid < B)
        own.finish(r0, r1, a.row_begin, a.row_end, a.rows_per_cta, a.item_rows);
    if ((unsigned) tid < B) {
        if constexpr (big_exp_table<M>())
            if ((*sh.theta_nan >> tid) & 1ull) own.S = own.Q = __longlong_as_double(0x7FF8000000000000ll);
        own.store(rec, tid);
        __threadfence();
    }
    named_bar_sync(PW_BAR_ALL, PW_CONSUMERS);
    if (tid == 0) {
        const unsigned prev = atomicAdd(a.ticket, 1u);
        flag[0] = (prev == gridDim.x - 1);
    }
    named_bar_sync(PW_BAR_ALL, PW_CONSUMERS);
    trace_mark(a.trace, 0, 4, tr0);  // record published
    trace_cta(a.trace, 1, tid == 0);
    if (!flag[0]) return;

    // ---------------------------------------------------------------------- last CTA: finalize
    trace_mark(a.trace, 1, 0, tid == 0);
    __threadfence();
    {
        double S, Q;
        finalize_items(a.part, gridDim.x, a.row_begin, a.row_end, a.rows_per_cta, M::ITEM1 ? 1ull : a.item_rows, tid,
                       B, red, PW_BAR_ALL, PW_CONSUMERS, S, Q);
        trace_mark(a.trace, 1, 1, tid == 0);  // records folded
        unsigned Bout = B;  // node split: the gathered frontier is wider than this rank's slice
        if (a.xchg.nranks > 1) {
            const double2 t = peer_allreduce(a.xchg, tid, B, PW_CONSUMERS, PW_BAR_ALL, red, S, Q);
            S = t.x;
            Q = t.y;
            if (a.xchg.gather_per) Bout = a.xchg.gather_total;
        }
        if ((unsigned) tid < Bout) {
            const unsigned iS = a.out_off + tid, iQ = (a.out_stride ? a.out_stride : Bout) + a.out_off + tid;
            if (a.done_flag) {  // host path: self-validating words into mapped host memory, no fence, no flag
                store_result_words(a.out, iS, S, a.done_seq);
                store_result_words(a.out, iQ, Q, a.done_seq);
            } else {
                a.out[iS] = S;
                a.out[iQ] = Q;
            }
        }
    }
    if (tid == 0) *a.ticket = 0u;  // ready for the next launch on this stream
    trace_mark(a.trace, 1, 2, tid == 0);
}


// ---------------------------------------------------------------------------------------------
// Partial sums in the reference's per-node visiting order (normal target only): item i of node b
// is data row (start_b + i * stride_b) mod n  (strideindex.hh:46-48; drawn in prepare(),
// normalsampler.hh:115-119).  Every node walks its own permutation, so this is a gather: one
// CTA column per node (blockIdx.y), rows fetched straight from global/L2 -- used for the partial
// updates that feed the predictive scheduler (worker.hh:141-162), never for whole-data scoring.
// Deterministic: per-CTA partials, summed in CTA order by ordered_finalize_kernel.
// ---------------------------------------------------------------------------------------------
template <int D, typename T, bool GENERAL>
__global__ void __launch_bounds__(256) normal_ordered_kernel(const T* __restrict__ data, unsigned long long n,
                                                             const double* __restrict__ theta, unsigned theta_len,
                                                             const unsigned long long* __restrict__ order,
                                                             unsigned long long first, unsigned long long count,
                                                             double* __restrict__ part) {
    using M = NormalModel<D, T, GENERAL>;
    using CT = typename M::CT;
    __shared__ double red[2][8];
    const unsigned b = blockIdx.y;
    typename M::Node nd;
    M::load(nd, theta + (size_t) b * theta_len, 0);
    const unsigned long long stride = order[2 * b], start = order[2 * b + 1];
    double S = 0.0, Q = 0.0;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned long long row = (start + (first + i) * stride) % n;
        CT x[D];
#pragma unroll
        for (int j = 0; j < D; ++j) x[j] = (CT) __ldg(&data[row * D + j]);
        const double v = (double) M::rowlp(nd, x);
        S += v;
        Q = fma(v, v, Q);
    }
    for (int m = 16; m >= 1; m >>= 1) {
        S += shfl_xor_f64(S, m);
        Q += shfl_xor_f64(Q, m);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red[0][warp] = S;
        red[1][warp] = Q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0, q = 0.0;
        for (int w = 0; w < 8; ++w) {
            s += red[0][w];
            q += red[1][w];
        }
        part[((size_t) b * gridDim.x + blockIdx.x) * 2] = s;
        part[((size_t) b * gridDim.x + blockIdx.x) * 2 + 1] = q;
    }
}

__global__ void ordered_finalize_kernel(const double* part, unsigned nblocks, unsigned B, double* out) {
    const unsigned b = threadIdx.x;
    if (b >= B) return;
    double s = 0.0, q = 0.0;
    for (unsigned c = 0; c < nblocks; ++c) {
        s += part[((size_t) b * nblocks + c) * 2];
        q += part[((size_t) b * nblocks + c) * 2 + 1];
    }
    out[b] = s;
    out[B + b] = q;
}

// scalar Boltzmann (boltzmannsampler.hh:71-82): x = -theta/T; sum = x*count; sum_sq = x*x*count
__global__ void boltzmann_scalar_kernel(const double* theta, double* out, unsigned B, double count, double inv_t) {
    const unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        const double x = -theta[b] * inv_t;
        out[b] = x * count;
        out[B + b] = x * x * count;
    }
}

// diagnostics: apply the transcendental kernels elementwise (tests/test_gpu_math.py)
__global__ void debug_math_kernel(int which, const double* in, double* out, size_t n) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    __shared__ double tab[kExp2TabSize];
    __shared__ __align__(16) double2 ltab[256];
    for (int i = threadIdx.x; i < kExp2TabSize; i += blockDim.x) tab[i] = kExp2Tab[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) ltab[i] = make_double2(kLogTab[2 * i], kLogTab[2 * i + 1]);
    __syncthreads();
    if (i >= n) return;
    const double x = in[i];
    double r;
    switch (which) {
        case 0: r = exp_neg_half(x, tab); break;
        case 1: r = log_pos(x); break;
        case 2: r = (double) exp_neg_half((float) x, tab); break;
        case 4: r = log_pos_tab(x, ltab); break;
        default: r = (double) log_pos((float) x); break;
    }
    out[i] = r;
}

// diagnostics for the mixture loop's own pair (mix_exp_accumulate / log_pos_fast_k): which = 5: exp(-x/2), 6: log(x)
__global__ void debug_math_kernel_big(int which, const double* in, double* out, size_t n) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    __shared__ __align__(128) double tab[kExpRepSize * kExpRepCopies];  // (>= kExpBigSize)
    __shared__ __align__(16) double2 ltab[256];
    if (which >= 7)
        for (int k = threadIdx.x; k < kExpRepSize * kExpRepCopies; k += blockDim.x) tab[k] = g_exp2_rep[k];
    else
        for (int k = threadIdx.x; k < kExpBigSize; k += blockDim.x) tab[k] = g_exp2_big[k];
    for (int k = threadIdx.x; k < 256; k += blockDim.x)
        ltab[k] = which == 8 ? make_double2(g_log_mid[2 * k], g_log_mid[2 * k + 1]) : make_double2(g_log_adj[2 * k], g_log_adj[2 * k + 1]);
    __syncthreads();
    if (i >= n) return;
    double r;
    if (which == 5) {
        double e[1] = {in[i]}, lp0 = 0.0, lp1 = 0.0;
        mix_exp_accumulate<1>(e, lp0, lp1, smem_u32(tab));
        r = lp0 + lp1;
    } else if (which == 7) {  // the factored loop's exp: 2^z
        double z[1] = {in[i]}, lp0 = 0.0, lp1 = 0.0;
        mix_exp_accumulate_z<1, true>(z, lp0, lp1, smem_u32(tab) + (threadIdx.x & 15) * 8u);
        r = lp0 + lp1;
    } else if (which == 9) {  // the clamped loop's exp, degree 2: exp(-x/2)
        double e[1] = {in[i]}, lp0 = 0.0, lp1 = 0.0;
        mix_exp_accumulate_e<1>(e, lp0, lp1, smem_u32(tab) + (threadIdx.x & 15) * 8u);
        r = lp0 + lp1;
    } else if (which == 8) {  // log with the exponent accumulated apart
        int n = 0;
        const double rest = log_pos_acc(in[i], smem_u32(ltab), n);
        r = log_n_times_ln2(n, rest);
    } else
        r = log_pos_fast_k(in[i], smem_u32(ltab));
    out[i] = r;
}

// ---------------------------------------------------------------------------------------------
// host side: dispatch
// ---------------------------------------------------------------------------------------------
namespace {

constexpr size_t kPwSmemFixed = (size_t) PW_STAGES * PW_STAGE_BYTES + 2 * PW_STAGES * sizeof(uint64_t) + 32;

template <class M, int NPL>
size_t pw_smem_bytes(unsigned theta_len, unsigned max_nodes) {
    const size_t base = kPwSmemFixed + (size_t) pw_red_doubles<M, NPL>() * sizeof(double);
    if (!M::THETA_IN_SMEM) return base;
    return base + (size_t) max_nodes * theta_smem_stride<typename M::CT>(pw_staged_len<M>(theta_len)) * sizeof(typename M::CT);
}

// g_exp2_big on the engine's device: B3 2^(j/2048) rounded from long double, (j << 9) subtracted from the high word
int ensure_exp_table(Engine& e) {
    static bool done[64] = {};
    if (done[e.device & 63]) return 0;
    static double host_tab[kExpBigSize];
    for (int j = 0; j < kExpBigSize; ++j) {
        const long double ln2 = 0.693147180559945309417232121458176568L;
        const double v = (double) (ln2 * ln2 * ln2 / 6.0L * exp2l((long double) j / kExpBigSize));  // B3 2^(j/2048)
        unsigned long long bits;
        memcpy(&bits, &v, 8);
        bits -= (unsigned long long) ((unsigned) j << 9) << 32;
        memcpy(&host_tab[j], &bits, 8);
    }
    static double host_rep[kExpRepSize * kExpRepCopies];
    for (int j = 0; j < kExpRepSize; ++j) {  // the kernels' table: B3 2^(j/256), high word - (j << 12), 16 copies side by side
        const double v = (double) (kExp3B3 * exp2l((long double) j / kExpRepSize));
        unsigned long long bits;
        memcpy(&bits, &v, 8);
        bits -= (unsigned long long) ((unsigned) j << 12) << 32;
        for (int c = 0; c < kExpRepCopies; ++c) memcpy(&host_rep[j * kExpRepCopies + c], &bits, 8);
    }
    static double host_log[512];
    if (cudaMemcpyFromSymbol(host_log, kLogTab, sizeof host_log) != cudaSuccess) {
        set_error("download of the log table failed: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    for (int j = kLogTabFirstHalf; j < 256; ++j) host_log[2 * j] *= 2.0;  // see log_pos_fast_k
    static double host_mid[512];  // log_pos_acc's table: {1/c_j (doubled in the upper half), -log of exactly that number}
    for (int j = 0; j < 256; ++j) {
        const long double c = 1.0L + ((long double) j + 0.5L) / 256.0L;
        const double inv = (double) ((j >= kLogTabFirstHalf ? 2.0L : 1.0L) / c);
        host_mid[2 * j] = inv;
        host_mid[2 * j + 1] = (double) -logl((long double) inv);
    }
    if (cudaMemcpyToSymbol(g_log_adj, host_log, sizeof host_log) != cudaSuccess ||
        cudaMemcpyToSymbol(g_log_mid, host_mid, sizeof host_mid) != cudaSuccess ||
        cudaMemcpyToSymbol(g_exp2_big, host_tab, sizeof host_tab) != cudaSuccess ||
        cudaMemcpyToSymbol(g_exp2_rep, host_rep, sizeof host_rep) != cudaSuccess) {
        set_error("upload of the exp table failed: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    done[e.device & 63] = true;
    return 0;
}

template <class M, typename T, int NPL>
int launch_one(Engine& e, const PointwiseArgs& proto, cudaStream_t s) {
    if constexpr (big_exp_table<M>())
        if (ensure_exp_table(e) != 0) return -1;
    void (*kern)(const PointwiseArgs);
    if constexpr (pw_two_ctas<M, NPL>())
        kern = pw_frontier_kernel2<M, T, NPL>;
    else
        kern = pw_frontier_kernel1<M, T, NPL>;
    static bool configured_dev[64] = {};
    static int occ_dev[64] = {};
    bool& configured = configured_dev[e.device & 63];
    int& occ = occ_dev[e.device & 63];
    // theta staging area sized for a full pass (32 * NPL nodes) so the occupancy is B-independent
    const size_t kPwSmem = pw_smem_bytes<M, NPL>(e.theta_len, 32 * NPL);
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kPwSmem);
        if (err == cudaSuccess)
            err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) {
            set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(err));
            return -1;
        }
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PW_THREADS, kPwSmem);
        if (occ < 1) occ = 1;
        if (occ > PF_PW_MAXOCC) occ = PF_PW_MAXOCC;
        configured = true;
    }
    PointwiseArgs a = proto;
    const unsigned long long rows = a.row_end - a.row_begin;
    constexpr unsigned TILE_ROWS = PW_TILE_BYTES / (M::DIM * sizeof(T));
    unsigned long long grid = (unsigned long long) e.sm_count * occ;
    const unsigned long long tiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
    if (grid > tiles) grid = tiles;
    if (grid > (unsigned long long) kMaxGrid) grid = kMaxGrid;
    if (grid < 1) grid = 1;
    a.rows_per_cta = (unsigned) ((rows + grid - 1) / grid);
    grid = (rows + a.rows_per_cta - 1) / a.rows_per_cta;
    if (grid < 1) grid = 1;
    kern<<<(unsigned) grid, PW_THREADS, kPwSmem, s>>>(a);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("pointwise kernel launch: %s", cudaGetErrorString(err));
        return -1;
    }
    return 1;
}

template <class M, typename T>
int launch_npl(Engine& e, const PointwiseArgs& a, cudaStream_t s, bool allow2) {
    if (a.B > 32 && allow2) return launch_one<M, T, 2>(e, a, s);
    return launch_one<M, T, 1>(e, a, s);
}

template <typename T>
int dispatch_normal(Engine& e, const PointwiseArgs& a, cudaStream_t s, bool general) {
#define PF_NORMAL_CASE(Dv)                                                                 \
    case Dv:                                                                               \
        return general ? launch_npl<NormalModel<Dv, T, true>, T>(e, a, s, Dv <= 2)         \
                       : launch_npl<NormalModel<Dv, T, false>, T>(e, a, s, Dv <= 4);
    switch (e.dim) {
        PF_NORMAL_CASE(1)
        PF_NORMAL_CASE(2)
        PF_NORMAL_CASE(3)
        PF_NORMAL_CASE(4)
        PF_NORMAL_CASE(8)
    }
#undef PF_NORMAL_CASE
    set_error("normal target: dimension %u not built (1,2,3,4,8)", e.dim);
    return -1;
}

template <typename T>
int dispatch_mixture(Engine& e, const PointwiseArgs& a, cudaStream_t s) {
    if (e.K == 8 && e.dim == 2) return launch_npl<MixtureModel<8, 2, T>, T>(e, a, s, true);
    if (e.K == 8 && e.dim == 8) return launch_npl<MixtureModel<8, 8, T>, T>(e, a, s, false);
    switch (e.dim) {
        case 1: return launch_npl<MixtureModelRT<1, T>, T>(e, a, s, true);
        case 2: return launch_npl<MixtureModelRT<2, T>, T>(e, a, s, true);
        case 3: return launch_npl<MixtureModelRT<3, T>, T>(e, a, s, true);
        case 4: return launch_npl<MixtureModelRT<4, T>, T>(e, a, s, true);
        case 8: return launch_npl<MixtureModelRT<8, T>, T>(e, a, s, true);
    }
    set_error("mixture: dimension %u not built (1,2,3,4,8)", e.dim);
    return -1;
}

}  // namespace

int ensure_exp_table_for(Engine& e) { return ensure_exp_table(e); }


bool pointwise_supported(const Engine& e) {
    const uint32_t d = e.dim;
    return d == 1 || d == 2 || d == 3 || d == 4 || d == 8;
}

uint32_t pointwise_max_frontier(const Engine& e) {
    if (e.desc.model == PF_MODEL_NORMAL) return e.dim <= 2 ? 64 : 32;  // general-covariance variant bound
    if (e.desc.model == PF_MODEL_MIXTURE) return (e.K == 8 && e.dim == 8) ? 32 : 64;
    return 64;
}

int launch_pointwise(Engine& e, uint32_t B, const double* d_theta, double* d_out, uint64_t first_item,
                     uint64_t n_items, cudaStream_t s) {
    PointwiseArgs a{};
    a.data = e.d_data;
    a.theta = d_theta;
    a.trace = e.d_trace;
    a.theta_inline_n = 0;
    if (e.inline_theta && e.inline_theta_n <= (unsigned) kInlineThetaDoubles) {
        a.theta_inline_n = e.inline_theta_n;
        memcpy(a.theta_inline, e.inline_theta, (size_t) e.inline_theta_n * sizeof(double));
    }
    a.part = e.d_part;
    a.out = d_out;
    a.ticket = e.d_ticket;
    a.item_rows = e.item_rows;
    a.row_begin = first_item * e.item_rows;
    a.row_end = (first_item + n_items) * e.item_rows;
    if (a.row_end > e.n_rows) a.row_end = e.n_rows;
    a.B = B;
    a.theta_len = e.theta_len;
    a.K = e.K;
    a.done_flag = e.done_flag;
    a.done_seq = e.done_seq;
    for (int j = 0; j < 8; ++j) a.xmax[j] = e.xmax[j];
    if (a.row_end <= a.row_begin) {  // no rows here: zeros (+ the same exchanges the other ranks' launches make)
        if (e.desc.model == PF_MODEL_MIXTURE && !e.f32 && B > 32 && e.gather_per == 0) {
            const int r1 = launch_exchange_only(e, 32, d_out, s, 0, B);
            if (r1 < 0) return r1;
            const int r2 = launch_exchange_only(e, B - 32, d_out, s, 32, B);
            return r2 < 0 ? r2 : r1 + r2;
        }
        return launch_exchange_only(e, B, d_out, s);
    }
    a.xchg = e.next_xchg();
    if (e.desc.model == PF_MODEL_NORMAL) {
        // The general-covariance variant reproduces the identity result exactly (W = I), so it is
        // always correct; the identity fast path is chosen by the caller when every node's
        // covariance is the identity (the only case the reference's proposals produce,
        // normalsampler.hh:73-75).
        const bool general = e.normal_general;
        return e.f32 ? dispatch_normal<float>(e, a, s, general) : dispatch_normal<double>(e, a, s, general);
    }
    if (!e.f32 && B > 32 && a.xchg.gather_per == 0) {
        // FP64 mixture, more than 32 nodes: two launches of <= 32 nodes (one node per lane) beat one launch with two
        // nodes per lane, which runs out of registers (N = 1e8, B = 64: 45.5 vs 52.7 ms); the data set is read twice,
        // which an FP64-bound kernel does not notice.  Each launch makes its own peer exchange.
        PointwiseArgs lo = a, hi = a;
        lo.B = 32;
        lo.out_off = 0;
        lo.out_stride = B;
        lo.theta_inline_n = 0;
        hi.B = B - 32;
        hi.theta = d_theta + (size_t) 32 * e.theta_len;
        hi.out_off = 32;
        hi.out_stride = B;
        hi.theta_inline_n = 0;
        if (a.theta_inline_n) {  // (tiny thetas only) the inline copy holds all B nodes: re-slice it
            lo.theta_inline_n = 32 * e.theta_len;
            hi.theta_inline_n = (B - 32) * e.theta_len;
            memcpy(hi.theta_inline, a.theta_inline + (size_t) 32 * e.theta_len, (size_t) hi.theta_inline_n * sizeof(double));
        }
        const int r1 = dispatch_mixture<double>(e, lo, s);
        if (r1 < 0) return r1;
        hi.xchg = e.next_xchg();
        const int r2 = dispatch_mixture<double>(e, hi, s);
        return r2 < 0 ? r2 : r1 + r2;
    }
    return e.f32 ? dispatch_mixture<float>(e, a, s) : dispatch_mixture<double>(e, a, s);
}


namespace {
template <int D, typename T>
int launch_ordered_t(Engine& e, uint32_t B, const double* d_theta, double* d_out, const unsigned long long* d_order,
                     uint64_t first, uint64_t count, bool general, cudaStream_t s) {
    unsigned nblocks = (unsigned) ((count + 2047) / 2048);
    if (nblocks > 128) nblocks = 128;
    if (nblocks < 1) nblocks = 1;
    const dim3 grid(nblocks, B);
    if (general)
        normal_ordered_kernel<D, T, true><<<grid, 256, 0, s>>>(static_cast<const T*>(e.d_data), e.n_rows, d_theta,
                                                             e.theta_len, d_order, first, count, e.d_part);
    else
        normal_ordered_kernel<D, T, false><<<grid, 256, 0, s>>>(static_cast<const T*>(e.d_data), e.n_rows, d_theta,
                                                              e.theta_len, d_order, first, count, e.d_part);
    ordered_finalize_kernel<<<1, kMaxNodesPerPass, 0, s>>>(e.d_part, nblocks, B, d_out);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("ordered range kernels: %s", cudaGetErrorString(err));
        return -1;
    }
    return 2;
}
}  // namespace

// d_order: device [B][2] = (stride, start) per node
int launch_normal_ordered(Engine& e, uint32_t B, const double* d_theta, double* d_out,
                          const unsigned long long* d_order, uint64_t first_item, uint64_t n_items, cudaStream_t s) {
    if (n_items == 0) {
        cudaMemsetAsync(d_out, 0, sizeof(double) * 2 * B, s);
        return 0;
    }
    const bool general = e.normal_general;
#define PF_ORD_CASE(Dv)                                                                                          \
    case Dv:                                                                                                     \
        return e.f32 ? launch_ordered_t<Dv, float>(e, B, d_theta, d_out, d_order, first_item, n_items, general, s) \
                     : launch_ordered_t<Dv, double>(e, B, d_theta, d_out, d_order, first_item, n_items, general, s);
    switch (e.dim) {
        PF_ORD_CASE(1)
        PF_ORD_CASE(2)
        PF_ORD_CASE(3)
        PF_ORD_CASE(4)
        PF_ORD_CASE(8)
    }
#undef PF_ORD_CASE
    set_error("normal target: dimension %u not built", e.dim);
    return -1;
}

// An empty local range on a rank of a peer-exchange group: one CTA that contributes zeros to the exchange (so the
// peers' kernels, which wait for this rank's record, and the call numbering stay in step) and delivers the totals
// exactly like a frontier kernel's finalizing CTA.  Without peers the result is simply zero.
__global__ void __launch_bounds__(256) exchange_only_kernel(const PeerXchg x, unsigned B, double* out,
                                                            unsigned long long* done_flag,
                                                            unsigned long long done_seq, unsigned out_off,
                                                            unsigned out_stride) {
    __shared__ double scratch[2 * kMaxNodesPerPass];
    const unsigned tid = threadIdx.x;
    double S = 0.0, Q = 0.0;
    unsigned Bout = B;
    if (x.nranks > 1) {
        const double2 t = peer_allreduce(x, tid, B, 256, 1, scratch, S, Q);
        S = t.x;
        Q = t.y;
        if (x.gather_per) Bout = x.gather_total;
    }
    if (tid < Bout) {
        const unsigned iS = out_off + tid, iQ = (out_stride ? out_stride : Bout) + out_off + tid;
        if (done_flag) {
            store_result_words(out, iS, S, done_seq);
            store_result_words(out, iQ, Q, done_seq);
        } else {
            out[iS] = S;
            out[iQ] = Q;
        }
    }
}

int launch_exchange_only(Engine& e, uint32_t B, double* d_out, cudaStream_t s, unsigned out_off, unsigned out_stride) {
    const PeerXchg x = e.next_xchg();
    exchange_only_kernel<<<1, 256, 0, s>>>(x, B, d_out, e.done_flag, e.done_seq, out_off, out_stride);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("exchange_only_kernel launch: %s", cudaGetErrorString(err));
        return -1;
    }
    return 1;
}

int launch_boltzmann_scalar(Engine& e, uint32_t B, const double* d_theta, double* d_out, uint64_t n_items,
                            cudaStream_t s) {
    boltzmann_scalar_kernel<<<(B + 63) / 64, 64, 0, s>>>(d_theta, d_out, B, (double) n_items, 1.0 / e.temperature);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("boltzmann_scalar_kernel launch: %s", cudaGetErrorString(err));
        return -1;
    }
    return 1;
}

}  // namespace pf

extern "C" PF_API int pf_debug_math(int which, const double* in, double* out, size_t n, int device) {
    if (!in || !out || cudaSetDevice(device) != cudaSuccess) return PF_ERR_INVALID;
    double *d_in = nullptr, *d_out = nullptr;
    if (cudaMalloc(&d_in, n * 8) != cudaSuccess || cudaMalloc(&d_out, n * 8) != cudaSuccess) return PF_ERR_NOMEM;
    cudaMemcpy(d_in, in, n * 8, cudaMemcpyHostToDevice);
    if (which >= 5 && which <= 9) {
        pf::Engine tmp;
        tmp.device = device;
        if (pf::ensure_exp_table_for(tmp) != 0) return PF_ERR_CUDA;
        pf::debug_math_kernel_big<<<(unsigned) ((n + 255) / 256), 256>>>(which, d_in, d_out, n);
    } else
        pf::debug_math_kernel<<<(unsigned) ((n + 255) / 256), 256>>>(which, d_in, d_out, n);
    cudaError_t err = cudaMemcpy(out, d_out, n * 8, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    return err == cudaSuccess ? PF_OK : PF_ERR_CUDA;
}
