// pf_pointwise.cuh -- model policies and the tile-scoring routine (pw_main) shared by the launched frontier kernels
// (pf_pointwise.cu) and the resident frontier kernel (pf_resident.cu).  See pf_pointwise.cu for the design notes.
#pragma once
#include <math.h>
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
//   rows per iteration 2, A_k re-read, single chain                        5.24 / 0.927 / 19.7
//   rows per iteration 3                                                   5.21 / 0.941 / 19.6
//   rows per iteration 1                                                   6.67 / 0.989 / 26.0
//   A_k and B_k both re-read (rows 2)                                      5.24 / 0.927 / 19.7
// Second half of round 2 (another box: 5.38 for the two-row default above): the rows of an iteration walk the components
// TOGETHER and share the A_k / B_k reads (PF_MIX_PAIR), and the per-row log is replaced by a running product (PF_MIX_PROD):
//   two rows together, per-row logs (PF_MIX_PROD = 0)                      5.24 / 0.908
//   per-row walk, running product (PF_MIX_PAIR = 0)                        5.12 / 0.953
//   two rows together, running product                                     4.74 / 0.869 / 18.0
//   four rows together (2 components x 4 rows), uniform masked tail step   4.67 / 0.861 / 17.7   (the default)
#ifndef PF_MIX_RB
#define PF_MIX_RB 4       // rows per iteration of the factored mixture loop (2 or 4 walk together; the clamped form keeps 2)
#endif
#ifndef PF_MIX_RELOAD
#define PF_MIX_RELOAD 1   // factored form: 0 = theta-derived constants wherever the compiler puts them (registers),
#endif                    // 1 = A_k re-read from shared memory for every row, 2 = A_k and B_k
#ifndef PF_MIX_TWO_ACC
#define PF_MIX_TWO_ACC 0  // factored mixture form: two partial sums per row (1) or a single accumulation chain (0)
#endif
// PF_MIX_PROD (factored form): no log per row -- the row sums are MULTIPLIED into a running product whose binary exponent is
// moved to an integer after every multiplication (one DMUL + SHF + IADD3 + LOP3 per row instead of 7 FP64 + 8 other for a
// table-driven log), and the log of the product is taken once per tile and lane.  Every row sum is a normal number in
// [2^-kFactoredLimit, K 2^kFactoredLimit] (kFactoredLimit = 500), so neither the product of two row sums nor its product with a mantissa in [1, 2) can leave
// the normal range.  Error: half an ulp per multiplication, i.e. <= 32 * 1.1e-16 absolute on a lane's tile sum
// (the per-row logs carried 5.7e-15 of truncation EACH).
#ifndef PF_MIX_PROD
#define PF_MIX_PROD 1
#endif
// PF_MIX_PAIR (factored form, two rows per iteration): the two rows walk the components TOGETHER -- A_k / B_k are read once
// for both, consecutive DFMAs share two of their three register operands (operand-reuse cache: a DFMA with three fresh
// register pairs costs 2.8 issue cycles instead of 2, scripts/fp64_probe.py), 4 components x 2 rows in flight.
#ifndef PF_MIX_PAIR
#define PF_MIX_PAIR 1
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
    // rows per loop iteration (factored: PF_MIX_RB; clamped FP64 form: 2; FP32: MUFU-bound, 4 is best)
    static constexpr unsigned RB = (FACT && BIG_EXP_TABLE) ? PF_MIX_RB : (BIG_EXP_TABLE ? 2 : 4);
    using Factored = MixtureModel<K, D, T, true>;           // the variant pw_frontier_body may switch to
    // doubles staged per node: theta as it is, or [K][D] B then [K] A
    static __host__ __device__ constexpr unsigned staged_len(unsigned theta_len) {
        return FACTORED ? theta_len + K : theta_len;
    }
    static constexpr bool PROD = FACTORED && (PF_MIX_PROD != 0);      // running product instead of a log per row
    static constexpr bool PAIR = FACTORED && (PF_MIX_PAIR != 0) && (RB == 2 || RB == 4) && (K % 2 == 0) && (D % 2 == 0);
    static constexpr bool QUAD = PAIR && RB == 4;
    struct RowState {
        unsigned redo;  // slow_path_accumulate() key, tested once per tile
        int nacc;       // sum of the rows' binary exponents (log_pos_acc / product form)
        double xsq;     // sum of the rows' |x|^2 (factored form)
        double prod;    // product form: mantissa in [1, 2) of the product of the row sums seen so far
        __device__ __forceinline__ void reset() {
            redo = 0u;
            nacc = 0;
            xsq = 0.0;
            prod = 1.0;
        }
    };
    // prod <- prod * v, exponent moved into nacc (v a positive normal number; prod stays in [1, 2))
    static __device__ __forceinline__ void prod_accumulate(RowState& st, double v) {
        const double p = st.prod * v;
        const int hi = __double2hiint(p);
        st.nacc += (hi >> 20) - 1023;  // SHF + IADD3
        st.prod = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(p));  // one LOP3
    }
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
    // two rows at once (PAIR): row sums into lpa / lpb
    static __device__ __forceinline__ void mix_sum2(const Node& nd, const CT (&xa)[D], const CT (&xb)[D], double& lpa, double& lpb) {
        constexpr int G2 = (K % 4 == 0) ? 4 : 2;  // components per group (K is even)
        lpa = 0.0;
        lpb = 0.0;
        const unsigned th_s = smem_u32(nd.th);
#pragma unroll
        for (int k0 = 0; k0 < K; k0 += G2) {
            double z[2 * G2], a[G2];
#pragma unroll
            for (int g = 0; g < G2; g += 2)
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a[g]), "=d"(a[g + 1]) : "r"(th_s + (unsigned) ((K * D + k0 + g) * 8)));
#pragma unroll
            for (int g = 0; g < G2; ++g) {
                double b[D];
#if PF_MIX_RELOAD >= 2
#pragma unroll
                for (int j = 0; j < D; j += 2)
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(b[j]), "=d"(b[j + 1]) : "r"(th_s + (unsigned) (((k0 + g) * D + j) * 8)));
#else
                load_row<double, D, double>(reinterpret_cast<const unsigned char*>(nd.th + (k0 + g) * D), b);
#endif
                z[2 * g] = fma(b[0], xa[0], a[g]);
                z[2 * g + 1] = fma(b[0], xb[0], a[g]);
#pragma unroll
                for (int j = 1; j < D; ++j) {
                    z[2 * g] = fma(b[j], xa[j], z[2 * g]);
                    z[2 * g + 1] = fma(b[j], xb[j], z[2 * g + 1]);
                }
            }
            mix_exp_accumulate_z<2 * G2, true>(z, lpa, lpb, nd.tab_s);  // even chains -> lpa, odd -> lpb
        }
    }
    // four rows at once (RB = 4): 2 components x 4 rows in flight; every A_k / B_k operand serves four consecutive DFMAs
    // MASKED (the uniform tail of a tile): rows whose bit in `valid` is clear arrive as x = 0 and count as a row sum of 1
    template <bool MASKED = false>
    static __device__ __forceinline__ double quad_fast(const Node& nd, const CT (&x)[4][D], RowState& st, unsigned valid = 15u) {
        constexpr int R = 4, GC = 2;
        double lp[R] = {0.0, 0.0, 0.0, 0.0};
        const unsigned th_s = smem_u32(nd.th);
#pragma unroll
        for (int k0 = 0; k0 < K; k0 += GC) {
            double z[R * GC], a[GC];
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a[0]), "=d"(a[1]) : "r"(th_s + (unsigned) ((K * D + k0) * 8)));
#pragma unroll
            for (int g = 0; g < GC; ++g) {
                double b[D];
#pragma unroll
                for (int j = 0; j < D; j += 2)
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(b[j]), "=d"(b[j + 1]) : "r"(th_s + (unsigned) (((k0 + g) * D + j) * 8)));
#pragma unroll
                for (int r = 0; r < R; ++r) z[R * g + r] = fma(b[0], x[r][0], a[g]);
#pragma unroll
                for (int j = 1; j < D; ++j)
#pragma unroll
                    for (int r = 0; r < R; ++r) z[R * g + r] = fma(b[j], x[r][j], z[R * g + r]);
            }
            mix_exp_accumulate_zr<R * GC, R>(z, lp, nd.tab_s);
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < D; ++j) st.xsq = fma(x[r][j], x[r][j], st.xsq);
        if constexpr (MASKED) {
#pragma unroll
            for (int r = 0; r < R; ++r) lp[r] = ((valid >> r) & 1u) ? lp[r] : 1.0;
        }
        if constexpr (PROD) {
            prod_accumulate(st, lp[0] * lp[1]);
            prod_accumulate(st, lp[2] * lp[3]);
            return 0.0;
        } else {
            double v = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) v += log_pos_acc(lp[r], nd.ltab_s, st.nacc);
            return v;
        }
    }
    // PAIR hot loop body: both rows' contributions go into the lane state (product form) or come back summed
    template <bool MASKED = false>
    static __device__ __forceinline__ double pair_fast(const Node& nd, const CT (&xa)[D], const CT (&xb)[D], RowState& st, unsigned valid = 3u) {
        double lpa, lpb;
        mix_sum2(nd, xa, xb, lpa, lpb);
        if constexpr (MASKED) {
            lpa = (valid & 1u) ? lpa : 1.0;
            lpb = (valid & 2u) ? lpb : 1.0;
        }
#pragma unroll
        for (int j = 0; j < D; ++j) st.xsq = fma(xa[j], xa[j], st.xsq);
#pragma unroll
        for (int j = 0; j < D; ++j) st.xsq = fma(xb[j], xb[j], st.xsq);
        if constexpr (PROD) {
            prod_accumulate(st, lpa * lpb);  // >= 2^-2 kFactoredLimit: normal
            return 0.0;
        } else
            return log_pos_acc(lpa, nd.ltab_s, st.nacc) + log_pos_acc(lpb, nd.ltab_s, st.nacc);
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
            if constexpr (PROD) {
                prod_accumulate(st, lp);
                return 0.0;
            } else
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
    static __device__ __forceinline__ void tile_fixup(const Node& nd, const RowState& st, double& acc) {
        if constexpr (PROD) {  // acc carries nothing yet: the tile's rows live in (prod, nacc, xsq)
            int n = st.nacc;
            const double lg = log_pos_acc(st.prod, nd.ltab_s, n);
            acc = log_n_times_ln2(n, fma(-0.5, st.xsq, lg));
        } else if constexpr (FACTORED)
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
template <class M, class = void>
struct ProdRows { static constexpr bool value = false; };
template <class M>
struct ProdRows<M, std::enable_if_t<M::PROD>> { static constexpr bool value = true; };
template <class M>
__host__ __device__ constexpr bool prod_rows() { return ProdRows<M>::value; }
template <class M, class = void>
struct PairRows { static constexpr bool value = false; };
template <class M>
struct PairRows<M, std::enable_if_t<M::PAIR>> { static constexpr bool value = true; };
template <class M>
__host__ __device__ constexpr bool pair_rows() { return PairRows<M>::value; }

template <class M, bool MASKED = false, class Node, class X, class St>
__device__ __forceinline__ double pair_call(const Node& nd, const X& xa, const X& xb, St& st, unsigned valid = 3u) {
    if constexpr (pair_rows<M>())
        return M::template pair_fast<MASKED>(nd, xa, xb, st, valid);
    else
        return 0.0;
}
template <class M, class = void>
struct QuadRows { static constexpr bool value = false; };
template <class M>
struct QuadRows<M, std::enable_if_t<M::QUAD>> { static constexpr bool value = true; };

template <class M, bool MASKED = false, class Node, class X, class St>
__device__ __forceinline__ double quad_call(const Node& nd, const X& x, St& st, unsigned valid = 15u) {
    if constexpr (pair_rows<M>())
        return M::template quad_fast<MASKED>(nd, x, st, valid);
    else
        return 0.0;
}

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

// One call of the resident kernel (pf_resident.cu) as pw_main sees it: the same fields as PointwiseArgs where the
// two share meaning, living in registers instead of the kernel parameters.
struct ResidentCall {
    const double* theta;          // raw thetas of this call, staged in SHARED memory ([B][theta_len]; normal target)
    double* part;
    double* out;                  // mapped host memory: self-validating result words
    unsigned long long* ticket;   // device, MONOTONIC over the calls of one launch (never reset)
    unsigned long long ticket_target;  // value of the ticket after the last CTA of this call arrived
    unsigned long long row_begin, row_end, item_rows;
    unsigned int rows_per_cta, B, theta_len, K;
    unsigned long long* done_flag;
    unsigned long long done_seq;
    PeerXchg xchg;
    unsigned int out_off, out_stride;
    unsigned long long* trace;
    const unsigned char* rows;    // this CTA's rows, resident in shared memory (16-byte aligned window like a stage)
    unsigned int tile_base;       // tiles this CTA scored in earlier calls (phases of the lead-limiting barriers)
};

// RES = false: rows arrive through the TMA ring (launched kernels, A = PointwiseArgs);
// RES = true: rows are resident in shared memory (A = ResidentCall).
template <class M, typename T, int NPL, bool RES = false, class A = PointwiseArgs>
__device__ __forceinline__ void pw_main(const A& a, const PwShared& sh);

// fields only a ResidentCall has (the launched kernels never evaluate these)
template <class A> __device__ __forceinline__ unsigned sh_tile_base(const A&) { return 0u; }
template <> __device__ __forceinline__ unsigned sh_tile_base<ResidentCall>(const ResidentCall& a) { return a.tile_base; }
template <class A> __device__ __forceinline__ const unsigned char* sh_rows(const A&) { return nullptr; }
template <> __device__ __forceinline__ const unsigned char* sh_rows<ResidentCall>(const ResidentCall& a) { return a.rows; }
template <class A> __device__ __forceinline__ unsigned long long sh_ticket_target(const A&) { return 0ull; }
template <> __device__ __forceinline__ unsigned long long sh_ticket_target<ResidentCall>(const ResidentCall& a) { return a.ticket_target; }

template <class M, typename T, int NPL, bool RES, class A>
__device__ __forceinline__ void pw_main(const A& a, const PwShared& sh) {
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
    [[maybe_unused]] auto issue_tile = [&](unsigned long long t0, unsigned long long t1, unsigned s) {
        if constexpr (!RES) {
            const unsigned long long b0 = t0 * RB_BYTES, b1 = t1 * RB_BYTES;
            const unsigned long long w0 = b0 & ~15ull, w1 = (b1 + 15ull) & ~15ull;
            const unsigned bytes = (unsigned) (w1 - w0);
            mbar_arrive_expect_tx(&full[s], bytes);
            bulk_g2s(stage_base + s * PW_STAGE_BYTES, static_cast<const unsigned char*>(a.data) + w0, bytes, &full[s]);
        }
    };

    // ---------------------------------------------------------------------- consumer warps
    const unsigned B = a.B;
    // lane = point-lane * NB + node: NB = min(B, 32) node lanes, PL = 32 / NB point lanes; when NB does not divide 32 the
    // last 32 - PL * NB lanes are spares (they re-walk the rows of point lane 0 and their sums are dropped).  B = 3 / 5 / 6
    // keep 30 of 32 lanes busy (a power-of-two layout: 24 / 20 / 24), B = 9 / 10: 27 / 30 (18 / 20).
    const unsigned NB = B < 32 ? B : 32;
    const unsigned PL = 32 / NB;
    const bool nb_pow2 = (NB & (NB - 1)) == 0;
    const unsigned nb = lane % NB, pl = lane / NB;
    const unsigned slot = warp * PL + (pl < PL ? pl : 0u), nslots = PW_CONSUMER_WARPS * PL;
    // sum over the point lanes of a node (valid in the lanes pl == 0): butterfly, or a tree of strided shuffles
    auto point_lane_sum = [&](double v) {
        if (nb_pow2) {
#pragma unroll
            for (unsigned m = 16; m >= 1; m >>= 1)  // (unrolled: the run-time loop cost 10 instructions per step -- at B = 1,
                if (m >= NB) v += shfl_xor_f64(v, m);  //  five steps per four rows of a lane; NB is warp-uniform)
        } else {
            for (unsigned off = 1; off < PL; off <<= 1) {
                const double o = shfl_down_f64(v, (int) (off * NB));
                if (pl + off < PL) v += o;
            }
        }
        return v;
    };

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
        const double* theta;
        if constexpr (RES)
            theta = a.theta;
        else
            theta = a.theta_inline_n ? a.theta_inline : a.theta;
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
    const bool issuer = PW_FOLD && !RES && tid == 0;
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
        // resident rows: tiles are numbered through all the calls of the launch (tt) -- the per-tile buffers and barriers
        // keep rotating, and the `empty` barriers only bound a warp's lead over the slowest one to PW_STAGES - 1 tiles
        // (what the ring does for the launched kernels: a tile-total buffer is never rewritten before its owners read it)
        const unsigned tt = RES ? sh_tile_base(a) + t : t;
        const unsigned s = tt % PW_STAGES;
        bool issue_pending = false;
        const unsigned char* base;
        if constexpr (RES) {
            if constexpr (!M::ITEM1)
                if (tt >= PW_STAGES) mbar_wait(&empty[s], ((tt / PW_STAGES) - 1) & 1);
            base = sh_rows(a) + (size_t) (t0 - r0) * RB_BYTES;
        } else {
            if (issuer && t >= 1 && pu == t + PW_STAGES - 1 && pit.cur < pit.end_all) issue_pending = !try_issue(t, false);
            mbar_wait(&full[s], (t / PW_STAGES) & 1);
            if (t == 0) trace_mark(a.trace, 0, 2, tr0);  // first tile landed
            base = stage_base + s * PW_STAGE_BYTES + (unsigned) ((t0 * RB_BYTES) & 15ull);
        }
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
                if constexpr (RB == 4 && pair_rows<M>()) {
#pragma unroll
                    for (int n = 0; n < NPL; ++n) {
                        const double v = quad_call<M>(nd[n], x, st[n]);
                        if constexpr (!prod_rows<M>()) accS[n] += v;
                    }
                } else if constexpr (RB == 2 && pair_rows<M>()) {
#pragma unroll
                    for (int n = 0; n < NPL; ++n) {
                        const double v = pair_call<M>(nd[n], x[0], x[1], st[n]);
                        if constexpr (!prod_rows<M>()) accS[n] += v;
                    }
                } else {
#pragma unroll
                    for (unsigned u = 0; u < RB; ++u)
#pragma unroll
                        for (int n = 0; n < NPL; ++n) {
                            double v;
                            if constexpr (M::HAS_REDO || M::HAS_FIXUP)
                                v = (double) M::rowlp_fast(nd[n], x[u], st[n]);
                            else
                                v = (double) M::rowlp(nd[n], x[u]);
                            if constexpr (!prod_rows<M>()) accS[n] += v;
                            if constexpr (M::ITEM1) accQ[n] = fma(v, v, accQ[n]);
                        }
                }
            }
        };
        if constexpr (pair_rows<M>() && NPL == 1) {
            // UNIFORM row steps: every lane of the CTA runs the same number of iterations.  Full steps of RBm rows per
            // slot first; what is left (< RBm * nslots rows) is one masked step in which the rows past the end arrive as
            // x = 0 and count as a row sum of 1.  (With per-lane trip counts the warp holding the last slots ran an extra
            // loop body per tile -- at B = 1, where a tile is four rows per lane, that warp bounded the CTA at +50 %.)
            constexpr unsigned RBm = QuadRows<M>::value ? 4u : 2u;
            unsigned left = nrows;
            for (; left >= RBm * nslots; left -= RBm * nslots, j += RBm * nslots) {
                CT x[RBm][DIM];
#pragma unroll
                for (unsigned u = 0; u < RBm; ++u) load_row<T, DIM, CT>(base + (size_t) (j + u * nslots) * RB_BYTES, x[u]);
                double v;
                if constexpr (RBm == 4)
                    v = quad_call<M>(nd[0], x, st[0]);
                else
                    v = pair_call<M>(nd[0], x[0], x[1], st[0]);
                if constexpr (!prod_rows<M>()) accS[0] += v;
            }
            auto masked_step = [&](auto rb_tag) {
                constexpr unsigned RBt = decltype(rb_tag)::value;
                CT x[RBt][DIM];
                unsigned valid = 0u;
#pragma unroll
                for (unsigned u = 0; u < RBt; ++u) {
                    const unsigned row = j + u * nslots;
                    const bool ok = row < nrows;
                    valid |= ok ? (1u << u) : 0u;
                    load_row<T, DIM, CT>(base + (size_t) (ok ? row : 0u) * RB_BYTES, x[u]);
#pragma unroll
                    for (int d = 0; d < DIM; ++d) x[u][d] = ok ? x[u][d] : (CT) 0;
                }
                double v;
                if constexpr (RBt == 4)
                    v = quad_call<M, true>(nd[0], x, st[0], valid);
                else
                    v = pair_call<M, true>(nd[0], x[0], x[1], st[0], valid);
                if constexpr (!prod_rows<M>()) accS[0] += v;
                j += RBt * nslots;
            };
            if constexpr (RBm == 4) {
                if (left > 2 * nslots)
                    masked_step(std::integral_constant<unsigned, 4>{});
                else if (left > 0)
                    masked_step(std::integral_constant<unsigned, 2>{});
            } else if (left > 0)
                masked_step(std::integral_constant<unsigned, 2>{});
            j = nrows;  // nothing left for the per-lane loops below
        } else if constexpr (M::ITEM1)
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
                if constexpr (!prod_rows<M>()) accS[n] += v;
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
            for (int n = 0; n < NPL; ++n) M::tile_fixup(nd[n], st[n], accS[n]);
        }
        if constexpr (M::ITEM1) {
            if constexpr (!RES) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
        } else {
            // Tile total per node: lanes -> warp (shuffles) -> per-warp slot in one of 4 rotating
            // buffers.  Every warp ARRIVES on the tile's named barrier and moves on; only the owner
            // warp(s) -- the ones holding threads tid < B -- WAIT on it, then fold the per-warp slots.  The
            // smem stage is released after the arrival, which bounds any warp's lead to 3 tiles, so a
            // buffer is never rewritten before the owners have read it.
            constexpr int RS = 32 * NPL;  // slot row stride
            double* r = red + (tt & (PW_RED_BUFS - 1)) * (PW_CONSUMER_WARPS * RS);
#pragma unroll
            for (int n = 0; n < NPL; ++n) {
                const double v = point_lane_sum(accS[n]);
                accS[n] = 0.0;
                if (pl == 0) r[warp * RS + nb + 32 * n] = v;
            }
            const int bar = PW_BAR_TILE0 + (int) (tt & (PW_RED_BUFS - 1));
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
    double* rec = a.part + blockIdx.x;  // column of this CTA in the [slot][node][CTA] records
    if constexpr (M::ITEM1) {
        // one block reduction at the very end: every row is its own item
#pragma unroll
        for (int n = 0; n < NPL; ++n) {
            const double v = point_lane_sum(accS[n]), q = point_lane_sum(accQ[n]);
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
        if constexpr (RES)
            flag[0] = (atomicAdd(a.ticket, 1ull) + 1ull == sh_ticket_target(a));
        else
            flag[0] = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
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
        finalize_items<!M::ITEM1>(a.part, gridDim.x, a.row_begin, a.row_end, a.rows_per_cta, M::ITEM1 ? 1ull : a.item_rows, tid,
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
    if constexpr (!RES)
        if (tid == 0) *a.ticket = 0u;  // ready for the next launch on this stream
    trace_mark(a.trace, 1, 2, tid == 0);
}

}  // namespace pf
