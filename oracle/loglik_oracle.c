/* oracle/loglik_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, CPU restatement of the arithmetic on the frontier log-likelihood path of
 * elaine84/fetching.  It is the CHECKER for the CUDA engine: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it.  The product (fetching_b200/) never links, imports or executes anything in oracle/.
 *
 * PINNING.  Every function below is checked in tests/test_oracle.py against
 *   - oracle/_ref/pf_ref, the reference's own C++ compiled in place (normal target,
 *     scalar Boltzmann, StrideIndex order, RNG stream, dataset recipe), bit-exact; and
 *   - tests/golden/py_models.npz, produced by executing the reference's own
 *     pymixture.py / pyblasso.py method bodies under numpy (oracle/make_goldens.py),
 *     to 1e-12 relative (numpy's pairwise sum and SIMD exp differ from libm in the last
 *     bits; this file sums sequentially like worker.hh:148-155 does across items).
 * The dense Boltzmann energy (BASELINE config 2) has NO reference implementation
 * (boltzmannsampler.hh:71-74 is a scalar toy): orc_boltzmann_dense_* DEFINES the new
 * workload and its parity is "unpinned" by construction.
 *
 * Compiled with -ffp-contract=off: no fused multiply-adds, like the reference's
 * `-O2` x86-64 build (GNUmakefile:1,8).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_LOG_2PI /* normalsampler.hh:24-26 */
#define M_LOG_2PI 1.837877066409345483560659472810
#endif

/* ------------------------------------------------------------------------------------------
 * Mersenne twister 19937 (boost::random::mt19937, default seed 5489; randomsequence.hh:19,
 * master.hh:131, normal.cc:41).  Independent C statement of the published algorithm.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    uint32_t mt[624];
    int idx;
    uint64_t count; /* raw outputs produced so far == RandomSequence position */
} orc_mt;

void orc_mt_seed(orc_mt* g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t) i;
    g->idx = 624;
    g->count = 0;
}

uint32_t orc_mt_next(orc_mt* g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    ++g->count;
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* first `count` raw outputs of a seeded stream, for RandomSequence checks
 * (testmaster.cc:295-332 uses seed 18340). */
void orc_mt_stream(uint32_t seed, size_t count, uint32_t* out) {
    orc_mt g;
    orc_mt_seed(&g, seed);
    for (size_t i = 0; i != count; ++i)
        out[i] = orc_mt_next(&g);
}

/* boost::random::uniform_01<double> on a 32-bit engine: one draw times 2^-32. */
static double orc_uniform01(orc_mt* g) {
    for (;;) {
        double r = (double) orc_mt_next(g) * (1.0 / 4294967296.0);
        if (r < 1.0)
            return r;
    }
}

/* boost::random::normal_distribution<> (Box-Muller with cached second variate). */
typedef struct {
    double r1, rho;
    int valid;
} orc_normal_state;

static double orc_normal(orc_mt* g, orc_normal_state* st, double mean, double sigma) {
    if (!st->valid) {
        st->r1 = orc_uniform01(g);
        double r2 = orc_uniform01(g);
        st->rho = sqrt(-2.0 * log(1.0 - r2));
        st->valid = 1;
    } else
        st->valid = 0;
    const double pi = 3.14159265358979323846;
    return st->rho * (st->valid ? cos(2.0 * pi * st->r1) : sin(2.0 * pi * st->r1)) * sigma + mean;
}

/* boost::random::uniform_int_distribution<int>(0, range) on a 32-bit engine. */
static uint32_t orc_uniform_int(orc_mt* g, uint32_t range) {
    if (range == 0)
        return 0;
    if (range == 0xffffffffu)
        return orc_mt_next(g);
    uint32_t bucket = 0xffffffffu / (range + 1u);
    if (0xffffffffu % (range + 1u) == range)
        ++bucket;
    for (;;) {
        uint32_t r = orc_mt_next(g) / bucket;
        if (r <= range)
            return r;
    }
}

/* The synthetic dataset of the `normal` sampler: samplers/normal.cc:41-46 (identical to
 * testmaster.cc:487-492): one default-seeded mt19937, one long-lived normal_distribution,
 * value i = N(0,1) + (i odd ? 100 : 200). */
void orc_normal_dataset(size_t n_points, double* out /* [n_points][2] */) {
    orc_mt g;
    orc_mt_seed(&g, 5489u);
    orc_normal_state st = {0, 0, 0};
    for (size_t i = 0; i != n_points * 2; ++i)
        out[i] = orc_normal(&g, &st, 0.0, 1.0) + ((i % 2) ? 100.0 : 200.0);
}

/* ------------------------------------------------------------------------------------------
 * StrideIndex: strideindex.hh:46-57
 * ------------------------------------------------------------------------------------------ */
static size_t orc_gcd(size_t a, size_t b) {
    while (b) {
        size_t t = a % b;
        a = b;
        b = t;
    }
    return a;
}

size_t orc_valid_stride(size_t n, size_t stride) { /* strideindex.hh:50-57 */
    if (stride <= 1)
        return 1;
    size_t g;
    while ((g = orc_gcd(n, stride)) > 1)
        stride = (stride + g - 1) % n;
    return stride <= 1 ? 1 : stride;
}

size_t orc_stride_index(size_t n, size_t stride, size_t start, size_t i) { /* :46-48 */
    return (start + i * stride) % n;
}

/* NormalSampler<D>::prepare draws (normalsampler.hh:115-119): two uniform_int(0, n-1)
 * from the replay stream at `random_position` (raw-output index of a default-seeded
 * mt19937; randomsequence.hh:121-136).  Returns the stream position afterwards. */
size_t orc_normal_prepare_draws(size_t n, size_t random_position, size_t* stride, size_t* start) {
    orc_mt g;
    orc_mt_seed(&g, 5489u);
    for (size_t i = 0; i != random_position; ++i)
        (void) orc_mt_next(&g);
    uint32_t range = (uint32_t) (n - 1);
    uint32_t s = orc_uniform_int(&g, range);
    uint32_t o = orc_uniform_int(&g, range);
    *stride = orc_valid_stride(n, s);
    *start = o;
    return (size_t) g.count;
}

/* NormalSampler<D>::load for a need_proposal job (normalsampler.hh:70-76): a FRESH
 * normal_distribution(0, 0.1) adds one variate to each of the D means, consuming the
 * replay stream from `random_position`.  Returns the stream position afterwards. */
size_t orc_normal_propose(int D, size_t random_position, double* theta /* in/out, D + D(D+1)/2 */) {
    orc_mt g;
    orc_mt_seed(&g, 5489u);
    for (size_t i = 0; i != random_position; ++i)
        (void) orc_mt_next(&g);
    orc_normal_state st = {0, 0, 0};
    for (int i = 0; i != D; ++i)
        theta[i] = theta[i] + orc_normal(&g, &st, 0.0, 0.1);
    return (size_t) g.count;
}

/* ------------------------------------------------------------------------------------------
 * a1/a2  NormalSampler<D>::evaluate / evaluate_many   normalsampler.hh:87-165
 *
 * theta = D means then packed lower-triangular covariance (normaltheta.hh:24,49,81-87).
 * Identity covariance -> scale = -D*log(2pi)/2 (:88-90), dot = sum (mu_i - x_i)^2 (:132-137).
 * Otherwise LU-invert the covariance (:91-112) and dot = (mu-x)' Winv (mu-x) with the
 * UPPER triangle of Winv (:138-150); scale = -(D log 2pi + log det)/2.
 * Items are visited in StrideIndex order (:129); sum += lp, sum_sq += lp*lp (:160-164).
 * ------------------------------------------------------------------------------------------ */
static int orc_identity_cov(int D, const double* theta) { /* normaltheta.hh:116-125 */
    for (int i = 0, x = 0; i != D; ++i)
        for (int j = 0; j <= i; ++j, ++x)
            if (theta[D + x] != (j == i ? 1.0 : 0.0))
                return 0;
    return 1;
}

/* Winv (row-major DxD) and log-density constant for a general covariance. */
static void orc_normal_general_prepare(int D, const double* theta, double* winv, double* scale) {
    double A[16 * 16];
    int perm[16];
    for (int i = 0, idx = 0; i != D; ++i)
        for (int j = 0; j <= i; ++j, ++idx) {
            A[i * D + j] = theta[D + idx];
            A[j * D + i] = theta[D + idx];
        }
    int signum = 1;
    for (int i = 0; i != D; ++i)
        perm[i] = i;
    for (int j = 0; j + 1 < D; ++j) { /* gsl_linalg_LU_decomp: partial pivoting */
        double max = fabs(A[j * D + j]);
        int piv = j;
        for (int i = j + 1; i < D; ++i)
            if (fabs(A[i * D + j]) > max) {
                max = fabs(A[i * D + j]);
                piv = i;
            }
        if (piv != j) {
            for (int k = 0; k != D; ++k) {
                double t = A[j * D + k];
                A[j * D + k] = A[piv * D + k];
                A[piv * D + k] = t;
            }
            int t = perm[j];
            perm[j] = perm[piv];
            perm[piv] = t;
            signum = -signum;
        }
        double ajj = A[j * D + j];
        if (ajj != 0.0)
            for (int i = j + 1; i < D; ++i) {
                double aij = A[i * D + j] / ajj;
                A[i * D + j] = aij;
                for (int k = j + 1; k < D; ++k)
                    A[i * D + k] -= aij * A[j * D + k];
            }
    }
    for (int c = 0; c != D; ++c) { /* gsl_linalg_LU_invert */
        double x[16];
        for (int i = 0; i != D; ++i)
            x[i] = (perm[i] == c) ? 1.0 : 0.0;
        for (int i = 0; i != D; ++i)
            for (int k = 0; k < i; ++k)
                x[i] -= A[i * D + k] * x[k];
        for (int i = D - 1; i >= 0; --i) {
            for (int k = i + 1; k < D; ++k)
                x[i] -= A[i * D + k] * x[k];
            x[i] /= A[i * D + i];
        }
        for (int i = 0; i != D; ++i)
            winv[i * D + c] = x[i];
    }
    double det = (double) signum;
    for (int i = 0; i != D; ++i)
        det *= A[i * D + i];
    *scale = -(D * M_LOG_2PI + log(det)) / 2;
}

static double orc_normal_point(int D, const double* theta, int identity, const double* winv,
                               double scale, const double* x) {
    double dot = 0.0;
    if (identity) {
        for (int i = 0; i != D; ++i) {
            double t = theta[i] - x[i];
            dot += t * t;
        }
    } else {
        double w1[16], w2[16];
        for (int i = 0; i != D; ++i) {
            w1[i] = theta[i] - x[i];
            w2[i] = 0.0;
        }
        for (int i = 0; i != D; ++i) { /* dsymv, CblasUpper, alpha 1, beta 0 */
            double temp1 = 1.0 * w1[i], temp2 = 0.0;
            w2[i] += temp1 * winv[i * D + i];
            for (int j = i + 1; j < D; ++j) {
                w2[j] += temp1 * winv[i * D + j];
                temp2 += w1[j] * winv[i * D + j];
            }
            w2[i] += 1.0 * temp2;
        }
        for (int i = 0; i != D; ++i)
            dot += w1[i] * w2[i];
    }
    return -0.5 * dot + scale;
}

/* Accumulates INTO *sum / *sum_sq like evaluate_many (normalsampler.hh:157-165).
 * Items [first, first+count) in the order (start + i*stride) % n. */
void orc_normal_eval(const double* data, size_t n, int D, const double* theta, size_t stride,
                     size_t start, size_t first, size_t count, double* sum, double* sum_sq) {
    int identity = orc_identity_cov(D, theta);
    double winv[16 * 16], scale;
    if (identity)
        scale = -D * M_LOG_2PI / 2;
    else
        orc_normal_general_prepare(D, theta, winv, &scale);
    double s = *sum, q = *sum_sq;
    for (size_t p = first; p != first + count; ++p) {
        const double* x = &data[(size_t) D * ((start + p * stride) % n)];
        double lp = orc_normal_point(D, theta, identity, winv, scale, x);
        s += lp;
        q += lp * lp;
    }
    *sum = s;
    *sum_sq = q;
}

/* extended-precision companion (tolerance budgeting only; natural order) */
void orc_normal_eval_ld(const double* data, size_t n, int D, const double* theta, long double* sum,
                        long double* sum_sq) {
    int identity = orc_identity_cov(D, theta);
    double winv[16 * 16], scale;
    if (identity)
        scale = -D * M_LOG_2PI / 2;
    else
        orc_normal_general_prepare(D, theta, winv, &scale);
    long double s = 0, q = 0;
    for (size_t p = 0; p != n; ++p) {
        long double dot = 0;
        const double* x = &data[(size_t) D * p];
        if (identity)
            for (int i = 0; i != D; ++i) {
                long double t = (long double) theta[i] - x[i];
                dot += t * t;
            }
        else
            for (int i = 0; i != D; ++i)
                for (int j = 0; j != D; ++j) {
                    long double w = (j >= i) ? winv[i * D + j] : winv[j * D + i];
                    dot += ((long double) theta[i] - x[i]) * w * ((long double) theta[j] - x[j]);
                }
        long double lp = -0.5L * dot + scale;
        s += lp;
        q += lp * lp;
    }
    *sum = s;
    *sum_sq = q;
}

/* ------------------------------------------------------------------------------------------
 * a4  BoltzmannSampler::evaluate / evaluate_many   boltzmannsampler.hh:71-82  (scalar toy)
 * ------------------------------------------------------------------------------------------ */
void orc_boltzmann_scalar_eval(double theta, double temperature, size_t count, double* sum,
                               double* sum_sq) {
    double x = -theta / temperature;
    *sum += x * count;
    *sum_sq += x * x * count;
}

/* NEW WORKLOAD (BASELINE config 2; no reference code).  Dense pairwise energy
 *   E(x) = x' W x,   log p(x) = -E(x)/T.
 * The data set is ONE item, as the reference's Boltzmann plugin registers data_size = 1
 * (samplers/boltzmann.cc:36): metric_sum = -x'Wx/T and metric_sum_sq = metric_sum^2.
 * orc_boltzmann_dense_eval exposes the row terms e_i = -x_i (W_i . x)/T (and their squares)
 * as a building block; orc_boltzmann_dense_logp is the item value the engine must return. */
void orc_boltzmann_dense_eval(const double* W, size_t D, const double* x, double temperature,
                              size_t first, size_t count, double* sum, double* sum_sq) {
    double s = *sum, q = *sum_sq;
    for (size_t i = first; i != first + count; ++i) {
        double dot = 0.0;
        for (size_t j = 0; j != D; ++j)
            dot += W[i * D + j] * x[j];
        double e = -(x[i] * dot) / temperature;
        s += e;
        q += e * e;
    }
    *sum = s;
    *sum_sq = q;
}

void orc_boltzmann_dense_eval_ld(const double* W, size_t D, const double* x, double temperature,
                                 long double* sum, long double* sum_sq) {
    long double s = 0, q = 0;
    for (size_t i = 0; i != D; ++i) {
        long double dot = 0;
        for (size_t j = 0; j != D; ++j)
            dot += (long double) W[i * D + j] * x[j];
        long double e = -((long double) x[i] * dot) / temperature;
        s += e;
        q += e * e;
    }
    *sum = s;
    *sum_sq = q;
}

/* ------------------------------------------------------------------------------------------
 * a5  Mixture.evaluate   pysamplers/pymixture.py:86-92
 *
 *   lp = zeros(batch)
 *   for i in range(K): td = theta[i] - data[pos*b:(pos+1)*b]; lp += exp(-(td**2)/2.).prod(axis=1)
 *   return log(lp).sum()
 *
 * One ITEM = one batch of `item_rows` whole rows.  K*d exps
 * per row, multiplied left to right, components added in order; UNNORMALISED.
 * ------------------------------------------------------------------------------------------ */
double orc_mixture_item(const double* data, size_t n_rows, int d, int K, const double* theta,
                        size_t item_rows, size_t position) {
    size_t r0 = position * item_rows, r1 = (position + 1) * item_rows;
    if (r1 > n_rows)
        /* a short item: `lp = zeros(batch)` cannot broadcast against the clamped slice
         * (pymixture.py:88-91), the call raises, and python.cc:188-191 returns 0 */
        return 0.0;
    double total = 0.0;
    for (size_t r = r0; r != r1; ++r) {
        const double* x = &data[r * (size_t) d];
        double lp = 0.0;
        for (int k = 0; k != K; ++k) {
            double prod = 1.0;
            for (int j = 0; j != d; ++j) {
                double td = theta[k * d + j] - x[j];
                prod *= exp(-(td * td) / 2.);
            }
            lp += prod;
        }
        total += log(lp);
    }
    return total;
}

/* worker.hh:148-155 over items [first, first+count): sum += lp_item; sum_sq += lp_item^2 */
void orc_mixture_eval(const double* data, size_t n_rows, int d, int K, const double* theta,
                      size_t item_rows, size_t first, size_t count, double* sum, double* sum_sq) {
    double s = *sum, q = *sum_sq;
    for (size_t p = first; p != first + count; ++p) {
        double lp = orc_mixture_item(data, n_rows, d, K, theta, item_rows, p);
        s += lp;
        q += lp * lp;
    }
    *sum = s;
    *sum_sq = q;
}

void orc_mixture_eval_ld(const double* data, size_t n_rows, int d, int K, const double* theta,
                         size_t item_rows, size_t n_items, long double* sum, long double* sum_sq) {
    long double s = 0, q = 0;
    for (size_t p = 0; p != n_items; ++p) {
        size_t r0 = p * item_rows, r1 = (p + 1) * item_rows;
        long double item = 0;
        if (r1 > n_rows)
            r0 = r1 = 0; /* short item contributes 0, see orc_mixture_item */
        for (size_t r = r0; r != r1; ++r) {
            long double lp = 0;
            for (int k = 0; k != K; ++k) {
                long double e = 0;
                for (int j = 0; j != d; ++j) {
                    long double td = (long double) theta[k * d + j] - data[r * (size_t) d + j];
                    e += td * td;
                }
                lp += expl(-e / 2);
            }
            item += logl(lp);
        }
        s += item;
        q += item * item;
    }
    *sum = s;
    *sum_sq = q;
}

/* ------------------------------------------------------------------------------------------
 * a7  BLasso.evaluate   pysamplers/pyblasso.py:135-141
 *
 *   lp = (log(1/sigma/sqrt(2 pi)) - (dot(beta, X.T) - y)**2 / 2 / sigma**2).sum()
 *
 * theta = (sigma, beta[0..p-1]); one ITEM = `item_rows` consecutive rows (18 000 in the
 * reference, :136-137).  Operation order kept: ((r*r)/2)/(sigma*sigma).
 * ------------------------------------------------------------------------------------------ */
double orc_blasso_item(const double* X, const double* y, size_t n_rows, int p, const double* theta,
                       size_t item_rows, size_t position) {
    size_t r0 = position * item_rows, r1 = (position + 1) * item_rows;
    if (r0 > n_rows)
        r0 = n_rows;
    if (r1 > n_rows)
        r1 = n_rows;
    double sigma = theta[0];
    const double* beta = theta + 1;
    double c = log(1 / sigma / sqrt(2 * M_PI));
    double s2 = sigma * sigma;
    double total = 0.0;
    for (size_t r = r0; r != r1; ++r) {
        double dot = 0.0;
        for (int j = 0; j != p; ++j)
            dot += beta[j] * X[r * (size_t) p + j];
        double res = dot - y[r];
        total += c - res * res / 2 / s2;
    }
    return total;
}

void orc_blasso_eval(const double* X, const double* y, size_t n_rows, int p, const double* theta,
                     size_t item_rows, size_t first, size_t count, double* sum, double* sum_sq) {
    double s = *sum, q = *sum_sq;
    for (size_t pos = first; pos != first + count; ++pos) {
        double lp = orc_blasso_item(X, y, n_rows, p, theta, item_rows, pos);
        s += lp;
        q += lp * lp;
    }
    *sum = s;
    *sum_sq = q;
}

void orc_blasso_eval_ld(const double* X, const double* y, size_t n_rows, int p, const double* theta,
                        size_t item_rows, size_t n_items, long double* sum, long double* sum_sq) {
    long double sigma = theta[0];
    long double c = logl(1 / sigma / sqrtl(2 * 3.14159265358979323846264338327950288L));
    long double s = 0, q = 0;
    for (size_t pos = 0; pos != n_items; ++pos) {
        size_t r0 = pos * item_rows, r1 = (pos + 1) * item_rows;
        if (r0 > n_rows)
            r0 = n_rows;
        if (r1 > n_rows)
            r1 = n_rows;
        long double item = 0;
        for (size_t r = r0; r != r1; ++r) {
            long double dot = 0;
            for (int j = 0; j != p; ++j)
                dot += (long double) theta[1 + j] * X[r * (size_t) p + j];
            long double res = dot - y[r];
            item += c - res * res / 2 / (sigma * sigma);
        }
        s += item;
        q += item * item;
    }
    *sum = s;
    *sum_sq = q;
}

/* a8  BLasso.log_prior   pysamplers/pyblasso.py:123-133
 *   -2 log sigma + p log(tau/2/sigma) - (tau * sum|beta|) / sigma          (tau = 5, :32,70) */
double orc_blasso_log_prior(const double* theta, int p, double tau) {
    double sigma = theta[0];
    double l1 = 0.0;
    for (int j = 0; j != p; ++j)
        l1 += fabs(theta[1 + j]);
    double log_prior_sigma = -2.0 * log(sigma);
    double log_laplace = p * log(tau / 2 / sigma) - (tau * l1) / sigma;
    return log_prior_sigma + log_laplace;
}

/* ------------------------------------------------------------------------------------------
 * CPU baseline helpers ("port"): score `nodes` frontier nodes over all items with every
 * host thread (OpenMP over node x item-block).  The reference would give each node to
 * one MPI worker holding the full dataset; splitting items as well
 * only helps the CPU side.  Same arithmetic as above; per-item values are combined in
 * item order so the result does not depend on the thread count.
 * ------------------------------------------------------------------------------------------ */
void orc_mixture_frontier_mt(const double* data, size_t n_rows, int d, int K, const double* thetas,
                             size_t nodes, size_t item_rows, size_t n_items, double* sum,
                             double* sum_sq) {
    double* items = (double*) malloc(sizeof(double) * nodes * n_items);
#pragma omp parallel for collapse(2) schedule(static)
    for (size_t b = 0; b < nodes; ++b)
        for (size_t p = 0; p < n_items; ++p)
            items[b * n_items + p] =
                orc_mixture_item(data, n_rows, d, K, thetas + b * (size_t) (K * d), item_rows, p);
    for (size_t b = 0; b != nodes; ++b) {
        double s = 0, q = 0;
        for (size_t p = 0; p != n_items; ++p) {
            double lp = items[b * n_items + p];
            s += lp;
            q += lp * lp;
        }
        sum[b] = s;
        sum_sq[b] = q;
    }
    free(items);
}

void orc_blasso_frontier_mt(const double* X, const double* y, size_t n_rows, int p,
                            const double* thetas, size_t nodes, size_t item_rows, size_t n_items,
                            double* sum, double* sum_sq) {
    double* items = (double*) malloc(sizeof(double) * nodes * n_items);
#pragma omp parallel for collapse(2) schedule(static)
    for (size_t b = 0; b < nodes; ++b)
        for (size_t pos = 0; pos < n_items; ++pos)
            items[b * n_items + pos] =
                orc_blasso_item(X, y, n_rows, p, thetas + b * (size_t) (p + 1), item_rows, pos);
    for (size_t b = 0; b != nodes; ++b) {
        double s = 0, q = 0;
        for (size_t pos = 0; pos != n_items; ++pos) {
            double lp = items[b * n_items + pos];
            s += lp;
            q += lp * lp;
        }
        sum[b] = s;
        sum_sq[b] = q;
    }
    free(items);
}

void orc_normal_frontier_mt(const double* data, size_t n, int D, const double* thetas, size_t nodes,
                            double* sum, double* sum_sq) {
    int tl = D + D * (D + 1) / 2;
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t b = 0; b < nodes; ++b) {
        double s = 0, q = 0;
        orc_normal_eval(data, n, D, thetas + b * (size_t) tl, 1, 0, 0, n, &s, &q);
        sum[b] = s;
        sum_sq[b] = q;
    }
}

double orc_boltzmann_dense_logp(const double* W, size_t D, const double* x, double temperature) {
    double s = 0, q = 0;
    orc_boltzmann_dense_eval(W, D, x, temperature, 0, D, &s, &q);
    return s;
}

void orc_boltzmann_dense_frontier_mt(const double* W, size_t D, const double* X, size_t nodes,
                                     double temperature, double* sum, double* sum_sq) {
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t b = 0; b < nodes; ++b) {
        double lp = orc_boltzmann_dense_logp(W, D, X + b * D, temperature);
        sum[b] = lp;          /* one item: worker.hh:151-152 with a single evaluate() */
        sum_sq[b] = lp * lp;
    }
}

/* bench.py only: fix the number of OpenMP threads explicitly (torchrun exports OMP_NUM_THREADS=1, which would
 * otherwise turn the CPU baseline of a multi-GPU launch into a one-thread number); returns what is in force */
int orc_set_threads(int n) {
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    extern void omp_set_dynamic(int);
    extern int omp_get_max_threads(void);
    if (n > 0) {
        omp_set_dynamic(0);
        omp_set_num_threads(n);
    }
    return omp_get_max_threads();
#else
    (void) n;
    return 1;
#endif
}

int orc_max_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
