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

#include "pf_pointwise.cuh"

namespace pf {

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
    if (e.K == 4 && e.dim == 2) return launch_npl<MixtureModel<4, 2, T>, T>(e, a, s, true);    // (K-unrolled loops: 2.3x the
    if (e.K == 16 && e.dim == 2) return launch_npl<MixtureModel<16, 2, T>, T>(e, a, s, true);  //  run-time component loop)
    if (e.K == 8 && e.dim == 4) return launch_npl<MixtureModel<8, 4, T>, T>(e, a, s, true);
    if (e.dim == 2) switch (e.K) {  // small two-dimensional mixtures
        case 2: return launch_npl<MixtureModel<2, 2, T>, T>(e, a, s, true);
        case 3: return launch_npl<MixtureModel<3, 2, T>, T>(e, a, s, true);
        case 5: return launch_npl<MixtureModel<5, 2, T>, T>(e, a, s, true);
        case 6: return launch_npl<MixtureModel<6, 2, T>, T>(e, a, s, true);
        default: break;
    }
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

// device addresses of the FP64 mixture loop's tables on e.device (uploaded on first use), for the resident kernel
int pointwise_table_ptrs(Engine& e, const double** exp2_rep, const double** log_mid) {
    if (ensure_exp_table(e) != 0) return -1;
    void *a = nullptr, *b = nullptr;
    if (cudaGetSymbolAddress(&a, g_exp2_rep) != cudaSuccess || cudaGetSymbolAddress(&b, g_log_mid) != cudaSuccess) {
        set_error("cudaGetSymbolAddress: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    *exp2_rep = static_cast<const double*>(a);
    *log_mid = static_cast<const double*>(b);
    return 0;
}


bool pointwise_supported(const Engine& e) {
    const uint32_t d = e.dim;
    return d == 1 || d == 2 || d == 3 || d == 4 || d == 8;
}

uint32_t pointwise_max_frontier(const Engine& e) {
    if (e.desc.model == PF_MODEL_NORMAL) return e.dim <= 2 ? 64 : 32;  // general-covariance variant bound
    if (e.desc.model == PF_MODEL_MIXTURE) return (e.K == 8 && e.dim == 8) ? 32 : 64;
    return 64;
}

// FP64 mixture: how a frontier of B nodes is cut into launches of <= 32 nodes.  The kernel is FP64-bound, so a launch
// costs what its LANE LAYOUT costs -- 11 .. 16 nodes all run with two point lanes (the time of 16), 17 .. 32 with one (the
// time of 32) -- and reading the data once more is free: 12 nodes as 8 + 4 take 7.3 ms per 1e8 rows instead of 9.1, 20 as
// 16 + 4 11.7 instead of 17.7.  Measured ms per 1e8 rows of one launch (K = 8, d = 2; other shapes scale alike) + 12 us per
// launch; the cheapest cut by dynamic programming.  (Engines of a multi-GPU group keep the fixed 32 + rest cut.)
static unsigned mixture_plan(unsigned B, unsigned long long rows, unsigned (&chunk)[64], double dtype_scale = 1.0) {
    static const double t1[33] = {0,    0.86, 1.41, 2.31, 2.55, 3.39, 3.94, 4.72, 4.72, 6.17, 6.15, 9.1,  9.1,  9.1,  9.1,  9.1, 9.1,
                                  17.7, 17.7, 17.7, 17.7, 17.7, 17.7, 17.7, 17.7, 17.7, 17.7, 17.7, 17.7, 17.7, 17.7, 17.7, 17.7};
    const double scale = (double) rows / 1e8 * dtype_scale, per_launch = 0.012;
    double cost[65];
    unsigned first[65];
    cost[0] = 0.0;
    first[0] = 0;
    for (unsigned b = 1; b <= B; ++b) {
        cost[b] = 1e300;
        for (unsigned c = 1; c <= 32 && c <= b; ++c) {
            const double v = cost[b - c] + t1[c] * scale + per_launch;
            if (v < cost[b] - 1e-12) {
                cost[b] = v;
                first[b] = c;
            }
        }
    }
    unsigned n = 0;
    for (unsigned b = B; b > 0; b -= first[b]) chunk[n++] = first[b];
    // widest launch first (deterministic order)
    for (unsigned i = 0; i + 1 < n; ++i)
        for (unsigned j = i + 1; j < n; ++j)
            if (chunk[j] > chunk[i]) {
                const unsigned t = chunk[i];
                chunk[i] = chunk[j];
                chunk[j] = t;
            }
    return n;
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
    // FP64 mixture: the frontier may be cut into several launches (mixture_plan)
    unsigned chunk[64], n_chunks = 1;
    chunk[0] = B;
    // PF_F32: the same lane layouts, 0.58x the time (12 / 16 nodes 5.08 ms per 1e8 rows, 24 / 32 9.78, and 33 .. 64 in one
    // launch with two nodes per lane 19.6: 40 nodes as 32 + 8 take 12.5) -- cut by the same plan on large single-GPU data
    // sets, otherwise launched as before (one launch)
    const bool cut_f32 = e.desc.model == PF_MODEL_MIXTURE && e.f32 && e.gather_per == 0 && B <= 64 && e.nranks <= 1 &&
                         n_items * e.item_rows >= 4000000ull;
    if (cut_f32) n_chunks = mixture_plan(B, n_items * e.item_rows, chunk, 0.58);
    const bool may_cut = e.desc.model == PF_MODEL_MIXTURE && !e.f32 && e.gather_per == 0 && B <= 64;
    if (may_cut) {
        // fixed cut (32 + rest) for the engines of a multi-GPU group -- their shares may be unequal and the cut must not
        // depend on the rows -- and for data sets within the resident kernel's reach: the resident kernel scores commands
        // of <= 32 nodes in one pass each, and the launched path reproduces its sums bit for bit
        if (e.nranks > 1 || n_items * e.item_rows < 4000000ull) {
            if (B > 32) {
                n_chunks = 2;
                chunk[0] = 32;
                chunk[1] = B - 32;
            }
        } else
            n_chunks = mixture_plan(B, n_items * e.item_rows, chunk);
    }
    if (a.row_end <= a.row_begin) {  // no rows here: zeros (+ the same exchanges the other ranks' launches make)
        int total = 0;
        for (unsigned i = 0, off = 0; i < n_chunks; off += chunk[i], ++i) {
            const int r = n_chunks == 1 ? launch_exchange_only(e, B, d_out, s) : launch_exchange_only(e, chunk[i], d_out, s, off, B);
            if (r < 0) return r;
            total += r;
        }
        return total;
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
    if ((may_cut || cut_f32) && n_chunks > 1) {
        // several launches of <= 32 nodes (one node per lane: two per lane run out of registers -- N = 1e8, B = 64: 45.5 vs
        // 52.7 ms); the data set is read once per launch, which an FP64-bound kernel does not notice.  Each launch makes
        // its own peer exchange.
        int total = 0;
        for (unsigned i = 0, off = 0; i < n_chunks; off += chunk[i], ++i) {
            PointwiseArgs c = a;
            c.B = chunk[i];
            c.theta = d_theta + (size_t) off * e.theta_len;
            c.out_off = off;
            c.out_stride = B;
            c.theta_inline_n = 0;
            if (a.theta_inline_n) {  // (tiny thetas only) the inline copy holds all B nodes: re-slice it
                c.theta_inline_n = chunk[i] * e.theta_len;
                memcpy(c.theta_inline, a.theta_inline + (size_t) off * e.theta_len, (size_t) c.theta_inline_n * sizeof(double));
            }
            if (i > 0) c.xchg = e.next_xchg();
            const int r = e.f32 ? dispatch_mixture<float>(e, c, s) : dispatch_mixture<double>(e, c, s);
            if (r < 0) return r;
            total += r;
        }
        return total;
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
