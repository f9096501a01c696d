// pf_resident.cu -- the RESIDENT frontier kernel of the pointwise models (normal target, mixture) for data sets that fit
// the shared memory of the whole GPU (N = 1e6 rows of two doubles = 16 MB over 296 CTAs: BASELINE configs 1 and 3).
//
// WHY.  On such a data set a frontier pass is a few microseconds of arithmetic (normal 1e6 x 8 nodes: 3 us at the FP64
// rate) behind ~8 us of launch + write-back + poll (pf_debug_launch_probe) and a prologue that re-stages tables, barriers
// and the rows themselves on every call.  The speculative chain (masterloop.cc:108-117) issues one such pass per
// scheduling burst, tens of thousands per second, so the kernel is launched ONCE and stays on the SMs between bursts:
//   * every CTA copies its contiguous row range into shared memory once (TMA bulk copies) and keeps it there -- later
//     calls read no global memory for the data at all;
//   * the exp / log tables, mbarriers and tile geometry survive across calls;
//   * the host rings a DOORBELL in mapped pinned memory: the command (header + thetas) travels as self-validating
//     8-byte words {call# << 32 | 32 payload bits} -- the mirror image of the result words (PointwiseArgs::done_flag) --
//     so one PCIe read round trip of CTA 0 both detects and fetches it, with no ordering between the words required;
//   * CTA 0 relays the decoded thetas through an L2-resident block with a release flag, every other CTA acquires it;
//   * scoring, per-CTA records, last-CTA finalize and the result words are pw_main<RES = true> -- the code of the
//     launched kernels reading resident rows instead of ring stages, so sums are bit-identical to the launched path
//     for the same geometry.
// The kernel leaves the SMs when the host asks (any other launch of this process on the device stops it first), or on
// its own after `idle_ns` without a command (default 2 ms), so it can never hold the GPU: the host notices through the
// state word and simply launches it again with the next call.
#include <stdio.h>
#include <string.h>

#include "pf_pointwise.cuh"

namespace pf {

static_assert(PW_FOLD, "the resident kernel is built on the producer-less CTA shape");

constexpr unsigned kResOpEval = 1u, kResOpQuit = 2u;
static_assert((size_t) (kResCmdWords - 1) * 4 <= (size_t) 2 * PW_CONSUMERS * sizeof(double), "command scratch fits the reduction area");

__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long res_next_seq(unsigned long long s) {
    ++s;
    if ((s & 0xFFFFFFFFull) == 0) ++s;  // the host never uses a tag of 0 (next_host_seq)
    return s;
}

// M0: the model (normal target: identity covariance); M1: the general-covariance variant of the normal target, or void.
// Two nodes per lane (frontiers of 33 .. 64 nodes in one command) for the normal target up to D = 2, like the launched path.
template <class M>
__host__ __device__ constexpr int res_max_npl() {
    return (!M::THETA_IN_SMEM && M::DIM <= 2) ? 2 : 1;
}

template <class M0, class M1, typename T>
__global__ void __launch_bounds__(PW_THREADS, 2) pw_resident_kernel(const __grid_constant__ ResidentArgs ra) {
    using M = M0;
    using CT = typename M::CT;
    constexpr int DIM = M::DIM;
    constexpr int RB_BYTES = DIM * (int) sizeof(T);
    constexpr unsigned TILE_ROWS = PW_TILE_BYTES / RB_BYTES;
    constexpr int NPL = 1;
    constexpr int NPLMAX = res_max_npl<M>();

    extern __shared__ __align__(128) unsigned char smem[];
    PwShared sh;
    sh.stage_base = smem;
    sh.red = reinterpret_cast<double*>(smem + ra.cap_bytes);
    sh.full = reinterpret_cast<uint64_t*>(sh.red + pw_red_doubles<M, NPLMAX>());
    sh.empty = sh.full + PW_STAGES;
    sh.flag = reinterpret_cast<int*>(sh.empty + PW_STAGES);
    CT* th_smem = reinterpret_cast<CT*>(reinterpret_cast<unsigned char*>(sh.flag) + 16);
    sh.th_smem = th_smem;
    sh.exp2_tab = nullptr;
    sh.log_tab = nullptr;

    const int tid = threadIdx.x;
    unsigned long long r0 = (unsigned long long) blockIdx.x * ra.rows_per_cta;
    if (r0 > ra.row_end) r0 = ra.row_end;
    unsigned long long r1 = r0 + ra.rows_per_cta;
    if (r1 > ra.row_end) r1 = ra.row_end;
    sh.r0 = r0;
    sh.r1 = r1;

    __shared__ unsigned long long theta_nan;
    __shared__ uint64_t loadbar;  // completion of the one-time copies (rows, tables)
    __shared__ int not_factorable;
    __shared__ unsigned s_cmdhdr;
    __shared__ int s_hdr_ok;
    __shared__ int s_quit;
    sh.theta_nan = &theta_nan;

    if (tid == 0) {
        for (int s = 0; s < PW_STAGES; ++s) {
            mbar_init(&sh.full[s], 1);
            mbar_init(&sh.empty[s], PW_CONSUMER_WARPS);
        }
        mbar_init(&loadbar, 1);
        fence_barrier_init();
        s_quit = 0;
    }
    __syncthreads();

    // ---------------------------------------------------------------- once per launch: tables and this CTA's rows
    constexpr bool BIG = big_exp_table<M>();
    const unsigned long long b0 = r0 * RB_BYTES, b1 = r1 * RB_BYTES;
    const unsigned long long w0 = b0 & ~15ull, w1 = (b1 + 15ull) & ~15ull;
    const unsigned char* const rows = smem + (unsigned) (b0 & 15ull);  // row r0 (the copy window is 16-byte aligned)
    {
        constexpr unsigned kTabBytesOnce = (M::THETA_IN_SMEM && BIG) ? (unsigned) (sizeof(double) * kExpRepSize * kExpRepCopies + sizeof(double2) * 256) : 0u;
        if (tid == 0) mbar_arrive_expect_tx(&loadbar, kTabBytesOnce + (unsigned) (w1 - w0));
    }
    if constexpr (M::THETA_IN_SMEM) {
        __shared__ __align__(128) double exp2_raw[BIG ? kExpRepSize * kExpRepCopies : 2 * kExp2TabSize];
        __shared__ __align__(16) double2 log_tab[256];
        constexpr unsigned kTabBytes = kExp2TabSize * 8;
        const unsigned raw_s = smem_u32(exp2_raw);
        double* const exp2_tab = BIG ? exp2_raw : exp2_raw + ((((raw_s + kTabBytes - 1) & ~(kTabBytes - 1)) - raw_s) >> 3);
        sh.exp2_tab = exp2_tab;
        sh.log_tab = log_tab;
        if constexpr (BIG) {
            constexpr unsigned kRepBytes = (unsigned) (sizeof(double) * kExpRepSize * kExpRepCopies);
            if (tid == 0) {
                bulk_g2s(exp2_tab, ra.exp2_rep, kRepBytes, &loadbar);
                bulk_g2s(log_tab, ra.log_mid, (unsigned) (sizeof(double2) * 256), &loadbar);
            }
        } else {
            for (int i = tid; i < kExp2TabSize; i += PW_CONSUMERS) exp2_tab[i] = kExp2Tab[i];
            for (int i = tid; i < 256; i += PW_CONSUMERS) log_tab[i] = make_double2(kLogTab[2 * i], kLogTab[2 * i + 1]);
        }
    }
    if (tid == 0) {
        // (the whole transaction count is announced before the first copy is issued: a barrier whose count reaches zero
        // between two announcements would complete early)
        for (unsigned long long o = w0; o < w1; o += 32768ull) {
            const unsigned bytes = (unsigned) (w1 - o < 32768ull ? w1 - o : 32768ull);
            bulk_g2s(smem + (o - w0), static_cast<const unsigned char*>(ra.data) + o, bytes, &loadbar);
        }
    }
    // tiles this CTA scores per call (every call covers the whole data set)
    unsigned ntiles = 0;
    {
        TileIter it{r0, r1, ra.item_rows, ra.row_end, TILE_ROWS, 0ull};
        unsigned long long t0, t1;
        while (it.next(t0, t1)) ++ntiles;
    }
    mbar_wait(&loadbar, 0);
    __syncthreads();

    // ---------------------------------------------------------------- the call loop
    unsigned long long seq = ra.first_seq;
    unsigned calls = 0, nw_poll = 1;
    unsigned reason = 0;
    unsigned long long idle0 = globaltimer_ns();
    // payload halves of the command words: the reduction area is free between calls
    unsigned* const cmdw = reinterpret_cast<unsigned*>(sh.red);
    const bool lead = blockIdx.x == 0;
    // CTA 0 polls the doorbell in mapped host memory and forwards every word of this call it finds -- as it is, tag and
    // all -- into the relay block in device memory; every other CTA polls that block the same way.  Nothing on the way
    // needs a fence or a flag: a word is either of this call or it is not.
    const unsigned long long* const src = lead ? ra.h_cmd : ra.d_relay;
    for (;;) {
        const unsigned tag = (unsigned) (seq & 0xFFFFFFFFull);
        unsigned hdr = kResOpQuit;
        bool got = false;
        const unsigned long long wait0 = globaltimer_ns();
        for (unsigned rounds = 1;; ++rounds) {
            // every thread fetches its share of the first nw_poll words and notes the first one that is not of this
            // call; how many words the command really has is only known from its header (word 0)
            unsigned first_bad = 0xFFFFFFFFu;
            for (unsigned w = tid; w < nw_poll; w += PW_CONSUMERS) {
                const unsigned long long v = ld_sys_u64(src + w);
                if ((unsigned) (v >> 32) != tag) {
                    if (w < first_bad) first_bad = w;
                    continue;
                }
                if (lead) *reinterpret_cast<volatile unsigned long long*>(ra.d_relay + w) = v;
                if (w == 0)
                    s_cmdhdr = (unsigned) v;
                else
                    cmdw[w - 1] = (unsigned) v;
            }
            if (tid == 0) s_hdr_ok = first_bad != 0u;
            if (tid == 32) {
                if (lead) {  // stop request / idle timeout, looked at while the command loads are in flight
                    const unsigned long long ctl = ld_sys_u64(ra.h_ctl);
                    if (ctl == ra.epoch)
                        s_quit = 1;
                    else if (globaltimer_ns() - idle0 > ra.idle_ns)
                        s_quit = 2;
                } else if ((rounds & 255u) == 0 && globaltimer_ns() - wait0 > ra.idle_ns + 3000000000ull)
                    s_quit = 3;  // CTA 0 always forwards a QUIT before it leaves; this only fires if it cannot (safety net)
            }
            __syncthreads();
            if (s_hdr_ok) {
                const unsigned h = s_cmdhdr;
                const unsigned nw = (h & 15u) == kResOpEval ? 1u + 2u * ((h >> 4) & 0xFFu) * ra.theta_len : 1u;
                const bool fetched = nw <= nw_poll;
                const bool complete = __syncthreads_and(first_bad >= nw) && fetched;
                nw_poll = nw;  // the next command is most likely of the same size: fetched in one round trip
                if (complete) {
                    got = true;  // (a complete command wins over a stop request or the idle timeout)
                    break;
                }
                continue;  // header seen, thetas not all here yet: one more round
            }
            if (s_quit) break;
            __syncthreads();  // (s_hdr_ok / s_quit are rewritten by the next round)
            if (!lead && rounds > 64u) __nanosleep(200);  // the host is thinking: ease off the L2
        }
        if (got) {
            hdr = s_cmdhdr;
            if (tid == 32 && s_quit == 2) s_quit = 0;  // not idle after all (read again behind the next barrier)
        } else if (lead) {
            reason = (unsigned) s_quit;
            if (tid == 0) *reinterpret_cast<volatile unsigned long long*>(ra.d_relay) = ((unsigned long long) tag << 32) | kResOpQuit;
        }
        trace_mark(ra.trace, 0, 0, tid == 0 && lead);  // CTA 0 has the command
        if ((hdr & 15u) != kResOpEval) break;
        trace_cta(ra.trace, 0, tid == 0);  // this CTA has the command

        const unsigned B = (hdr >> 4) & 0xFFu;
        const bool general = ((hdr >> 12) & 1u) != 0;
        // theta double i of the command, from its two payload halves
        auto theta_at = [&](unsigned i) {
            return __hiloint2double((int) cmdw[2 * i + 1], (int) cmdw[2 * i]);
        };

        ResidentCall c;
        c.theta = reinterpret_cast<const double*>(th_smem);
        c.part = ra.part;
        c.out = ra.out;
        c.ticket = ra.ticket;
        c.ticket_target = (unsigned long long) (calls + 1) * gridDim.x;
        c.row_begin = 0;
        c.row_end = ra.row_end;
        c.item_rows = ra.item_rows;
        c.rows_per_cta = ra.rows_per_cta;
        c.B = B;
        c.theta_len = ra.theta_len;
        c.K = ra.K;
        c.done_flag = reinterpret_cast<unsigned long long*>(ra.out);  // non-null: deliver self-validating words
        c.done_seq = seq;
        c.xchg = PeerXchg{};
        c.out_off = 0;
        c.out_stride = 0;
        c.trace = ra.trace;
        c.rows = rows;
        c.tile_base = calls * ntiles;

        if (tid == 0) {
            theta_nan = 0ull;
            not_factorable = 0;
        }
        if constexpr (M::THETA_IN_SMEM) {
            __syncthreads();
            using MF = typename FactoredOf<M>::type;
            [[maybe_unused]] bool factored = false;
            if constexpr (!std::is_void<MF>::value) {
                // the factored layout and its range test (see pw_frontier_body); thetas come from the relay block (L2)
                constexpr unsigned KC = MF::NCOMP;
                const unsigned stride = theta_smem_stride<double>(MF::staged_len(ra.theta_len));
                bool bad = false;
                for (unsigned i = tid; i < B * KC; i += PW_CONSUMERS) {
                    const unsigned node = i / KC, k = i - node * KC;
                    const unsigned th = node * ra.theta_len + k * DIM;
                    double* dst = reinterpret_cast<double*>(th_smem) + node * stride;
                    double nrm = 0.0, bound = 0.0;
#pragma unroll
                    for (int j = 0; j < DIM; ++j) {
                        const double v = theta_at(th + j), bj = v * kLog2e;
                        dst[k * DIM + j] = bj;
                        nrm = fma(v, v, nrm);
                        bound = fma(fabs(bj), ra.xmax[j], bound);
                    }
                    const double A = nrm * kNegHalfLog2e;
                    dst[KC * DIM + k] = A;
                    bad |= !(fabs(A) + bound <= kFactoredLimit);
                }
                if (bad) atomicOr(&not_factorable, 1);
                __syncthreads();
                factored = not_factorable == 0;
            }
            if (!factored) {
                const unsigned stride = theta_smem_stride<CT>(ra.theta_len);
                for (unsigned i = tid; i < B * ra.theta_len; i += PW_CONSUMERS) {
                    const unsigned node = i / ra.theta_len, j = i - node * ra.theta_len;
                    const double v = theta_at(i);
                    th_smem[node * stride + j] = (CT) v;
                    if constexpr (BIG)
                        if (v != v) atomicOr(&theta_nan, 1ull << node);
                }
                __syncthreads();
            }
            if (ra.trace != nullptr && tid == 0 && blockIdx.x == 0) ra.trace[6] = factored ? 1ull : 2ull;
            if constexpr (!std::is_void<MF>::value) {
                if (factored)
                    pw_main<MF, T, NPL, true, ResidentCall>(c, sh);
                else
                    pw_main<M, T, NPL, true, ResidentCall>(c, sh);
            } else
                pw_main<M, T, NPL, true, ResidentCall>(c, sh);
        } else {
            // normal target: the raw thetas into shared memory (a persistent kernel must not trust L1 for data that
            // changes between calls), Node::load reads them from there
            double* const raw = reinterpret_cast<double*>(th_smem);
            for (unsigned i = tid; i < B * ra.theta_len; i += PW_CONSUMERS) raw[i] = theta_at(i);
            __syncthreads();
            if constexpr (!std::is_void<M1>::value) {
                if constexpr (NPLMAX == 2) {
                    if (B > 32) {
                        if (general)
                            pw_main<M1, T, 2, true, ResidentCall>(c, sh);
                        else
                            pw_main<M, T, 2, true, ResidentCall>(c, sh);
                    } else if (general)
                        pw_main<M1, T, 1, true, ResidentCall>(c, sh);
                    else
                        pw_main<M, T, 1, true, ResidentCall>(c, sh);
                } else if (general)
                    pw_main<M1, T, NPL, true, ResidentCall>(c, sh);
                else
                    pw_main<M, T, NPL, true, ResidentCall>(c, sh);
            } else
                pw_main<M, T, NPL, true, ResidentCall>(c, sh);
        }
        ++calls;
        seq = res_next_seq(seq);
        idle0 = globaltimer_ns();
        __syncthreads();
    }
    // ---------------------------------------------------------------- leaving: tell the host (and why)
    if (blockIdx.x == 0 && tid == 0) {
        *reinterpret_cast<volatile unsigned long long*>(ra.h_state) = (ra.epoch << 32) | ((unsigned long long) reason << 8) | 1ull;
        __threadfence_system();
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
namespace {

template <class M>
size_t res_smem_bytes(unsigned cap_bytes, unsigned theta_len) {
    size_t b = (size_t) cap_bytes + (size_t) pw_red_doubles<M, res_max_npl<M>()>() * sizeof(double) + 2 * PW_STAGES * sizeof(uint64_t) + 32;
    // theta staging: 32 nodes of the largest layout a call may use; raw doubles (up to 64 nodes) for the normal target
    if (M::THETA_IN_SMEM)
        b += (size_t) 32 * theta_smem_stride<typename M::CT>(pw_staged_len<M>(theta_len)) * sizeof(typename M::CT);
    else
        b += (size_t) 32 * res_max_npl<M>() * theta_len * sizeof(double);
    return b;
}

template <class M0, class M1, typename T>
int res_configure(Engine& e) {
    Resident& r = e.res;
    const void* kern = reinterpret_cast<const void*>(&pw_resident_kernel<M0, M1, T>);
    const unsigned long long rows = e.n_items * e.item_rows;
    constexpr unsigned RB = M0::DIM * (unsigned) sizeof(T);
    // the geometry of the launched kernels (launch_one in pf_pointwise.cu: two CTAs per SM, never more CTAs than 16 KB
    // tiles), so that both paths add the same numbers in the same order: bit-identical sums
    constexpr unsigned TILE_ROWS = PW_TILE_BYTES / RB;
    unsigned long long grid = (unsigned long long) e.sm_count * 2;
    const unsigned long long tiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
    if (grid > tiles) grid = tiles;
    if (grid > (unsigned long long) kMaxGrid) grid = kMaxGrid;
    if (grid < 1) return 0;
    const unsigned long long per = (rows + grid - 1) / grid;
    grid = (rows + per - 1) / per;
    const unsigned long long cap = (per * RB + 32 + 127) & ~127ull;
    if (cap > 200000ull) return 0;
    const size_t smem = res_smem_bytes<M0>((unsigned) cap, e.theta_len);
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int dev_max = 0;
    cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, e.device);
    if (smem + fa.sharedSizeBytes > (size_t) dev_max) return 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess ||
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PW_THREADS, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if ((unsigned long long) occ * e.sm_count < grid) return 0;  // every CTA must be on an SM at the same time
    r.kern = kern;
    r.grid = (unsigned) grid;
    r.rows_per_cta = (unsigned) per;
    r.cap_bytes = (unsigned) cap;
    r.smem = smem;
    r.max_nodes = 32u * (unsigned) res_max_npl<M0>();
    return 1;
}

template <typename T>
int res_configure_t(Engine& e) {
    if (e.desc.model == PF_MODEL_NORMAL) {
        switch (e.dim) {
            case 1: return res_configure<NormalModel<1, T, false>, NormalModel<1, T, true>, T>(e);
            case 2: return res_configure<NormalModel<2, T, false>, NormalModel<2, T, true>, T>(e);
            case 3: return res_configure<NormalModel<3, T, false>, NormalModel<3, T, true>, T>(e);
            case 4: return res_configure<NormalModel<4, T, false>, NormalModel<4, T, true>, T>(e);
        }
        return 0;  // (D = 8: the launched general-covariance kernel runs one CTA per SM -- another geometry; not resident)
    }
    if (e.desc.model == PF_MODEL_MIXTURE) {
        if (e.K == 8 && e.dim == 2) return res_configure<MixtureModel<8, 2, T>, void, T>(e);
        if (e.K == 8 && e.dim == 8) return res_configure<MixtureModel<8, 8, T>, void, T>(e);
        if (e.K == 4 && e.dim == 2) return res_configure<MixtureModel<4, 2, T>, void, T>(e);
        if (e.K == 16 && e.dim == 2) return res_configure<MixtureModel<16, 2, T>, void, T>(e);
        if (e.K == 8 && e.dim == 4) return res_configure<MixtureModel<8, 4, T>, void, T>(e);
        if (e.dim == 2) switch (e.K) {
            case 2: return res_configure<MixtureModel<2, 2, T>, void, T>(e);
            case 3: return res_configure<MixtureModel<3, 2, T>, void, T>(e);
            case 5: return res_configure<MixtureModel<5, 2, T>, void, T>(e);
            case 6: return res_configure<MixtureModel<6, 2, T>, void, T>(e);
            default: break;
        }
        switch (e.dim) {
            case 1: return res_configure<MixtureModelRT<1, T>, void, T>(e);
            case 2: return res_configure<MixtureModelRT<2, T>, void, T>(e);
            case 3: return res_configure<MixtureModelRT<3, T>, void, T>(e);
            case 4: return res_configure<MixtureModelRT<4, T>, void, T>(e);
            case 8: return res_configure<MixtureModelRT<8, T>, void, T>(e);
        }
    }
    return 0;
}

}  // namespace

// Can this engine's data set live in shared memory?  Fills e.res geometry.  1 = yes, 0 = no (not an error).
int resident_configure(Engine& e) {
    if (e.desc.model != PF_MODEL_NORMAL && e.desc.model != PF_MODEL_MIXTURE) return 0;
    return e.f32 ? res_configure_t<float>(e) : res_configure_t<double>(e);
}

// Launch the resident kernel; the first command it accepts is call number `first_seq`.
int resident_launch(Engine& e, unsigned long long first_seq) {
    Resident& r = e.res;
    ResidentArgs a{};
    a.data = e.d_data;
    a.part = e.d_part;
    a.out = e.d_hout;
    a.ticket = r.d_ticket;
    a.h_cmd = r.d_hcmd;
    a.h_ctl = r.d_hcmd + kResCtlWord;
    a.h_state = r.d_hcmd + kResStateWord;
    a.d_relay = r.d_relay;
    a.exp2_rep = r.exp2_rep;
    a.log_mid = r.log_mid;
    a.row_end = e.n_items * e.item_rows;
    a.item_rows = e.item_rows;
    a.rows_per_cta = r.rows_per_cta;
    a.theta_len = e.theta_len;
    a.K = e.K;
    a.first_seq = first_seq;
    a.epoch = r.epoch;
    a.idle_ns = r.idle_ns;
    a.trace = e.d_trace;
    a.cap_bytes = r.cap_bytes;
    for (int j = 0; j < 8; ++j) a.xmax[j] = e.xmax[j];
    // (the relay block may hold the QUIT an earlier launch left for the very call number this one starts with)
    if (cudaMemsetAsync(r.d_ticket, 0, sizeof(unsigned long long), e.stream) != cudaSuccess ||
        cudaMemsetAsync(r.d_relay, 0, sizeof(unsigned long long), e.stream) != cudaSuccess) {
        set_error("resident kernel: ticket reset failed: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    void* params[] = {&a};
    const cudaError_t err = cudaLaunchCooperativeKernel(r.kern, dim3(r.grid), dim3(PW_THREADS), params, r.smem, e.stream);
    if (err != cudaSuccess) {
        set_error("resident kernel launch: %s", cudaGetErrorString(err));
        cudaGetLastError();
        return -1;
    }
    return 1;
}

}  // namespace pf
