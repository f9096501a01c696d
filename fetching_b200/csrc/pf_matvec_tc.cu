// pf_matvec_tc.cu -- PF_F32 lasso frontier kernel on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   pyblasso.py:135-141   lp_item = sum_rows [ log(1/sigma/sqrt(2pi)) - (beta.x_r - y_r)^2 / 2 / sigma^2 ]
//
// The contraction X[rows x p] . beta[p x B] is a real dense GEMM, and in PF_F32 mode (data stored as float, 1e-5
// contract) it may run on the tensor cores as long as it stays FP32-FAITHFUL.  tcgen05 has no FP32 kind; kind::tf32
// reads FP32 words and uses their top 19 bits.  So every operand is split   v = hi + lo,   hi = the tf32 the hardware
// sees in v, lo = v - hi rounded to tf32, and the product is formed as   hi.hi + lo.hi + hi.lo   with FP32
// accumulation in TMEM ("3xTF32": ~2^-21 per operand instead of tf32's 2^-10; the lo.lo term is below 2^-20 of that).
//
// DESIGN (B200).  One CTA per SM, 192 threads:
//   warp 4 lane 0   producer: per (128-row tile, 32-column chunk) one 2-D TMA box of X (128 x 128 B, SWIZZLE_128B --
//                   exactly the canonical K-major UMMA operand layout) + the chunk's beta_hi / beta_lo panels (written
//                   by the prep kernel in that same swizzled layout, so a plain bulk copy lands them ready to use);
//   warps 0-3       split: X_lo = X - tf32(X) for the landed tile, element for element at the same shared-memory
//                   offsets (the swizzle is a permutation of 16-byte vectors, an elementwise pass does not care),
//                   conflict-free LDS.128 / STS.128, then fence.proxy.async so the tensor core's async proxy sees it;
//   warp 5 lane 0   MMA issuer: per chunk 4 k-steps x {hi.hi, lo.hi, hi.lo} = 12 tcgen05.mma.kind::tf32
//                   (M = 128, N = Npad, K = 8) into one FP32 accumulator tile in TMEM; tcgen05.commit frees the stage;
//   warps 0-3       epilogue per tile: tcgen05.ld (32x32b: thread = row) -> residual, square, per-node constants in
//                   FP64 -> item totals exactly like pf_matvec.cu (ItemOwner / finalize_items / peer exchange).
// Shared memory: a deep ring of landed X boxes with their panels (8 stages at N = 16, 5 at N = 64: what keeps HBM busy)
// and a 2-stage ring of X_lo tiles, which only live from the split to the completion of the MMAs that read them.
#include <math.h>
#include <stdio.h>

#include "pf_common.cuh"
#include "pf_engine.h"

namespace pf {

constexpr int TC_ROWS = 128;  // UMMA M: rows per tile
constexpr int TC_KCH = 32;    // floats per 128-byte chunk
constexpr int TC_MAX_STAGES = 8;   // ring of landed X boxes (+ their panels): sized at launch to what shared memory holds
#ifndef PF_TC_LO_STAGES
#define PF_TC_LO_STAGES 2
#endif
#ifndef PF_TC_EXP
#define PF_TC_EXP 0  // timing experiments (results WRONG unless 0): 1 no split arithmetic, 2 no lo MMAs, 3 no proxy fence
#endif
constexpr int TC_LO_STAGES = PF_TC_LO_STAGES;    // ring of X_lo tiles: produced right before the MMAs that read them
constexpr int TC_XBYTES = TC_ROWS * 128;                   // 16 KB
constexpr int TC_WORKERS = 128;                            // split + epilogue threads (warps 0-3)
constexpr int TC_THREADS = 192;                            // + producer warp + MMA warp
constexpr int TC_RED_DOUBLES = 2 * 2 * 4 * kMaxNodesPerPass;  // 2 buffers x (part A | part B) x 4 warps x nodes
constexpr int TC_TMEM_COLS = 64;                           // accumulator: 128 lanes x Npad (<= 64) FP32 columns
constexpr size_t kTcSmemFixed = 1024 /* alignment slack */ + (size_t) TC_LO_STAGES * TC_XBYTES + TC_RED_DOUBLES * sizeof(double) +
                                (2 * TC_MAX_STAGES + 2 * TC_LO_STAGES + 2) * sizeof(uint64_t) + 64;
constexpr size_t kTcSmemBudget = 224 * 1024;               // of the 227 KB a CTA may have (node constants are static)
// one stage of the X ring: the 128 x 128 B box + the chunk's beta_hi and beta_lo panels (Npad x 128 B each)
__host__ __device__ constexpr unsigned tc_stage_bytes(unsigned Npad) { return TC_XBYTES + 2 * Npad * 128; }
inline unsigned tc_stages_for(unsigned Npad) {
    const size_t n = (kTcSmemBudget - kTcSmemFixed) / tc_stage_bytes(Npad);
    return n > (size_t) TC_MAX_STAGES ? TC_MAX_STAGES : (unsigned) n;
}

// ---- tcgen05 wrappers (PTX ISA 8.6+, sm_100a) -------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major operand, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (1), version 1
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
    uint64_t d = (uint64_t) ((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t) 1 << 16;
    d |= (uint64_t) (1024 >> 4) << 32;
    d |= (uint64_t) 1 << 46;
    d |= (uint64_t) 2 << 61;
    return d;
}
// D[tmem] (+)= A[smem] . B[smem]^T, kind::tf32, issued by ONE thread for the CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when they are done (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// tf32 pieces of an FP32 / FP64 value.  The tensor core TRUNCATES an FP32 word to tf32 (drops the low 13 mantissa bits).
__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
// bits + half a tf32 ulp: the hardware's truncation of the result is then a round-to-nearest of v
__device__ __forceinline__ float tf32_pre_round(float v) { return __uint_as_float(__float_as_uint(v) + 0x1000u); }

// The coefficient panel for the tensor path: for 32-column chunk c and part h in {hi, lo} an [Npad rows = nodes][32
// k] K-major tile of 128-byte rows in the 128-byte-swizzled layout the MMA descriptor expects -- element (n, kk) of part
// h lives at   ((c * 2 + h) * Npad + n) * 128 + (((kk / 4) ^ (n & 7)) * 16) + (kk % 4) * 4   bytes.
// hi = beta rounded to tf32, lo = (beta - hi) rounded to tf32, both from the DOUBLE theta: beta is good to ~2^-22.
__global__ void mv_prep_panel_tc32(const double* theta, float* panel, unsigned B, unsigned Npad, unsigned theta_len,
                                   unsigned off, unsigned cols, unsigned cols_pad) {
    asm volatile("griddepcontrol.launch_dependents;");  // the frontier kernel may start its prologue now
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cols_pad * Npad) return;
    const unsigned n = idx % Npad, k = idx / Npad, c = k / TC_KCH, kk = k % TC_KCH;
    double v = 0.0;
    if (k < cols && n < B) v = theta[(size_t) n * theta_len + off + k];
    const float hi = tf32_trunc(tf32_pre_round((float) v));
    const float lo = tf32_trunc(tf32_pre_round((float) (v - (double) hi)));
    const size_t word = (((kk >> 2) ^ (n & 7u)) << 2) + (kk & 3u);
    panel[((size_t) (c * 2 + 0) * Npad + n) * 32 + word] = hi;
    panel[((size_t) (c * 2 + 1) * Npad + n) * 32 + word] = lo;
}

__global__ void __launch_bounds__(TC_THREADS, 1) mv_tc32_frontier_kernel(const __grid_constant__ MatvecArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const unsigned NS = a.n_units;  // stages of the X ring (tc_stages_for(Npad), passed by the launcher)
    const unsigned stage_bytes = tc_stage_bytes(a.Bpad);
    unsigned char* lo_ring = smem + (size_t) NS * stage_bytes;  // (stage_bytes is a multiple of 1024)
    double* red = reinterpret_cast<double*>(lo_ring + TC_LO_STAGES * TC_XBYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(red + TC_RED_DOUBLES);  // X box + panels landed
    uint64_t* empty = full + TC_MAX_STAGES;                              // the MMAs that read the stage are done
    uint64_t* split = empty + TC_MAX_STAGES;                             // X_lo written (per lo stage)
    uint64_t* lo_free = split + TC_LO_STAGES;                            // the MMAs that read the lo stage are done
    uint64_t* acc_full = lo_free + TC_LO_STAGES;                         // the tile's accumulator is complete
    uint64_t* acc_empty = acc_full + 1;                                  // ... and has been read
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
    int* flag = reinterpret_cast<int*>(tmem_slot + 2);
    __shared__ double2 node_cst[kMaxNodesPerPass];  // {log(1/sigma/sqrt(2 pi)), 1/(2 sigma^2)}

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned B = a.B, Npad = a.Bpad;  // Npad: multiple of 16, <= 64
    const unsigned n_chunks = a.cols_pad / TC_KCH;
    const unsigned pbytes = Npad * 128;  // one panel part

    unsigned long long r0 = a.row_begin + (unsigned long long) blockIdx.x * a.rows_per_cta;
    if (r0 > a.row_end) r0 = a.row_end;
    unsigned long long r1 = r0 + a.rows_per_cta;
    if (r1 > a.row_end) r1 = a.row_end;
    // tiles of this CTA: 128 rows, cut at item boundaries when items are shorter than a tile
    auto next_tile = [&](unsigned long long cur, unsigned long long& t1) {
        unsigned long long e = cur + TC_ROWS;
        if (e > r1) e = r1;
        if (a.item_rows < (unsigned long long) TC_ROWS) {
            unsigned long long ie = (cur / a.item_rows + 1) * a.item_rows;
            if (ie > a.row_end) ie = a.row_end;
            if (e > ie) e = ie;
        }
        t1 = e;
    };

    const bool tr0 = tid == 0 && blockIdx.x == 0;
    trace_mark(a.trace, 0, 0, tr0);
    trace_cta(a.trace, 0, tid == 0);
    if (tid == 0) {
        for (unsigned s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < TC_LO_STAGES; ++s) {
            mbar_init(&split[s], 4);
            mbar_init(&lo_free[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        fence_barrier_init();
    }
    if (warp == 5) {  // TMEM: one warp allocates (and later frees) the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < kMaxNodesPerPass) {
        const unsigned node = (unsigned) tid < B ? tid : B - 1;
        const double sigma = a.theta[(size_t) node * a.theta_len];
        node_cst[tid] = make_double2(log(1 / sigma / sqrt(2 * 3.14159265358979323846)), 1.0 / (2.0 * sigma * sigma));  // pyblasso.py:140
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 4) {
        // ------------------------------------------------------------------ producer
        if (lane == 0 && r0 < r1) {
            prefetch_tensormap(&a.tmap);
            asm volatile("griddepcontrol.wait;" ::: "memory");  // the panel (prep kernel) is first touched below
            unsigned q = 0;
            for (unsigned long long t0 = r0, t1; t0 < r1; t0 = t1) {
                next_tile(t0, t1);
                for (unsigned c = 0; c < n_chunks; ++c, ++q) {
                    const unsigned s = q % NS;
                    if (q >= NS) mbar_wait(&empty[s], ((q / NS) - 1) & 1);
                    unsigned char* st = smem + (size_t) s * stage_bytes;
                    mbar_arrive_expect_tx(&full[s], TC_XBYTES + 2 * pbytes);
                    tma_load_2d(st, &a.tmap, (int) (c * TC_KCH), (int) t0, &full[s]);
                    // beta_hi | beta_lo of this chunk are adjacent in the panel: one copy
                    bulk_g2s(st + TC_XBYTES, static_cast<const unsigned char*>(a.betaT) + (size_t) c * 2 * pbytes, 2 * pbytes,
                             &full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0 && r0 < r1) {
            // instruction descriptor: D = F32, A = B = TF32, both K-major, N = Npad, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((Npad >> 3) << 17) | ((TC_ROWS >> 4) << 24);
            unsigned q = 0, tile = 0;
            for (unsigned long long t0 = r0, t1; t0 < r1; t0 = t1, ++tile) {
                next_tile(t0, t1);
                if (tile > 0) mbar_wait(acc_empty, (tile - 1) & 1);  // the previous tile's accumulator has been read
                tc_fence_after();
                for (unsigned c = 0; c < n_chunks; ++c, ++q) {
                    const unsigned s = q % NS, sl = q % TC_LO_STAGES;
                    mbar_wait(&split[sl], (q / TC_LO_STAGES) & 1);  // X_hi landed (the split warps saw it), X_lo written
                    tc_fence_after();
                    const uint32_t sb = smem_u32(smem + (size_t) s * stage_bytes);
                    const uint64_t a_hi = umma_desc_k_sw128(sb), a_lo = umma_desc_k_sw128(smem_u32(lo_ring + sl * TC_XBYTES));
                    const uint64_t b_hi = umma_desc_k_sw128(sb + TC_XBYTES), b_lo = umma_desc_k_sw128(sb + TC_XBYTES + pbytes);
#if PF_TC_EXP != 4
#pragma unroll
                    for (unsigned k = 0; k < 4; ++k) {  // 8 floats = 32 bytes = 2 descriptor units per k-step
                        umma_tf32(tmem_acc, a_hi + 2 * k, b_hi + 2 * k, idesc, (c | k) != 0);
#if PF_TC_EXP != 2
                        umma_tf32(tmem_acc, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
                        umma_tf32(tmem_acc, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
#endif
                    }
#endif
                    umma_commit(&empty[s]);     // the stage may be refilled once these MMAs have read it
                    umma_commit(&lo_free[sl]);  // ... and the X_lo tile overwritten
                }
                umma_commit(acc_full);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ split + epilogue warps (0..3)
        ItemOwner own;
        own.init(r0, a.item_rows);
        unsigned q = 0, tile = 0;
        for (unsigned long long t0 = r0, t1; t0 < r1; t0 = t1, ++tile) {
            next_tile(t0, t1);
            for (unsigned c = 0; c < n_chunks; ++c, ++q) {
                const unsigned s = q % NS, sl = q % TC_LO_STAGES;
                if (q >= TC_LO_STAGES) mbar_wait(&lo_free[sl], ((q / TC_LO_STAGES) - 1) & 1);
                mbar_wait(&full[s], (q / NS) & 1);
                if (q == 0) trace_mark(a.trace, 0, 2, tr0);  // first box landed
                const uint4* hi = reinterpret_cast<const uint4*>(smem + (size_t) s * stage_bytes);
                uint4* lo = reinterpret_cast<uint4*>(lo_ring + sl * TC_XBYTES);
#if PF_TC_EXP != 1 && PF_TC_EXP != 4 && PF_TC_EXP != 5
#pragma unroll
                for (int i = 0; i < TC_XBYTES / 16 / TC_WORKERS; ++i) {  // 8 vectors per thread, consecutive threads consecutive
                    const uint4 v = hi[i * TC_WORKERS + tid];
                    uint4 w;
                    w.x = __float_as_uint(tf32_pre_round(__uint_as_float(v.x) - tf32_trunc(__uint_as_float(v.x))));
                    w.y = __float_as_uint(tf32_pre_round(__uint_as_float(v.y) - tf32_trunc(__uint_as_float(v.y))));
                    w.z = __float_as_uint(tf32_pre_round(__uint_as_float(v.z) - tf32_trunc(__uint_as_float(v.z))));
                    w.w = __float_as_uint(tf32_pre_round(__uint_as_float(v.w) - tf32_trunc(__uint_as_float(v.w))));
                    lo[i * TC_WORKERS + tid] = w;
                }
#endif
#if PF_TC_EXP != 3
                fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
#endif
                __syncwarp();
                if (lane == 0) mbar_arrive(&split[sl]);
            }
            // ---- epilogue of this tile: thread = row
            mbar_wait(acc_full, tile & 1);
            tc_fence_after();
            unsigned long long ts = t1;  // a tile holds at most ONE item boundary (items shorter than a tile cut the tiles)
            {
                const unsigned long long ie = (t0 / a.item_rows + 1) * a.item_rows;
                if (ie < t1) ts = ie;
            }
            const bool has_b = ts < t1;
            const unsigned long long r = t0 + tid;
            const bool live = r < t1, first = r < ts;
            const double yv = live ? __ldg(&a.y[r]) : 0.0;
            double* rbuf = red + (tile & 1) * (2 * 4 * kMaxNodesPerPass);
            for (unsigned n0 = 0; n0 < Npad; n0 += 16) {
                uint32_t d[16];
                tmem_ld16(tmem_acc + ((uint32_t) (warp * 32) << 16) + n0, d);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const unsigned node = n0 + j;
                    if (node >= B) break;  // (uniform)
                    const double2 ch = node_cst[node];
                    const double res = (double) __uint_as_float(d[j]) - yv;
                    const double lp = live ? ch.x - res * res * ch.y : 0.0;
                    double pa = first ? lp : 0.0;
                    for (int m = 16; m >= 1; m >>= 1) pa += shfl_xor_f64(pa, m);
                    if (lane == 0) rbuf[warp * kMaxNodesPerPass + node] = pa;
                    if (has_b) {
                        double pb = first ? 0.0 : lp;
                        for (int m = 16; m >= 1; m >>= 1) pb += shfl_xor_f64(pb, m);
                        if (lane == 0) rbuf[(4 + warp) * kMaxNodesPerPass + node] = pb;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);  // the accumulator may be overwritten
            named_bar_sync(1, TC_WORKERS);
            if ((unsigned) tid < B) {
                double tot = 0.0;
                for (unsigned w = 0; w < 4; ++w) tot += rbuf[w * kMaxNodesPerPass + tid];
                own.add_tile(tot, ts, a.item_rows, a.row_end);
                if (has_b) {
                    double totB = 0.0;
                    for (unsigned w = 0; w < 4; ++w) totB += rbuf[(4 + w) * kMaxNodesPerPass + tid];
                    own.add_tile(totB, t1, a.item_rows, a.row_end);
                }
            }
        }
        trace_mark(a.trace, 0, 4, tr0);  // all tiles done
        double* rec = a.part + blockIdx.x;  // column of this CTA in the [slot][node][CTA] records
        if ((unsigned) tid < B) {
            own.finish(r0, r1, a.row_begin, a.row_end, a.rows_per_cta, a.item_rows);
            own.store(rec, tid);
            __threadfence();
        }
        named_bar_sync(1, TC_WORKERS);
        if (tid == 0) {
            const unsigned prev = atomicAdd(a.ticket, 1u);
            flag[0] = (prev == gridDim.x - 1);
        }
        named_bar_sync(1, TC_WORKERS);
        trace_mark(a.trace, 0, 5, tr0);
        trace_cta(a.trace, 1, tid == 0);
        if (flag[0]) {  // last CTA: finalize (item stitching, exchange between GPUs, delivery)
            trace_mark(a.trace, 1, 0, tid == 0);
            __threadfence();
            double S, Q;
            finalize_items(a.part, gridDim.x, a.row_begin, a.row_end, a.rows_per_cta, a.item_rows, tid, B, red, 1,
                           TC_WORKERS, S, Q);
            trace_mark(a.trace, 1, 1, tid == 0);
            unsigned Bout = B;
            if (a.xchg.nranks > 1) {
                const double2 t = peer_allreduce(a.xchg, tid, B, TC_WORKERS, 1, red, S, Q);
                S = t.x;
                Q = t.y;
                if (a.xchg.gather_per) Bout = a.xchg.gather_total;
            }
            if ((unsigned) tid < Bout) {
                if (a.done_flag) {
                    store_result_words(a.out, tid, S, a.done_seq);
                    store_result_words(a.out, Bout + tid, Q, a.done_seq);
                } else {
                    a.out[tid] = S;
                    a.out[Bout + tid] = Q;
                }
            }
            if (tid == 0) *a.ticket = 0u;
            trace_mark(a.trace, 1, 2, tid == 0);
        }
    }

    // ---- teardown: every tcgen05 operation of this CTA is complete (the epilogue waited for the last commit)
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(TC_TMEM_COLS) : "memory");
    }
}

// host side -----------------------------------------------------------------------------------------------------------
int launch_matvec_tc32(Engine& e, MatvecArgs& a, const double* d_theta, cudaStream_t s) {
    static bool configured_dev[64] = {};
    if (!configured_dev[e.device & 63]) {
        cudaError_t err = cudaFuncSetAttribute(mv_tc32_frontier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int) kTcSmemBudget);
        if (err == cudaSuccess)
            err = cudaFuncSetAttribute(mv_tc32_frontier_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared);
        if (err == cudaSuccess)
            err = cudaFuncSetAttribute(mv_prep_panel_tc32, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) {
            set_error("cudaFuncSetAttribute(matvec tc32): %s", cudaGetErrorString(err));
            return -1;
        }
        configured_dev[e.device & 63] = true;
    }
    a.Bpad = (a.B + 15) / 16 * 16;  // UMMA N: a multiple of 16 for M = 128
    a.tile_rows = TC_ROWS;
    const unsigned long long rows = a.row_end - a.row_begin;
    // contiguous rows per CTA, a whole number of 128-row tiles (no row fetched twice)
    const unsigned long long tiles = (rows + TC_ROWS - 1) / TC_ROWS;
    unsigned long long g = tiles < (unsigned long long) e.sm_count ? tiles : (unsigned long long) e.sm_count;
    if (g < 1) g = 1;
    const unsigned long long tpc = (tiles + g - 1) / g;
    a.rows_per_cta = tpc * TC_ROWS;
    a.k_splits = 1;
    a.chunks_per_split = a.cols_pad / TC_KCH;
    a.n_units = tc_stages_for(a.Bpad);  // stages of the X ring (8 at N = 16, 5 at N = 64)
    const size_t smem_bytes = kTcSmemFixed + (size_t) a.n_units * tc_stage_bytes(a.Bpad);
    unsigned grid = (unsigned) ((rows + a.rows_per_cta - 1) / a.rows_per_cta);
    if (grid < 1) grid = 1;
    if (make_tmap(e, TC_ROWS, &a.tmap) != 0) return -1;
    const unsigned total = e.cols_pad * a.Bpad;
    mv_prep_panel_tc32<<<(total + 255) / 256, 256, 0, s>>>(d_theta, static_cast<float*>(e.d_betaT), a.B, a.Bpad, e.theta_len, 1,
                                                          e.dim, e.cols_pad);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("mv_prep_panel_tc32 launch: %s", cudaGetErrorString(err));
        return -1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // start under the tail of the prep kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    err = cudaLaunchKernelEx(&cfg, mv_tc32_frontier_kernel, a);
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("matvec tc32 kernel launch: %s", cudaGetErrorString(err));
        return -1;
    }
    return 2;
}

}  // namespace pf
