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
// DESIGN (B200).  One CTA per SM, 320 threads (DESIGN.md section 4.3 has the measurements behind every choice):
//   warp 4          TMA producer (whole warp walks, an elected lane issues): per step -- one 128-row tile x one 32-column
//                   chunk -- two 2-D boxes of X (64 rows x 128 B each, SWIZZLE_128B, L2 evict_first) into a deep ring;
//                   per chunk, ONCE for all tiles of the group, the beta_hi | beta_lo panel (written by the prep kernel in
//                   the swizzled K-major UMMA layout, L2 evict_last);
//   warps 0-3, 6-9  two split teams on alternate steps (one warp per TMEM lane quarter): thread = row reads the landed box
//                   (8 x LDS.128 through the swizzle), releases the shared-memory stage, and stores X_hi (the raw words)
//                   and X_lo = X - tf32(X) into the team's A stage in TENSOR MEMORY (tcgen05.st);
//   warp 5          MMA issuer (whole warp walks, an elected lane issues): per step 4 k-steps x {hi.hi, lo.hi, hi.lo} = 12
//                   tcgen05.mma.kind::tf32 with A from TMEM and B by shared-memory descriptor (M = 128, N = Npad, K = 8)
//                   into the tile's own FP32 accumulator in TMEM -- (512 - 128) / Npad accumulators: the chunk loop runs
//                   over all tiles of a group; tcgen05.commit frees the A stage / the panel stage / hands over the group;
//   warps 0-3       epilogue per item segment: tcgen05.ld (thread = row) -> residual, square, per-node constants in FP64,
//                   summed per thread over the segment's tiles, one transposed 16-value warp reduction -> item totals
//                   exactly like pf_matvec.cu (ItemOwner / finalize_items / peer exchange; both teams run the finalize).
#include <math.h>
#include <stdio.h>

#include "pf_common.cuh"
#include "pf_engine.h"

namespace pf {

// -DPF_TC_TIMELINE: CTA 0 stamps the first 200 steps of every role into the trace buffer (PF_TRACE=1 prints them)
#ifdef PF_TC_TIMELINE
#define TC_TL(role, q) do { if (blockIdx.x == 0 && a.trace != nullptr && (q) < 200u) a.trace[400 + (role) * 200 + (q)] = globaltimer_ns(); } while (0)
#else
#define TC_TL(role, q) do { } while (0)
#endif

constexpr int TC_ROWS = 128;  // UMMA M: rows per tile
constexpr int TC_BOX_ROWS = 64;  // rows per TMA box: two boxes per tile, one for a tile of <= 64 rows (row ranges of the
                                 // CTAs are multiples of 64 rows, so the last tile of a CTA fetches no row it does not score)
constexpr int TC_KCH = 32;    // floats per 128-byte chunk
constexpr int TC_MAX_STAGES = 12;  // ring of landed X tiles: sized at launch to what shared memory holds
constexpr int TC_P_STAGES = 2;                   // ring of coefficient panels (one per 32-column chunk, shared by all tiles)
constexpr int TC_XBYTES = TC_ROWS * 128;                   // 16 KB
constexpr int TC_WORKERS = 128;                            // threads of a split team (4 warps: one per TMEM lane quarter)
constexpr int TC_TEAMS = 2;                                // split teams: team t takes the steps q = t (mod 2), A stage t
constexpr int TC_THREADS = 192 + (TC_TEAMS - 1) * 128;     // team 0 (also the epilogue) | producer warp | MMA warp | team 1
constexpr int TC_RED_DOUBLES = 2 * 2 * 4 * kMaxNodesPerPass;  // reduction / finalize scratch
constexpr int TC_TMEM_COLS = 512;                          // all of TMEM
constexpr int TC_A_COLS = 64;                              // one A stage in TMEM: X_hi (32 columns) | X_lo (32 columns)
constexpr int TC_A_STAGES = TC_TEAMS;                      // one per team.  (Measured alternative: X_hi read from shared memory
                                                           // by descriptor and four X_lo-only stages -- 6 % slower: the MMAs with
                                                           // A in shared memory take longer to issue, and the step is bounded
                                                           // by the MMA thread, not by the split teams.)
constexpr int TC_A_BASE = TC_TMEM_COLS - TC_A_STAGES * TC_A_COLS;  // accumulator tiles live below: (512 - 128) / Npad of them
constexpr size_t kTcSmemFixed = 1024 /* alignment slack */ + TC_RED_DOUBLES * sizeof(double) +
                                (2 * TC_MAX_STAGES + 2 * TC_P_STAGES + 2 * TC_A_STAGES + 2) * sizeof(uint64_t) + 64;
constexpr size_t kTcSmemBudget = 224 * 1024;               // of the 227 KB a CTA may have (node constants are static)
// one stage of the panel ring: the chunk's beta_hi and beta_lo panels (Npad x 128 B each)
__host__ __device__ constexpr unsigned tc_panel_bytes(unsigned Npad) { return 2 * Npad * 128; }
inline unsigned tc_stages_for(unsigned Npad) {
    const size_t n = (kTcSmemBudget - kTcSmemFixed - (size_t) TC_P_STAGES * tc_panel_bytes(Npad)) / TC_XBYTES;
    return n > (size_t) TC_MAX_STAGES ? TC_MAX_STAGES : (unsigned) n;
}

// ---- tcgen05 wrappers (PTX ISA 8.6+, sm_100a) -------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one lane of a converged warp, chosen by the hardware: inside `if (elect_one())` ptxas knows that a single thread runs,
// so uniform operands reach the tensor-core / TMA instructions straight from uniform registers
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

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
// same with the A operand in TENSOR MEMORY (lane = row, one 32-bit column per k): [tmem_a] instead of a descriptor
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 consecutive columns of this thread's TMEM lane <- 32 registers (32x32b: lane = 32 (warp % 4) + lane id)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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

// Sums of 16 per-lane values over the 32 lanes of a warp in 16 shuffles instead of 80: in every round a lane keeps half of
// its values and hands the other half to its partner.  Returns the total of value `warp_reduce16_index(lane)` (every
// lane pair (2k, 2k + 1) ends with the same one).
__device__ __forceinline__ unsigned warp_reduce16_index(unsigned lane) {
    return ((lane >> 4) & 1u) * 8u + ((lane >> 3) & 1u) * 4u + ((lane >> 2) & 1u) * 2u + ((lane >> 1) & 1u);
}
__device__ __forceinline__ double warp_reduce16(double (&v)[16], unsigned lane) {
#pragma unroll
    for (int h = 8; h >= 1; h >>= 1) {  // partner = lane ^ (2 h): the upper lane of the pair keeps v[h .. 2h)
        const bool up = (lane & (2u * h)) != 0u;
#pragma unroll
        for (int j = 0; j < h; ++j) {
            const double send = up ? v[j] : v[j + h];
            const double keep = up ? v[j + h] : v[j];
            v[j] = keep + shfl_xor_f64(send, 2 * h);
        }
    }
    return v[0] + shfl_xor_f64(v[0], 1);
}

__global__ void __launch_bounds__(TC_THREADS, 1) mv_tc32_frontier_kernel(const __grid_constant__ MatvecArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const unsigned NS = a.n_units;  // stages of the X ring (tc_stages_for(Npad), passed by the launcher)
    const unsigned B = a.B, Npad = a.Bpad;  // Npad: multiple of 16, <= 64
    const unsigned pbytes = Npad * 128;     // one panel part (hi or lo)
    unsigned char* x_ring = smem;
    unsigned char* p_ring = x_ring + (size_t) NS * TC_XBYTES;               // (multiples of 1024 bytes throughout)
    double* red = reinterpret_cast<double*>(p_ring + (size_t) TC_P_STAGES * 2 * pbytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(red + TC_RED_DOUBLES);  // X tile landed
    uint64_t* empty = full + TC_MAX_STAGES;                              // the split team has read the X stage
    uint64_t* pfull = empty + TC_MAX_STAGES;                             // panel of a chunk landed
    uint64_t* pempty = pfull + TC_P_STAGES;                              // ... and every tile's MMAs of that chunk are done
    uint64_t* a_ready = pempty + TC_P_STAGES;                            // X_hi | X_lo of a step written to TMEM (per team)
    uint64_t* a_free = a_ready + TC_A_STAGES;                            // the MMAs that read that stage are done
    uint64_t* acc_full = a_free + TC_A_STAGES;                           // the group's accumulators are complete
    uint64_t* acc_empty = acc_full + 1;                                  // ... and have been read
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
    int* flag = reinterpret_cast<int*>(tmem_slot + 2);
    __shared__ double2 node_cst[kMaxNodesPerPass];  // {log(1/sigma/sqrt(2 pi)), 1/(2 sigma^2)}

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned n_chunks = a.cols_pad / TC_KCH;
    const unsigned GT = TC_A_BASE / Npad;  // accumulator tiles TMEM holds next to the two A stages: tiles per group

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
    // the tiles of one GROUP (as many as TMEM has accumulators for), starting at row g0: returns their number and the
    // first row after them.  Loop order inside a group is CHUNK-major -- for every 32-column chunk the panel is fetched
    // ONCE and all tiles of the group contract against it, each into its own accumulator -- so the panel traffic is
    // 1 / (tiles per group) of the tile-major order's (which re-read 16 KB of panel per 16 KB of X at Npad = 64).
    auto group_tiles = [&](unsigned long long g0, unsigned long long& g1) {
        unsigned n = 0;
        unsigned long long cur = g0, t1;
        while (cur < r1 && n < GT) {
            next_tile(cur, t1);
            cur = t1;
            ++n;
        }
        g1 = cur;
        return n;
    };

    const bool tr0 = tid == 0 && blockIdx.x == 0;
    trace_mark(a.trace, 0, 0, tr0);
    trace_cta(a.trace, 0, tid == 0);
    if (tid == 0) {
        for (unsigned s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 4);
        }
        for (int s = 0; s < TC_P_STAGES; ++s) {
            mbar_init(&pfull[s], 1);
            mbar_init(&pempty[s], 1);
        }
        for (int s = 0; s < TC_A_STAGES; ++s) {
            mbar_init(&a_ready[s], 4);
            mbar_init(&a_free[s], 1);
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
        // ------------------------------------------------------------------ producer (whole warp walks, elected lane issues)
        if (r0 < r1) {
            if (elect_one()) prefetch_tensormap(&a.tmap);
            asm volatile("griddepcontrol.wait;" ::: "memory");  // the panel (prep kernel) is first touched below
            unsigned q = 0, pc = 0, s = 0, sph = 1;  // s = q % NS, sph = parity of q / NS - 1 (no division per step)
            const uint64_t pol_x = l2_policy_evict_first(), pol_p = l2_policy_evict_last();  // X streams, the panel is re-read
            unsigned long long g1;
            for (unsigned long long g0 = r0; g0 < r1; g0 = g1) {
                const unsigned nt = group_tiles(g0, g1);
                for (unsigned c = 0; c < n_chunks; ++c, ++pc) {
                    const unsigned ps = pc % TC_P_STAGES;
                    if (pc >= TC_P_STAGES) mbar_wait(&pempty[ps], ((pc / TC_P_STAGES) - 1) & 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&pfull[ps], 2 * pbytes);  // beta_hi | beta_lo of this chunk are adjacent: one copy
                        bulk_g2s_hint(p_ring + (size_t) ps * 2 * pbytes,
                                      static_cast<const unsigned char*>(a.betaT) + (size_t) c * 2 * pbytes, 2 * pbytes, &pfull[ps], pol_p);
                    }
                    __syncwarp();
                    unsigned long long t0 = g0, t1;
                    for (unsigned i = 0; i < nt; ++i, ++q, t0 = t1) {
                        next_tile(t0, t1);
                        if (q >= NS) mbar_wait(&empty[s], sph);
                        unsigned char* st = x_ring + (size_t) s * TC_XBYTES;
                        const bool two = t1 - t0 > (unsigned long long) TC_BOX_ROWS;
                        if (elect_one()) {
                            TC_TL(0, q);
                            mbar_arrive_expect_tx(&full[s], two ? TC_XBYTES : TC_XBYTES / 2);
                            tma_load_2d_hint(st, &a.tmap, (int) (c * TC_KCH), (int) t0, &full[s], pol_x);
                            if (two)
                                tma_load_2d_hint(st + TC_XBYTES / 2, &a.tmap, (int) (c * TC_KCH), (int) t0 + TC_BOX_ROWS, &full[s], pol_x);
                        }
                        __syncwarp();
                        if (++s == NS) {
                            s = 0;
                            sph ^= 1u;
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ------------------------------------------------------------------ MMA issuer: the WHOLE warp walks the loop
        // (uniform control flow, operands in uniform registers) and an elected lane issues.  With the loop inside `if (lane == 0)`
        // every descriptor reached the tensor-core instructions through a per-thread-to-uniform conversion loop
        // (ELECT / R2UR.BROADCAST / BRA.U.ANY) and the issuing thread, not the tensor core, bounded the step.
        if (r0 < r1) {
            // instruction descriptor: D = F32, A = B = TF32, both K-major, N = Npad, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((Npad >> 3) << 17) | ((TC_ROWS >> 4) << 24);
            const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_acc, 0);
            unsigned q = 0, pc = 0, grp = 0;
            unsigned long long g1;
            for (unsigned long long g0 = r0; g0 < r1; g0 = g1, ++grp) {
                const unsigned nt = group_tiles(g0, g1);
                if (grp > 0) mbar_wait(acc_empty, (grp - 1) & 1);  // the previous group's accumulators have been read
                tc_fence_after();
                for (unsigned c = 0; c < n_chunks; ++c, ++pc) {
                    const unsigned ps = pc % TC_P_STAGES;
                    mbar_wait(&pfull[ps], (pc / TC_P_STAGES) & 1);
                    const uint32_t pb = smem_u32(p_ring + (size_t) ps * 2 * pbytes);
                    const uint64_t b_hi = umma_desc_k_sw128(pb), b_lo = umma_desc_k_sw128(pb + pbytes);
                    for (unsigned i = 0; i < nt; ++i, ++q) {
                        const unsigned at = q % TC_A_STAGES;
                        mbar_wait(&a_ready[at], (q / TC_A_STAGES) & 1);  // X_hi | X_lo of this step are in TMEM
                        tc_fence_after();
                        const uint32_t a_hi = tbase + TC_A_BASE + at * TC_A_COLS, a_lo = a_hi + TC_KCH;
                        const uint32_t acc = tbase + i * Npad;
                        if (elect_one()) {
                            TC_TL(2, q);
#pragma unroll
                            for (unsigned k = 0; k < 4; ++k) {  // 8 floats per k-step: 8 TMEM columns of A, 2 descriptor units of B
                                umma_tf32_ts(acc, a_hi + 8 * k, b_hi + 2 * k, idesc, (c | k) != 0);
                                umma_tf32_ts(acc, a_lo + 8 * k, b_hi + 2 * k, idesc, 1u);
                                umma_tf32_ts(acc, a_hi + 8 * k, b_lo + 2 * k, idesc, 1u);
                            }
                            umma_commit(&a_free[at]);  // the A stage may be rewritten once these MMAs have read it
                            TC_TL(3, q);
                        }
                        __syncwarp();
                    }
                    if (elect_one()) umma_commit(&pempty[ps]);  // ... and the panel stage
                    __syncwarp();
                }
                if (elect_one()) umma_commit(acc_full);
                __syncwarp();
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ split teams (warps 0..3 and 6..9); team 0 is
        // also the epilogue.  A step: read the landed X tile from shared memory ONCE (thread = row), release the stage, store
        // X_hi and X_lo = X - tf32(X) into the team's A stage in TENSOR MEMORY (tcgen05.st).  The 12 MMAs of the step read
        // A from TMEM and only the small coefficient panel from shared memory: no second copy of the tile in shared memory,
        // no proxy fence, and an X stage is held from landing to the split only (not until its MMAs have completed).
        const unsigned team = warp < 4 ? 0u : 1u;
        const unsigned tt = (unsigned) tid & 127u;  // thread within the team (tid 0..127 | 192..319)
        ItemOwner own;
        own.init(r0, a.item_rows);
        unsigned q = team, grp = 0, seg = 0;
        unsigned s = team % NS, sph = 0;  // s = q % NS, sph = parity of q / NS
        unsigned qbase = 0;               // steps of the groups before this one
        unsigned long long g1;
        for (unsigned long long g0 = r0; g0 < r1; g0 = g1, ++grp) {
            const unsigned nt = group_tiles(g0, g1);
            const unsigned qend = qbase + n_chunks * nt;
            for (; q < qend; q += TC_TEAMS) {
                const unsigned u = q / TC_TEAMS;  // uses of this team's A stage so far
                mbar_wait(&full[s], sph);
                if (tt == 0) TC_TL(1, q);
                if (q == 0) trace_mark(a.trace, 0, 2, tr0);  // first tile landed
                // thread = row of the tile (TMEM lane 32 (warp % 4) + lane): its 128 bytes sit in 8 swizzled 16-byte vectors
                const unsigned row = (unsigned) (warp & 3) * 32u + (unsigned) lane;
                const unsigned row_s = smem_u32(x_ring + (size_t) s * TC_XBYTES) + row * 128u;
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int v = 0; v < 8; ++v)  // every load in flight before the first use; a quarter warp covers all banks
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(hi[4 * v]), "=r"(hi[4 * v + 1]), "=r"(hi[4 * v + 2]), "=r"(hi[4 * v + 3])
                                 : "r"(row_s + (((unsigned) v ^ (row & 7u)) << 4)));
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    lo[j] = __float_as_uint(tf32_pre_round(__uint_as_float(hi[j]) - tf32_trunc(__uint_as_float(hi[j]))));
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);  // the X stage may be refilled: its words are in registers
                // only now the A stage is needed: the wait for the MMAs of step q - 2 comes AFTER the loads and the split
                // arithmetic, so the chain  MMAs done -> A rewritten -> next MMAs  is two TMEM stores long
                if (u >= 1) {
                    mbar_wait(&a_free[team], (u - 1) & 1);
                    tc_fence_after();
                }
                const uint32_t a_t = tmem_acc + ((uint32_t) ((warp & 3) * 32) << 16) + TC_A_BASE + team * TC_A_COLS;
                tmem_st32(a_t, hi);            // the tensor core truncates the FP32 words to tf32 itself
                tmem_st32(a_t + TC_KCH, lo);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[team]);
                s += TC_TEAMS;
                if (s >= NS) {
                    s -= NS;
                    sph ^= 1u;
                }
            }
            qbase = qend;
            if (team != 0) continue;
            // ---- epilogue of the group (team 0): thread = row of a tile.  The rows of one ITEM SEGMENT (the part of an item
            // inside this group: at C4 the whole group, or two pieces) are summed per thread over the tiles first, in FP64,
            // and reduced across the lanes once per segment and 16 nodes (warp_reduce16) -- not once per tile and node.
            mbar_wait(acc_full, grp & 1);
            tc_fence_after();
            for (unsigned long long s0 = g0, s1; s0 < g1; s0 = s1, ++seg) {
                s1 = (s0 / a.item_rows + 1) * a.item_rows;
                if (s1 > g1) s1 = g1;
                double* rbuf = red + (seg & 1) * (4 * kMaxNodesPerPass);
                for (unsigned n0 = 0; n0 < Npad; n0 += 16) {
                    double acc[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] = 0.0;
                    unsigned long long t0 = g0, t1;
                    for (unsigned i = 0; i < nt; ++i, t0 = t1) {
                        next_tile(t0, t1);
                        if (t1 <= s0 || t0 >= s1) continue;  // (uniform) tile outside the segment
                        const unsigned long long r = t0 + tt;
                        const bool live = r < t1 && r >= s0 && r < s1;
                        const double yv = live ? __ldg(&a.y[r]) : 0.0;
                        uint32_t d[16];
                        tmem_ld16(tmem_acc + ((uint32_t) (warp * 32) << 16) + i * Npad + n0, d);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const double2 ch = node_cst[n0 + j];
                            const double res = (double) __uint_as_float(d[j]) - yv;
                            acc[j] += live ? ch.x - res * res * ch.y : 0.0;
                        }
                    }
                    const double tot = warp_reduce16(acc, (unsigned) lane);
                    if ((lane & 1) == 0) rbuf[warp * kMaxNodesPerPass + n0 + warp_reduce16_index((unsigned) lane)] = tot;
                }
                named_bar_sync(1, TC_WORKERS);
                if ((unsigned) tid < B) {
                    double tot = 0.0;
                    for (unsigned w = 0; w < 4; ++w) tot += rbuf[w * kMaxNodesPerPass + tid];
                    own.add_tile(tot, s1, a.item_rows, a.row_end);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);  // the accumulators may be overwritten by the next group
        }
        trace_mark(a.trace, 0, 4, tr0);  // all tiles done
        const unsigned ftid = team ? (unsigned) tid - 64u : (unsigned) tid;  // 0..255 over both teams
        if (team == 0) {
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
        }
        named_bar_sync(2, TC_TEAMS * TC_WORKERS);
        trace_mark(a.trace, 0, 5, tr0);
        trace_cta(a.trace, 1, tid == 0);
        if (flag[0]) {  // last CTA: finalize (item stitching, exchange between GPUs, delivery), both teams
            trace_mark(a.trace, 1, 0, tid == 0);
            __threadfence();
            double S, Q;
            // every TMA load of this CTA has landed and been consumed: the X ring is scratch for the stitch walks
            if (gridDim.x <= 160 && B > 2 * TC_TEAMS * 4)  // (at most 148 CTAs here) more than two nodes per warp: four at a time
                finalize_items<true, 5, 4>(a.part, gridDim.x, a.row_begin, a.row_end, a.rows_per_cta, a.item_rows, ftid, B, red, 2,
                                           TC_TEAMS * TC_WORKERS, S, Q, reinterpret_cast<double*>(x_ring),
                                           (size_t) NS * TC_XBYTES / sizeof(double), a.trace);
            else
                finalize_items(a.part, gridDim.x, a.row_begin, a.row_end, a.rows_per_cta, a.item_rows, ftid, B, red, 2,
                               TC_TEAMS * TC_WORKERS, S, Q, reinterpret_cast<double*>(x_ring),
                               (size_t) NS * TC_XBYTES / sizeof(double));
            trace_mark(a.trace, 1, 1, tid == 0);
            unsigned Bout = B;
            if (a.xchg.nranks > 1) {
                const double2 t = peer_allreduce(a.xchg, ftid, B, TC_TEAMS * TC_WORKERS, 2, red, S, Q);
                S = t.x;
                Q = t.y;
                if (a.xchg.gather_per) Bout = a.xchg.gather_total;
            }
            if (ftid < Bout) {
                if (a.done_flag) {
                    store_result_words(a.out, ftid, S, a.done_seq);
                    store_result_words(a.out, Bout + ftid, Q, a.done_seq);
                } else {
                    a.out[ftid] = S;
                    a.out[Bout + ftid] = Q;
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
    // contiguous rows per CTA, a whole number of 64-row TMA boxes (no row fetched twice; C4: 143 CTAs x 704 rows)
    const unsigned long long units = (rows + TC_BOX_ROWS - 1) / TC_BOX_ROWS;
    unsigned long long g = units < (unsigned long long) e.sm_count ? units : (unsigned long long) e.sm_count;
    if (g < 1) g = 1;
    const unsigned long long upc = (units + g - 1) / g;
    a.rows_per_cta = upc * TC_BOX_ROWS;
    a.k_splits = 1;
    a.chunks_per_split = a.cols_pad / TC_KCH;
    a.n_units = tc_stages_for(a.Bpad);  // stages of the X ring
    const size_t smem_bytes = kTcSmemFixed + (size_t) TC_P_STAGES * tc_panel_bytes(a.Bpad) + (size_t) a.n_units * TC_XBYTES;
    unsigned grid = (unsigned) ((rows + a.rows_per_cta - 1) / a.rows_per_cta);
    if (grid < 1) grid = 1;
    if (make_tmap(e, TC_BOX_ROWS, &a.tmap) != 0) return -1;
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
