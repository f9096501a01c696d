// pf_matvec.cu -- frontier kernels for the models whose row term needs a dot product of the
// data row with a per-node vector:
//
//   Bayesian lasso   pyblasso.py:135-141   lp_item = sum_rows [ log(1/sigma/sqrt(2pi)) - (beta.x_r - y_r)^2 / 2 / sigma^2 ]
//   dense Boltzmann  (new workload)        log p   = -x'Wx / T  = -(1/T) sum_i x_i (W_i . x)
//
// Both stream a row-major matrix (X: n x p, W: D x D) exactly ONCE per launch and multiply it
// by the B node vectors at the same time -- a tall-skinny GEMM fused with its row-wise epilogue
// and the reduction over rows, so the n x B product never exists in memory.
//
// DESIGN (B200).  A producer warp walks (row tile x 128-byte column chunk) boxes of the matrix
// with 2-D TMA (cp.async.bulk.tensor, SWIZZLE_128B) into a 4-stage shared-memory ring, together
// with the matching [chunk][B] slice of the coefficient panel (1-D bulk copy; the panel is the
// thetas transposed once per call by a tiny prep kernel and stays L2 resident).  Eight consumer
// warps are arranged ROW-WARPS x NODE-GROUPS; a lane owns RPL rows x 8 nodes of accumulators.
// Row operands are read with conflict-free LDS.128 thanks to the swizzle, coefficient operands
// are warp-wide broadcasts, and every operand fetched from shared memory feeds 8 (row) or
// RPL (coefficient) FP64 FMAs.  The epilogue (residual, square, scale / multiply by x_i) runs
// in registers; item totals are assembled exactly like in pf_pointwise.cu.
//
// Arithmetic is IEEE double like the reference (sm_100a has no FP64 tcgen05 kind).  Two consumer variants:
//   * vector (NBK = 0): FMA on the CUDA cores as described above -- PF_F32 (X/W and the panel stored as float,
//     float accumulation) and, on request (PF_OPT_MATVEC_TENSOR = 0), FP64;
//   * tensor (NBK = 1, 2, 4, 8 blocks of 8 nodes; the FP64 default): mma.sync.m8n8k4.f64 (SASS DMMA.884) -- the FP64 tensor-core
//     instruction.  One DMMA is 256 FMAs from 4 register pairs, so the contraction is no longer limited by
//     register-file bandwidth (the 3-operand DFMA form tops out at 43 of 64 FMA/clk/SM, scripts/fp64_probe.py,
//     while DMMA issues at 63.8).  A warp owns 32 rows x all node blocks; A fragments come straight out of the
//     swizzled TMA stage (conflict-free LDS.64), B fragments from a panel the prep kernel wrote in fragment
//     order (lane-contiguous LDS.64).  Products and sums are exact IEEE FP64 FMAs, only the summation order
//     over k differs from the vector variant.
#include <math.h>
#include <stdio.h>

#include "pf_common.cuh"
#include "pf_engine.h"

namespace pf {

constexpr int MV_CONSUMER_WARPS = 8;
constexpr int MV_CONSUMERS = MV_CONSUMER_WARPS * 32;
constexpr int MV_THREADS = MV_CONSUMERS + 32;
constexpr int MV_STAGES = 4;
constexpr int MV_CHUNK_BYTES = 128;  // bytes of one matrix row per stage == the swizzle span
constexpr int MV_MAX_TILE_ROWS = 256;
constexpr int MV_NODES_PER_LANE = 8;
constexpr int MV_XSTAGE_BYTES = MV_MAX_TILE_ROWS * MV_CHUNK_BYTES;                  // 32 KB
constexpr int MV_BSTAGE_BYTES = MV_CHUNK_BYTES * kMaxNodesPerPass;                  // [chunk k][64 nodes] = 8 KB
constexpr int MV_RED_DOUBLES = 2 * 2 * MV_CONSUMER_WARPS * kMaxNodesPerPass;  // 2 buffers x (part A | part B) x warps x nodes

constexpr size_t kMvSmem = 1024 /* alignment slack */ + (size_t) MV_STAGES * (MV_XSTAGE_BYTES + MV_BSTAGE_BYTES) +
                           MV_RED_DOUBLES * sizeof(double) + 2 * MV_STAGES * sizeof(uint64_t) + 16;

// transpose thetas into the coefficient panel: panel[k][b] = theta[b][off + k], zero padded
template <typename T>
__global__ void mv_prep_panel(const double* theta, T* panel, unsigned B, unsigned Bpad, unsigned theta_len,
                              unsigned off, unsigned cols, unsigned cols_pad) {
    asm volatile("griddepcontrol.launch_dependents;");  // the frontier kernel may start its prologue now
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cols_pad * Bpad) return;
    const unsigned k = idx / Bpad, b = idx % Bpad;
    double v = 0.0;
    if (k < cols && b < B) v = theta[(size_t) b * theta_len + off + k];
    panel[idx] = (T) v;
}

// The same panel in DMMA fragment order: for 16-column chunk c, k-step s (4 columns), node block nb (8 nodes),
// lane l holds B[k = l % 4][n = l / 4]:  panel[((c * 4 + s) * NBK + nb) * 32 + l] = theta[nb * 8 + l / 4][off + 16 c + 4 s + l % 4]
__global__ void mv_prep_panel_frag(const double* theta, double* panel, unsigned B, unsigned NBK, unsigned theta_len,
                                   unsigned off, unsigned cols, unsigned cols_pad) {
    asm volatile("griddepcontrol.launch_dependents;");
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cols_pad * NBK * 8) return;
    const unsigned l = idx & 31, nb = (idx >> 5) % NBK, ks = (idx >> 5) / NBK;
    const unsigned k = ks * 4 + (l & 3), b = nb * 8 + (l >> 2);
    double v = 0.0;
    if (k < cols && b < B) v = theta[(size_t) b * theta_len + off + k];
    panel[idx] = v;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Work units.  Lasso: the CTA's contiguous row range cut into tiles that never straddle an item,
// each over ALL column chunks (the epilogue is non-linear in the dot product).  Boltzmann: units
// are (row tile, column split) pairs dealt round-robin to CTAs (the epilogue is linear).
template <bool BOLTZ>
struct UnitIter {
    // lasso
    unsigned long long cur, end_all, item_rows, row_end;
    // boltzmann
    unsigned u, n_units, stride, k_splits, chunks_per_split, n_chunks;
    unsigned tile_rows;
    unsigned long long row_begin;
    __device__ __forceinline__ bool next(unsigned long long& t0, unsigned long long& t1, unsigned& c0, unsigned& c1) {
        if constexpr (BOLTZ) {
            if (u >= n_units) return false;
            const unsigned tile = u / k_splits, split = u % k_splits;
            t0 = row_begin + (unsigned long long) tile * tile_rows;
            t1 = t0 + tile_rows;
            if (t1 > row_end) t1 = row_end;
            c0 = split * chunks_per_split;
            c1 = c0 + chunks_per_split;
            if (c1 > n_chunks) c1 = n_chunks;
            u += stride;
            return true;
        } else {
            if (cur >= end_all) return false;
            t0 = cur;
            unsigned long long e = cur + tile_rows;
            if (e > end_all) e = end_all;
            if (item_rows < tile_rows) {  // small items: a tile never straddles one
                unsigned long long ie = (cur / item_rows + 1) * item_rows;
                if (ie > row_end) ie = row_end;
                if (e > ie) e = ie;
            }
            t1 = e;
            cur = e;
            c0 = 0;
            c1 = n_chunks;
            return true;
        }
    }
};

template <typename T, int RPL, bool BOLTZ, int NBK, int MB = 4>
__global__ void __launch_bounds__(MV_THREADS, 1) mv_frontier_kernel(const __grid_constant__ MatvecArgs a) {
    using CT = typename std::conditional<sizeof(T) == 4, float, double>::type;
    // tensor variant: NBK blocks of 8 nodes and MB blocks of 8 rows per warp (tile = 64 MB rows).  MB = 4 for the
    // lasso; the small dense Boltzmann matrix takes MB = 1: its epilogue (one x_i load per row x node) is repeated
    // for every column split, so short tiles x few splits make it 4x cheaper (3.9 -> ~1 us of a 16 us call)
    constexpr bool MMA = NBK > 0;
    static_assert(!MMA || sizeof(T) == 8, "DMMA is the FP64 path");
    constexpr int KPV = 16 / (int) sizeof(T);             // k values per 16-byte vector
    constexpr int KCH = MV_CHUNK_BYTES / (int) sizeof(T);  // k values per chunk

    // 1024-byte alignment (the 128-byte swizzle pattern repeats every 8 rows = 1024 B).  The pad is ADDED
    // to the array so the compiler still knows these are shared-memory addresses (LDS, not generic LD).
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* xstage = smem;
    unsigned char* bstage = smem + MV_STAGES * MV_XSTAGE_BYTES;
    double* red = reinterpret_cast<double*>(bstage + MV_STAGES * MV_BSTAGE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(red + MV_RED_DOUBLES);
    uint64_t* empty = full + MV_STAGES;
    int* flag = reinterpret_cast<int*>(empty + MV_STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned B = a.B, Bpad = a.Bpad;
    const unsigned NG = Bpad / MV_NODES_PER_LANE;  // node groups: 1, 2, 4 or 8
    const unsigned tile_rows = a.tile_rows;        // = (8 / NG) * 32 * RPL
    const unsigned n_chunks = a.cols_pad / KCH;
    const unsigned xbytes = tile_rows * MV_CHUNK_BYTES;
    const unsigned bbytes = KCH * Bpad * (unsigned) sizeof(T);

    unsigned long long r0 = a.row_begin, r1 = a.row_end;
    if constexpr (!BOLTZ) {
        const unsigned long long per = a.rows_per_cta;
        r0 = a.row_begin + (unsigned long long) blockIdx.x * per;
        if (r0 > a.row_end) r0 = a.row_end;
        r1 = r0 + per;
        if (r1 > a.row_end) r1 = a.row_end;
    }
    UnitIter<BOLTZ> proto{r0, r1, a.item_rows, a.row_end, blockIdx.x, a.n_units, gridDim.x, a.k_splits,
                          a.chunks_per_split, n_chunks, tile_rows, a.row_begin};

    const bool tr0 = tid == 0 && blockIdx.x == 0;
    trace_mark(a.trace, 0, 0, tr0);
    trace_cta(a.trace, 0, tid == 0);
    if (tid == 0) {
        for (int s = 0; s < MV_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], MV_CONSUMER_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == MV_CONSUMER_WARPS) {
        // ------------------------------------------------------------------ producer warp
        if (lane == 0) {
            prefetch_tensormap(&a.tmap);
            // programmatic dependent launch: everything above overlapped the panel prep kernel; its output
            // (a.betaT) is first touched below
            asm volatile("griddepcontrol.wait;" ::: "memory");
            UnitIter<BOLTZ> it = proto;
            unsigned long long t0, t1;
            unsigned c0, c1, q = 0;
            [[maybe_unused]] const uint64_t pol_x = l2_policy_evict_first(), pol_p = l2_policy_evict_last();
            while (it.next(t0, t1, c0, c1)) {
                for (unsigned c = c0; c < c1; ++c, ++q) {
                    const unsigned s = q % MV_STAGES;
                    if (q >= MV_STAGES) mbar_wait(&empty[s], ((q / MV_STAGES) - 1) & 1);
                    mbar_arrive_expect_tx(&full[s], xbytes + bbytes);
                    if constexpr (BOLTZ) {  // W (8 MB) stays in L2 between calls: default policy
                        tma_load_2d(xstage + s * MV_XSTAGE_BYTES, &a.tmap, (int) (c * KCH), (int) t0, &full[s]);
                        bulk_g2s(bstage + s * MV_BSTAGE_BYTES,
                                 static_cast<const unsigned char*>(a.betaT) + (size_t) c * bbytes, bbytes, &full[s]);
                    } else {  // the matrix streams through L2 once per call; the panel and the CTA records must survive it
                        tma_load_2d_hint(xstage + s * MV_XSTAGE_BYTES, &a.tmap, (int) (c * KCH), (int) t0, &full[s], pol_x);
                        bulk_g2s_hint(bstage + s * MV_BSTAGE_BYTES,
                                      static_cast<const unsigned char*>(a.betaT) + (size_t) c * bbytes, bbytes, &full[s], pol_p);
                    }
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const unsigned g = warp % NG, rw = warp / NG;  // node group, row warp
    const unsigned row_in_tile = rw * (32 * RPL) + lane;  // + 32*i
    const unsigned sw = lane & 7;                          // swizzle phase of all my rows

    // per-node constants of the epilogue (tensor variant: in shared memory, a lane meets 2 * NBK nodes)
    __shared__ double2 node_cst[MMA ? kMaxNodesPerPass : 1];  // {log(1/sigma/sqrt(2 pi)), 1/(2 sigma^2)}
    [[maybe_unused]] double cst[MMA ? 1 : MV_NODES_PER_LANE], hinv[MMA ? 1 : MV_NODES_PER_LANE];
    if constexpr (MMA) {
        if (!BOLTZ && tid < kMaxNodesPerPass) {
            const unsigned node = (unsigned) tid < B ? tid : B - 1;
            const double sigma = a.theta[(size_t) node * a.theta_len];
            node_cst[tid] = make_double2(log(1 / sigma / sqrt(2 * 3.14159265358979323846)),  // pyblasso.py:140
                                         1.0 / (2.0 * sigma * sigma));
        }
        named_bar_sync(1, MV_CONSUMERS);
    } else {
#pragma unroll
        for (int n = 0; n < MV_NODES_PER_LANE; ++n) {
            unsigned node = g * MV_NODES_PER_LANE + n;
            if (node >= B) node = B - 1;
            if constexpr (BOLTZ) {
                cst[n] = 0.0;
                hinv[n] = 0.0;
            } else {
                const double sigma = a.theta[(size_t) node * a.theta_len];
                cst[n] = log(1 / sigma / sqrt(2 * 3.14159265358979323846));  // pyblasso.py:140
                hinv[n] = 1.0 / (2.0 * sigma * sigma);
            }
        }
    }

    ItemOwner own;
    own.init(r0, BOLTZ ? 1ull : a.item_rows);

    UnitIter<BOLTZ> it = proto;
    unsigned long long t0, t1;
    unsigned c0, c1, q = 0, unit = 0;
    while (it.next(t0, t1, c0, c1)) {
        CT acc[MMA ? 1 : RPL][MV_NODES_PER_LANE];
        [[maybe_unused]] double cacc[MMA ? MB : 1][MMA ? NBK : 1][2];  // C fragments: row lane/4, nodes 2*(lane%4) + {0,1}
        if constexpr (MMA) {
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int nb = 0; nb < NBK; ++nb) cacc[mb][nb][0] = cacc[mb][nb][1] = 0.0;
        } else {
#pragma unroll
            for (int i = 0; i < RPL; ++i)
#pragma unroll
                for (int n = 0; n < MV_NODES_PER_LANE; ++n) acc[i][n] = 0;
        }

        for (unsigned c = c0; c < c1; ++c, ++q) {
            const unsigned s = q % MV_STAGES;
            mbar_wait(&full[s], (q / MV_STAGES) & 1);
            if (q == 0) trace_mark(a.trace, 0, 2, tr0);  // first box landed
            if constexpr (MMA) {
                // A[row = lane/4][k = lane%4] of row block mb, k-step ks: 16-byte chunk (2 ks + (lane%4)/2) of the
                // row, swizzled with row & 7 == lane / 4 (row blocks start at multiples of 8), half lane & 1
                const unsigned char* xw = xstage + s * MV_XSTAGE_BYTES +
                                          (size_t) (warp * (8 * MB) + (lane >> 2)) * MV_CHUNK_BYTES + ((lane & 1) << 3);
                const unsigned c0x = (((lane >> 1) & 1) ^ (lane >> 2)) << 4;
                const double* bw = reinterpret_cast<const double*>(bstage + s * MV_BSTAGE_BYTES) + lane;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    double bf[NBK];
#pragma unroll
                    for (int nb = 0; nb < NBK; ++nb) bf[nb] = bw[(ks * NBK + nb) * 32];
                    // row blocks past the end of a short tile hold stale rows: they only feed their own (masked)
                    // accumulator rows, and an unconditional instruction stream keeps every load ahead of its DMMAs
                    double af[MB];
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb)
                        af[mb] = *reinterpret_cast<const double*>(xw + mb * 8 * MV_CHUNK_BYTES + (c0x ^ (unsigned) (ks << 5)));
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                        for (int nb = 0; nb < NBK; ++nb) dmma884(cacc[mb][nb][0], cacc[mb][nb][1], af[mb], bf[nb]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                continue;
            }
            const unsigned char* xs = xstage + s * MV_XSTAGE_BYTES + (size_t) row_in_tile * MV_CHUNK_BYTES;
            const T* bs = reinterpret_cast<const T*>(bstage + s * MV_BSTAGE_BYTES) + g * MV_NODES_PER_LANE;
#pragma unroll
            for (int v = 0; v < MV_CHUNK_BYTES / 16; ++v) {
                T xv[RPL][KPV];
#pragma unroll
                for (int i = 0; i < RPL; ++i)
                    *reinterpret_cast<int4*>(xv[i]) =
                        *reinterpret_cast<const int4*>(xs + (size_t) i * 32 * MV_CHUNK_BYTES + ((v ^ sw) << 4));
#pragma unroll
                for (int kk = 0; kk < KPV; ++kk) {
                    T bv[MV_NODES_PER_LANE];
                    const T* bp = bs + (size_t) (v * KPV + kk) * Bpad;
#pragma unroll
                    for (int w = 0; w < MV_NODES_PER_LANE * (int) sizeof(T) / 16; ++w)
                        reinterpret_cast<int4*>(bv)[w] = reinterpret_cast<const int4*>(bp)[w];
#pragma unroll
                    for (int i = 0; i < RPL; ++i)
#pragma unroll
                        for (int n = 0; n < MV_NODES_PER_LANE; ++n)
                            acc[i][n] = fma((CT) xv[i][kk], (CT) bv[n], acc[i][n]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }

        if (unit == 0) trace_mark(a.trace, 0, 3, tr0);  // first unit's contraction done
        // ---- epilogue for this unit, in registers.  A lasso tile may contain ONE item boundary `ts`
        // (items are at least a tile tall in that mode): rows before it go to partA, rows after to partB.
        unsigned long long ts = t1;
        if constexpr (!BOLTZ) {
            const unsigned long long ie = (t0 / a.item_rows + 1) * a.item_rows;
            if (ie < t1) ts = ie;
        }
        const bool split = ts < t1;
        double* rbuf = red + (unit & 1) * (2 * MV_CONSUMER_WARPS * kMaxNodesPerPass);
        if constexpr (MMA) {
            // my rows: block mb, row lane/4; my nodes: block nb, 2*(lane%4) + {0,1}
            double yv[MB];
            bool live[MB], first[MB];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                const unsigned long long r = t0 + warp * (8 * MB) + mb * 8 + (lane >> 2);
                live[mb] = r < t1;
                first[mb] = r < ts;
                yv[mb] = 0.0;
                if constexpr (!BOLTZ)
                    if (live[mb]) yv[mb] = __ldg(&a.y[r]);
            }
#pragma unroll
            for (int nb = 0; nb < NBK; ++nb) {
                const unsigned node0 = nb * 8 + 2 * (lane & 3);
                double pA[2] = {0.0, 0.0}, pB[2] = {0.0, 0.0};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const unsigned node = node0 + h < B ? node0 + h : B - 1;
                    [[maybe_unused]] double2 ch = make_double2(0.0, 0.0);
                    if constexpr (!BOLTZ) ch = node_cst[node];
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb) {
                        if (!live[mb]) continue;
                        if constexpr (BOLTZ) {
                            const unsigned long long r = t0 + warp * (8 * MB) + mb * 8 + (lane >> 2);
                            const double xi = __ldg(&a.theta[(size_t) node * a.theta_len + r]);
                            pA[h] = fma(xi, cacc[mb][nb][h], pA[h]);
                        } else {
                            const double res = cacc[mb][nb][h] - yv[mb];
                            const double lp = ch.x - res * res * ch.y;
                            pA[h] += first[mb] ? lp : 0.0;
                            pB[h] += first[mb] ? 0.0 : lp;
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    for (int m = 16; m >= 4; m >>= 1) {
                        pA[h] += shfl_xor_f64(pA[h], m);
                        pB[h] += shfl_xor_f64(pB[h], m);
                    }
                    if (lane < 4) {
                        rbuf[warp * kMaxNodesPerPass + node0 + h] = pA[h];
                        rbuf[(MV_CONSUMER_WARPS + warp) * kMaxNodesPerPass + node0 + h] = pB[h];
                    }
                }
            }
        }
        double partA[MV_NODES_PER_LANE], partB[MV_NODES_PER_LANE];
#pragma unroll
        for (int n = 0; n < MV_NODES_PER_LANE; ++n) partA[n] = partB[n] = 0.0;
#pragma unroll
        for (int i = 0; i < (MMA ? 0 : RPL); ++i) {
            const unsigned long long r = t0 + row_in_tile + 32 * i;
            if (r < t1) {
                if constexpr (BOLTZ) {
#pragma unroll
                    for (int n = 0; n < MV_NODES_PER_LANE; ++n) {
                        unsigned node = g * MV_NODES_PER_LANE + n;
                        if (node >= B) node = B - 1;
                        const double xi = __ldg(&a.theta[(size_t) node * a.theta_len + r]);
                        partA[n] = fma(xi, (double) acc[i][n], partA[n]);
                    }
                } else {
                    const double yv = __ldg(&a.y[r]);
                    const bool first = r < ts;
#pragma unroll
                    for (int n = 0; n < MV_NODES_PER_LANE; ++n) {
                        const double res = (double) acc[i][n] - yv;
                        const double lp = cst[n] - res * res * hinv[n];
                        partA[n] += first ? lp : 0.0;
                        partB[n] += first ? 0.0 : lp;
                    }
                }
            }
        }
        if constexpr (!MMA) {
#pragma unroll
            for (int n = 0; n < MV_NODES_PER_LANE; ++n) {
                double v = partA[n];
                for (int m = 16; m >= 1; m >>= 1) v += shfl_xor_f64(v, m);
                if (lane == 0) rbuf[rw * kMaxNodesPerPass + g * MV_NODES_PER_LANE + n] = v;
            }
        }
        if (!MMA && split) {
#pragma unroll
            for (int n = 0; n < MV_NODES_PER_LANE; ++n) {
                double v = partB[n];
                for (int m = 16; m >= 1; m >>= 1) v += shfl_xor_f64(v, m);
                if (lane == 0)
                    rbuf[(MV_CONSUMER_WARPS + rw) * kMaxNodesPerPass + g * MV_NODES_PER_LANE + n] = v;
            }
        }
        named_bar_sync(1, MV_CONSUMERS);
        if ((unsigned) tid < B) {
            const unsigned RW = MMA ? MV_CONSUMER_WARPS : MV_CONSUMER_WARPS / NG;
            double tot = 0.0;
            for (unsigned w = 0; w < RW; ++w) tot += rbuf[w * kMaxNodesPerPass + tid];
            if constexpr (BOLTZ)
                own.S += tot;
            else {
                own.add_tile(tot, ts, a.item_rows, a.row_end);
                if (split) {
                    double totB = 0.0;
                    for (unsigned w = 0; w < RW; ++w) totB += rbuf[(MV_CONSUMER_WARPS + w) * kMaxNodesPerPass + tid];
                    own.add_tile(totB, t1, a.item_rows, a.row_end);
                }
            }
        }
        ++unit;
    }

    trace_mark(a.trace, 0, 4, tr0);  // all units done (epilogues included)
    double* rec = a.part + blockIdx.x;  // column of this CTA in the [slot][node][CTA] records
    if ((unsigned) tid < B) {
        if constexpr (!BOLTZ) own.finish(r0, r1, a.row_begin, a.row_end, a.rows_per_cta, a.item_rows);
        own.store(rec, tid);
        __threadfence();
    }
    named_bar_sync(1, MV_CONSUMERS);
    if (tid == 0) {
        const unsigned prev = atomicAdd(a.ticket, 1u);
        flag[0] = (prev == gridDim.x - 1);
    }
    named_bar_sync(1, MV_CONSUMERS);
    trace_mark(a.trace, 0, 5, tr0);  // record published
    trace_cta(a.trace, 1, tid == 0);
    if (!flag[0]) return;

    trace_mark(a.trace, 1, 0, tid == 0);
    __threadfence();
    {
        // Boltzmann: the records hold plain energy partials (one item): finalize with item_rows = 1
        double S, Q;
        // (every load of this CTA has landed and been consumed: the X stages are scratch for the stitch walks)
        if (gridDim.x <= 160 && B > 2 * MV_CONSUMER_WARPS)  // more than two nodes per warp: four at a time
            finalize_items<true, 5, 4>(a.part, gridDim.x, a.row_begin, a.row_end, BOLTZ ? 0ull : a.rows_per_cta,
                                       BOLTZ ? 1ull : a.item_rows, tid, B, red, 1, MV_CONSUMERS, S, Q, reinterpret_cast<double*>(smem),
                                       (size_t) MV_STAGES * MV_XSTAGE_BYTES / sizeof(double));
        else
            finalize_items(a.part, gridDim.x, a.row_begin, a.row_end, BOLTZ ? 0ull : a.rows_per_cta,
                           BOLTZ ? 1ull : a.item_rows, tid, B, red, 1, MV_CONSUMERS, S, Q, reinterpret_cast<double*>(smem),
                           (size_t) MV_STAGES * MV_XSTAGE_BYTES / sizeof(double));
        trace_mark(a.trace, 1, 1, tid == 0);  // records folded
        unsigned Bout = B;  // node split: the gathered frontier is wider than this rank's slice
        if (a.xchg.nranks > 1) {
            const double2 t = peer_allreduce(a.xchg, tid, B, MV_CONSUMERS, 1, red, S, Q);
            S = t.x;
            Q = t.y;
            if (a.xchg.gather_per) Bout = a.xchg.gather_total;
        }
        if ((unsigned) tid < Bout) {
            if constexpr (BOLTZ) {
                const double lp = -S * a.inv_temperature;  // one item: the whole energy
                S = lp;
                Q = lp * lp;
            }
            if (a.done_flag) {  // host path: self-validating words into mapped host memory, no fence, no flag
                store_result_words(a.out, tid, S, a.done_seq);
                store_result_words(a.out, Bout + tid, Q, a.done_seq);
            } else {
                a.out[tid] = S;
                a.out[Bout + tid] = Q;
            }
        }
    }
    if (tid == 0) *a.ticket = 0u;
    trace_mark(a.trace, 1, 2, tid == 0);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace

int make_tmap(Engine& e, unsigned tile_rows, CUtensorMap* out) {
    for (int i = 0; i < e.n_tmaps; ++i)
        if (e.tmap_rows[i] == tile_rows) {
            *out = e.tmaps[i];
            return 0;
        }
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return -1;
    }
    const size_t esz = e.f32 ? 4 : 8;
    const cuuint64_t dims[2] = {(cuuint64_t) e.dim, (cuuint64_t) e.n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t) e.pitch_bytes};
    const cuuint32_t box[2] = {(cuuint32_t) (MV_CHUNK_BYTES / esz), (cuuint32_t) tile_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, e.f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, e.d_data, dims,
                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d)", (int) r);
        return -1;
    }
    if (e.n_tmaps < 8) {
        e.tmaps[e.n_tmaps] = *out;
        e.tmap_rows[e.n_tmaps++] = tile_rows;
    }
    return 0;
}

namespace {

template <typename T, int RPL, bool BOLTZ, int NBK = 0, int MB = 4>
int launch_mv(Engine& e, MatvecArgs& a, cudaStream_t s) {
    auto kern = mv_frontier_kernel<T, RPL, BOLTZ, NBK, MB>;
    static bool configured_dev[64] = {};
    if (!configured_dev[e.device & 63]) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMvSmem);
        if (err == cudaSuccess)
            err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) {
            set_error("cudaFuncSetAttribute(matvec): %s", cudaGetErrorString(err));
            return -1;
        }
        configured_dev[e.device & 63] = true;
    }
    const unsigned long long rows = a.row_end - a.row_begin;
    const unsigned KCH = MV_CHUNK_BYTES / sizeof(T);
    const unsigned n_chunks = a.cols_pad / KCH;
    unsigned grid;
    if (BOLTZ) {
        // enough (row tile, column split) units to occupy every SM
        const unsigned long long tiles = (rows + a.tile_rows - 1) / a.tile_rows;
        // one unit per CTA when the matrix is small (a second round would double the latency-bound run time)
        unsigned splits = tiles < (unsigned long long) e.sm_count ? (unsigned) ((unsigned long long) e.sm_count / tiles) : 1;
        if (splits > n_chunks) splits = n_chunks;
        if (splits < 1) splits = 1;
        a.chunks_per_split = (n_chunks + splits - 1) / splits;
        a.k_splits = (n_chunks + a.chunks_per_split - 1) / a.chunks_per_split;
        a.n_units = (unsigned) tiles * a.k_splits;
        a.rows_per_cta = rows;
        grid = a.n_units < (unsigned) e.sm_count ? a.n_units : (unsigned) e.sm_count;
    } else {
        // Contiguous rows per CTA, cut into equal tiles no taller than the lane layout allows, so
        // every CTA streams the same number of full boxes and no box reads rows it will not use
        // (HBM traffic == algorithmic bytes, except at item boundaries).
        const unsigned max_tr = a.tile_rows;
        unsigned long long g = (rows + 63) / 64;
        if (g > (unsigned long long) e.sm_count) g = e.sm_count;
        if (g < 1) g = 1;
        unsigned long long per = (rows + g - 1) / g;
        const unsigned long long tpc = (per + max_tr - 1) / max_tr;
        unsigned long long tr = (per + tpc - 1) / tpc;
        tr = (tr + 7) & ~7ull;
        if (tr > max_tr) tr = max_tr;
        per = tpc * tr;
        a.tile_rows = (unsigned) tr;
        a.rows_per_cta = per;
        a.k_splits = 1;
        a.chunks_per_split = n_chunks;
        a.n_units = 0;
        grid = (unsigned) ((rows + per - 1) / per);
    }
    if (make_tmap(e, a.tile_rows, &a.tmap) != 0) return -1;
    if (grid < 1) grid = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(MV_THREADS);
    cfg.dynamicSmemBytes = kMvSmem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // start under the tail of mv_prep_panel*
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t err = cudaLaunchKernelEx(&cfg, kern, a);
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("matvec kernel launch: %s", cudaGetErrorString(err));
        return -1;
    }
    return 1;
}

template <typename T, bool BOLTZ>
int dispatch_rpl(Engine& e, MatvecArgs& a, unsigned rpl, cudaStream_t s) {
    switch (rpl) {
        case 1: return launch_mv<T, 1, BOLTZ>(e, a, s);
        case 2: return launch_mv<T, 2, BOLTZ>(e, a, s);
        default: return launch_mv<T, 4, BOLTZ>(e, a, s);
    }
}

template <bool BOLTZ, int MB>
int dispatch_mma(Engine& e, MatvecArgs& a, unsigned nbk, cudaStream_t s) {
    switch (nbk) {
        case 1: return launch_mv<double, 1, BOLTZ, 1, MB>(e, a, s);
        case 2: return launch_mv<double, 1, BOLTZ, 2, MB>(e, a, s);
        case 4: return launch_mv<double, 1, BOLTZ, 4, MB>(e, a, s);
        case 3: if constexpr (!BOLTZ) return launch_mv<double, 1, BOLTZ, 3, MB>(e, a, s);
        case 5: if constexpr (!BOLTZ) return launch_mv<double, 1, BOLTZ, 5, MB>(e, a, s);
        case 6: if constexpr (!BOLTZ) return launch_mv<double, 1, BOLTZ, 6, MB>(e, a, s);
        case 7: if constexpr (!BOLTZ) return launch_mv<double, 1, BOLTZ, 7, MB>(e, a, s);
        default: return launch_mv<double, 1, BOLTZ, 8, MB>(e, a, s);
    }
}

}  // namespace

int launch_matvec(Engine& e, uint32_t B, const double* d_theta, double* d_out, uint64_t first_item,
                  uint64_t n_items, cudaStream_t s) {
    const bool boltz = e.desc.model == PF_MODEL_BOLTZMANN_DENSE;
    MatvecArgs a{};
    unsigned NG = 1;
    while (NG * MV_NODES_PER_LANE < B) NG *= 2;  // 1, 2, 4, 8
    const unsigned rpl = NG == 1 ? 1 : (NG == 2 ? 2 : 4);
    const bool mma = !e.f32 && e.mv_tensor;  // FP64: the contraction runs on the tensor cores (DMMA)
    // lasso on the tensor cores: node blocks of 8 in ANY number up to 8 (the DMMA-bound frontiers cost what their padded
    // width costs: B = 40 ran as 64 nodes); the CUDA-core layouts and the Boltzmann unit split keep powers of two
    if (mma && !boltz) NG = (B + MV_NODES_PER_LANE - 1) / MV_NODES_PER_LANE;
    // short tiles for a Boltzmann matrix whose 64-row tiles do not outnumber the SMs (see mv_frontier_kernel)
    const bool short_tiles = mma && boltz && (e.n_rows + 63) / 64 <= (uint64_t) e.sm_count;
    a.tile_rows = mma ? (short_tiles ? 64 : 256) : (MV_CONSUMER_WARPS / NG) * 32 * rpl;  // vector: 256, 256, 256, 128
    a.B = B;
    a.Bpad = NG * MV_NODES_PER_LANE;
    a.theta = d_theta;
    a.theta_len = e.theta_len;
    a.betaT = e.d_betaT;
    a.y = e.d_resp;
    a.part = e.d_part;
    a.out = d_out;
    a.ticket = e.d_ticket;
    a.cols = e.dim;
    a.cols_pad = e.cols_pad;
    a.item_rows = boltz ? e.n_rows : e.item_rows;
    a.row_begin = boltz ? 0 : first_item * e.item_rows;
    a.row_end = boltz ? e.n_rows : (first_item + n_items) * e.item_rows;
    if (a.row_end > e.n_rows) a.row_end = e.n_rows;
    a.inv_temperature = 1.0 / e.temperature;
    a.done_flag = e.done_flag;
    a.done_seq = e.done_seq;
    a.trace = e.d_trace;
    if (a.row_end <= a.row_begin) return launch_exchange_only(e, B, d_out, s);  // no rows here: zeros (+ the exchange)
    a.xchg = e.next_xchg();
    // PF_F32 lasso: the contraction on the tcgen05 tensor cores, FP32-faithful (3xTF32, pf_matvec_tc.cu).  Measured at C4
    // (n = 1e5, p = 1000; ms per call, tcgen05 / CUDA cores): B = 1 0.080 / 0.089, 8 0.078 / 0.091, 16 0.083 / 0.118,
    // 32 0.092 / 0.188, 64 0.122 / 0.361 -- the default for every B; PF_OPT_MATVEC_TENSOR = 0 selects the CUDA cores.
    if (e.f32 && !boltz && e.mv_tensor) return launch_matvec_tc32(e, a, d_theta, s);  // (faster than the CUDA cores at every B)
    const unsigned off = boltz ? 0 : 1;
    const unsigned total = e.cols_pad * a.Bpad;
    static bool prep_configured[64] = {};
    if (!prep_configured[e.device & 63]) {
        // same shared-memory carve-out as the frontier kernel, so the SMs are not reconfigured twice per call
        cudaFuncSetAttribute(mv_prep_panel_frag, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(mv_prep_panel<double>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(mv_prep_panel<float>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        prep_configured[e.device & 63] = true;
    }
    if (mma)
        mv_prep_panel_frag<<<(total + 255) / 256, 256, 0, s>>>(d_theta, static_cast<double*>(e.d_betaT), B, NG, e.theta_len,
                                                               off, e.dim, e.cols_pad);
    else if (e.f32)
        mv_prep_panel<float><<<(total + 255) / 256, 256, 0, s>>>(d_theta, static_cast<float*>(e.d_betaT), B, a.Bpad,
                                                                 e.theta_len, off, e.dim, e.cols_pad);
    else
        mv_prep_panel<double><<<(total + 255) / 256, 256, 0, s>>>(d_theta, static_cast<double*>(e.d_betaT), B, a.Bpad,
                                                                  e.theta_len, off, e.dim, e.cols_pad);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("mv_prep_panel launch: %s", cudaGetErrorString(err));
        return -1;
    }
    int r;
    if (mma)
        r = boltz ? (short_tiles ? dispatch_mma<true, 1>(e, a, NG, s) : dispatch_mma<true, 4>(e, a, NG, s))
                  : dispatch_mma<false, 4>(e, a, NG, s);
    else if (boltz)
        r = e.f32 ? dispatch_rpl<float, true>(e, a, rpl, s) : dispatch_rpl<double, true>(e, a, rpl, s);
    else
        r = e.f32 ? dispatch_rpl<float, false>(e, a, rpl, s) : dispatch_rpl<double, false>(e, a, rpl, s);
    return r < 0 ? r : r + 1;
}

}  // namespace pf
