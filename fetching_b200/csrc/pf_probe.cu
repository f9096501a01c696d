// pf_probe.cu -- FP64 pipe probe (diagnostics).  BASELINE.md: "FP64 peak is not in
// MEASURED_PEAKS.json; the builder must measure it".  Measures DFMA throughput per SM per clock
// for ILP in {1,2,4,8} dependent chains per thread and a given number of warps per SM, from which
// both the FP64 roofline denominator and the pipe's dependent-issue latency follow.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pfetch.h"

namespace pf {

template <int ILP>
__global__ void fp64_probe_kernel(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], b, a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;  // keep the chains alive
}

// FP64 chains with KINT integer multiply-adds interleaved per DFMA: does integer work issue in the
// shadow of an FP64 instruction (which holds its pipe for 2 cycles per warp), or does it cost issue slots?
template <int ILP, int KINT>
__global__ void fp64_mix_probe_kernel(double* out, int iters, double a, double b, int c) {
    double x[ILP];
    int y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        x[i] = a + i + threadIdx.x * 1e-3;
        y[i] = threadIdx.x + i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                x[i] = fma(x[i], b, a);
#pragma unroll
                for (int k = 0; k < KINT; ++k) y[i] = y[i] * c + (k + 1);
            }
    }
    double s = 0;
    int t = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        s += x[i];
        t += y[i];
    }
    if (s == 12345.678 || t == 123456789) out[0] = s + t;
}

// DFMA with 1, 2 or 3 DISTINCT register operands (the rest loop-invariant / uniform): register-file bandwidth.
//   NREG = 1:  x = fma(x, B, A)        (B, A uniform)
//   NREG = 2:  x = fma(x, y, A)        (y a per-thread register)
//   NREG = 3:  x = fma(y, w, x)        (the accumulate form of a dot product: three distinct register pairs)
template <int NREG>
__global__ void fp64_rf_probe_kernel(double* out, int iters, double a, double b) {
    double x[8], y[8], w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = a + i + threadIdx.x * 1e-3;
        y[i] = b - i * 1e-9 - threadIdx.x * 1e-12;
        w[i] = 1e-9 * (i + 1) + threadIdx.x * 1e-13;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (NREG == 1) x[i] = fma(x[i], b, a);
                else if (NREG == 2) x[i] = fma(x[i], y[i], a);
                else x[i] = fma(y[i], w[(i + u) & 7], x[i]);
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + y[i] + w[i];
    if (s == 12345.678) out[0] = s;
}

// DFMA with two distinct register pairs + KINT cheap one-register integer ops per DFMA: does the second
// register-file cycle of the DFMA block the dispatch of the integer op?
template <int KINT>
__global__ void fp64_rf_mix_probe_kernel(double* out, int iters, double a, double b) {
    double x[8], y[8];
    unsigned z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = a + i + threadIdx.x * 1e-3;
        y[i] = 0.5 * b - i * 1e-9 - threadIdx.x * 1e-12;
        z[i] = threadIdx.x * 7 + i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x[i] = fma(x[i], y[i], 1.5);  // immediate addend: exactly two register pairs
#pragma unroll
                for (int k = 0; k < KINT; ++k) z[i] = (z[i] ^ (0x9e3779b9u + 16 * u + k)) + 0x7f4a7c15u;  // LOP3 + IADD
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + y[i] + (double) z[i];
    if (s == 12345.678) out[0] = s;
}

// FP64 tensor-core probe: mma.sync.m8n8k4.f64 (SASS DMMA.884), ILP independent 8x8 accumulator tiles per warp.
// One instruction = 8*8*4 = 256 FMAs per warp.  On B200 the FP64 tensor rate is specified equal to the vector
// rate; this measures what the part actually issues (the B300 guide's "vestigial FP64" does not apply here).
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int ILP>
__global__ void dmma_probe_kernel(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        c0[i] = a + i + threadIdx.x * 1e-3;
        c1[i] = a - i;
    }
    const double av = a * 1e-3 + threadIdx.x * 1e-9, bv = b * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
static float run_dmma_probe(int sms, int warps_per_sm, int iters, double* d_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int threads = 256, ctas_per_sm = warps_per_sm / 8;
    dmma_probe_kernel<ILP><<<sms * ctas_per_sm, threads>>>(d_out, 16, 1.0000001, 0.9999999);
    cudaEventRecord(e0);
    dmma_probe_kernel<ILP><<<sms * ctas_per_sm, threads>>>(d_out, iters, 1.0000001, 0.9999999);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

template <int KINT>
static float run_rf_mix_probe(int sms, int iters, double* d_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fp64_rf_mix_probe_kernel<KINT><<<sms * 2, 256>>>(d_out, 16, 1.0000001, 0.9999999);
    cudaEventRecord(e0);
    fp64_rf_mix_probe_kernel<KINT><<<sms * 2, 256>>>(d_out, iters, 1.0000001, 0.9999999);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

template <int NREG>
static float run_rf_probe(int sms, int iters, double* d_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fp64_rf_probe_kernel<NREG><<<sms * 2, 256>>>(d_out, 16, 1.0000001, 0.9999999);
    cudaEventRecord(e0);
    fp64_rf_probe_kernel<NREG><<<sms * 2, 256>>>(d_out, iters, 1.0000001, 0.9999999);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

template <int KINT>
static float run_mix_probe(int sms, int iters, double* d_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fp64_mix_probe_kernel<8, KINT><<<sms * 2, 256>>>(d_out, 16, 1.0000001, 0.9999999, 3);
    cudaEventRecord(e0);
    fp64_mix_probe_kernel<8, KINT><<<sms * 2, 256>>>(d_out, iters, 1.0000001, 0.9999999, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

template <int ILP>
static float run_probe(int sms, int warps_per_sm, int iters, double* d_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int threads = warps_per_sm * 32 > 1024 ? 1024 : warps_per_sm * 32;
    const int ctas_per_sm = (warps_per_sm * 32 + threads - 1) / threads;
    fp64_probe_kernel<ILP><<<sms * ctas_per_sm, threads>>>(d_out, 16, 1.0000001, 0.9999999);
    cudaEventRecord(e0);
    fp64_probe_kernel<ILP><<<sms * ctas_per_sm, threads>>>(d_out, iters, 1.0000001, 0.9999999);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

}  // namespace pf

// out[4 * w + i] = DFMA per second (1e12 units: TFMA/s) for warps_per_sm = {4, 8, 16, 32}[w], ILP = {1,2,4,8}[i]
extern "C" PF_API int pf_debug_fp64_probe(int device, double* out16) {
    if (!out16 || cudaSetDevice(device) != cudaSuccess) return PF_ERR_INVALID;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PF_ERR_CUDA;
    const int sms = prop.multiProcessorCount;
    double* d_out = nullptr;
    if (cudaMalloc(&d_out, 64) != cudaSuccess) return PF_ERR_NOMEM;
    const int iters = 2048;
    const int wlist[4] = {4, 8, 16, 32};
    for (int w = 0; w < 4; ++w) {
        float ms[4];
        ms[0] = pf::run_probe<1>(sms, wlist[w], iters, d_out);
        ms[1] = pf::run_probe<2>(sms, wlist[w], iters, d_out);
        ms[2] = pf::run_probe<4>(sms, wlist[w], iters, d_out);
        ms[3] = pf::run_probe<8>(sms, wlist[w], iters, d_out);
        const int ilp[4] = {1, 2, 4, 8};
        for (int i = 0; i < 4; ++i) {
            const double fmas = (double) sms * wlist[w] * 32.0 * iters * 16.0 * ilp[i];
            out16[4 * w + i] = fmas / (ms[i] * 1e-3) / 1e12;
        }
    }
    cudaFree(d_out);
    return cudaGetLastError() == cudaSuccess ? PF_OK : PF_ERR_CUDA;
}

// out8[4 * w + i] = FMA/s (1e12) through DMMA m8n8k4 for warps_per_sm = {8, 16}[w], ILP = {1,2,4,8}[i]
extern "C" PF_API int pf_debug_dmma_probe(int device, double* out8) {
    if (!out8 || cudaSetDevice(device) != cudaSuccess) return PF_ERR_INVALID;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PF_ERR_CUDA;
    const int sms = prop.multiProcessorCount, iters = 1024;
    double* d_out = nullptr;
    if (cudaMalloc(&d_out, 64) != cudaSuccess) return PF_ERR_NOMEM;
    const int wlist[2] = {8, 16};
    for (int w = 0; w < 2; ++w) {
        const float ms[4] = {pf::run_dmma_probe<1>(sms, wlist[w], iters, d_out), pf::run_dmma_probe<2>(sms, wlist[w], iters, d_out),
                             pf::run_dmma_probe<4>(sms, wlist[w], iters, d_out), pf::run_dmma_probe<8>(sms, wlist[w], iters, d_out)};
        const int ilp[4] = {1, 2, 4, 8};
        for (int i = 0; i < 4; ++i)
            out8[4 * w + i] = (double) sms * wlist[w] * iters * 16.0 * ilp[i] * 256.0 / (ms[i] * 1e-3) / 1e12;
    }
    cudaFree(d_out);
    return cudaGetLastError() == cudaSuccess ? PF_OK : PF_ERR_CUDA;
}

// out4[k] = DFMA rate (1e12/s) at 16 warps/SM, ILP 8, with k = 0..3 integer multiply-adds per DFMA
extern "C" PF_API int pf_debug_issue_probe(int device, double* out4) {
    if (!out4 || cudaSetDevice(device) != cudaSuccess) return PF_ERR_INVALID;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PF_ERR_CUDA;
    const int sms = prop.multiProcessorCount, iters = 1024;
    double* d_out = nullptr;
    if (cudaMalloc(&d_out, 64) != cudaSuccess) return PF_ERR_NOMEM;
    const float ms[4] = {pf::run_mix_probe<0>(sms, iters, d_out), pf::run_mix_probe<1>(sms, iters, d_out),
                         pf::run_mix_probe<2>(sms, iters, d_out), pf::run_mix_probe<3>(sms, iters, d_out)};
    for (int k = 0; k < 4; ++k) out4[k] = (double) sms * 2 * 256 * iters * 16.0 * 8 / (ms[k] * 1e-3) / 1e12;
    cudaFree(d_out);
    return cudaGetLastError() == cudaSuccess ? PF_OK : PF_ERR_CUDA;
}

// out3[n-1] = DFMA rate (1e12/s) at 16 warps/SM, 8 chains, with n = 1..3 distinct register operands per DFMA
extern "C" PF_API int pf_debug_rf_probe(int device, double* out3) {
    if (!out3 || cudaSetDevice(device) != cudaSuccess) return PF_ERR_INVALID;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PF_ERR_CUDA;
    const int sms = prop.multiProcessorCount, iters = 1024;
    double* d_out = nullptr;
    if (cudaMalloc(&d_out, 64) != cudaSuccess) return PF_ERR_NOMEM;
    const float ms[3] = {pf::run_rf_probe<1>(sms, iters, d_out), pf::run_rf_probe<2>(sms, iters, d_out),
                         pf::run_rf_probe<3>(sms, iters, d_out)};
    for (int k = 0; k < 3; ++k) out3[k] = (double) sms * 2 * 256 * iters * 16.0 * 8 / (ms[k] * 1e-3) / 1e12;
    // out3[3], out3[4]: two-register DFMA with 1 resp. 2 single-register ALU ops (2 resp. 4 instructions) between
    const float mm[2] = {pf::run_rf_mix_probe<1>(sms, iters, d_out), pf::run_rf_mix_probe<2>(sms, iters, d_out)};
    for (int k = 0; k < 2; ++k) out3[3 + k] = (double) sms * 2 * 256 * iters * 16.0 * 8 / (mm[k] * 1e-3) / 1e12;
    cudaFree(d_out);
    return cudaGetLastError() == cudaSuccess ? PF_OK : PF_ERR_CUDA;
}

// Do DFMA (FP64 pipe) and DMMA m8n8k4 (tensor pipe) overlap?  NF two-register DFMAs (8 chains) and ND DMMAs with fresh
// (zero) accumulators per block, 16 warps per SM.  out[i] = microseconds per 1000 blocks per warp-slot for
// (NF, ND) = (16,0) (0,2) (0,4) (16,1) (16,2) (16,4) (32,4) (8,4).
namespace pf {
template <int NF, int ND>
__global__ void __launch_bounds__(256, 2) dfma_dmma_probe_kernel(double* out, int iters, double a, double b) {
    double x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = a + i + threadIdx.x * 1e-3;
        y[i] = 0.5 * b - i * 1e-9 - threadIdx.x * 1e-12;
    }
    double av = a * 1e-3 + threadIdx.x * 1e-9;
    const double bv = b * 1e-3;
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        double c0[ND > 0 ? ND : 1], c1[ND > 0 ? ND : 1];
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%4};"
                         : "=d"(c0[d]), "=d"(c1[d])
                         : "d"(av), "d"(bv + d), "d"(0.0));
        }
#pragma unroll
        for (int u = 0; u < NF; ++u) x[u & 7] = fma(x[u & 7], y[u & 7], 1.5);
#pragma unroll
        for (int d = 0; d < ND; ++d) acc += c0[d] * 1e-30 + c1[d] * 1e-30;  // 2 more DFMAs per DMMA consume the tile
        av += 1e-12;
    }
    double s = acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}
template <int NF, int ND>
static float run_dfma_dmma_probe(int sms, int iters, double* d_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dfma_dmma_probe_kernel<NF, ND><<<sms * 2, 256>>>(d_out, 16, 1.0000001, 0.9999999);
    cudaEventRecord(e0);
    dfma_dmma_probe_kernel<NF, ND><<<sms * 2, 256>>>(d_out, iters, 1.0000001, 0.9999999);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}
}  // namespace pf

// out8[i] = SM clocks (at 1.965 GHz) per block per SM sub-partition (4 warps each) for the (NF, ND) list above
extern "C" PF_API int pf_debug_dfma_dmma_probe(int device, double* out8) {
    if (!out8 || cudaSetDevice(device) != cudaSuccess) return PF_ERR_INVALID;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PF_ERR_CUDA;
    const int sms = prop.multiProcessorCount, iters = 20000;
    double* d_out = nullptr;
    if (cudaMalloc(&d_out, 64) != cudaSuccess) return PF_ERR_NOMEM;
    const float ms[8] = {pf::run_dfma_dmma_probe<16, 0>(sms, iters, d_out), pf::run_dfma_dmma_probe<0, 2>(sms, iters, d_out),
                         pf::run_dfma_dmma_probe<0, 4>(sms, iters, d_out),  pf::run_dfma_dmma_probe<16, 1>(sms, iters, d_out),
                         pf::run_dfma_dmma_probe<16, 2>(sms, iters, d_out), pf::run_dfma_dmma_probe<16, 4>(sms, iters, d_out),
                         pf::run_dfma_dmma_probe<32, 4>(sms, iters, d_out), pf::run_dfma_dmma_probe<8, 4>(sms, iters, d_out)};
    // 16 warps per SM = 4 per sub-partition: clocks per block per warp = time * f / iters / 4
    for (int i = 0; i < 8; ++i) out8[i] = (double) ms[i] * 1e-3 * 1.965e9 / iters / 4.0;
    cudaFree(d_out);
    return cudaGetLastError() == cudaSuccess ? PF_OK : PF_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------------------------
// Host-buffer call floor: one kernel launch whose last thread delivers a self-validating word to mapped pinned
// memory, polled by the host -- the skeleton of pf_eval_frontier without any scoring.  Measured for a small and a
// 3.6 KB parameter block (PointwiseArgs carries up to 432 inline doubles) and for 1 CTA vs a full grid of
// 2 x 148 x 256 threads with a ticket (what the frontier kernels do).  out[4] = microseconds per call:
// {small params, 1 CTA}, {3.6 KB params, 1 CTA}, {small, full grid + ticket}, {3.6 KB, full grid + ticket}.
// ---------------------------------------------------------------------------------------------------------------
namespace pf {
struct ProbeSmall {
    unsigned long long* out;
    unsigned int* ticket;
    unsigned long long seq;
};
struct ProbeBig {
    unsigned long long* out;
    unsigned int* ticket;
    unsigned long long seq;
    double pad[448];
};
template <class A>
__global__ void __launch_bounds__(256) launch_probe_kernel(const __grid_constant__ A a) {
    __shared__ int last;
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(a.ticket, 1u);
        last = prev == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        *a.ticket = 0u;
        *reinterpret_cast<volatile unsigned long long*>(a.out) = a.seq;
    }
}
template <class A>
static double launch_probe_run(unsigned grid, int iters, unsigned long long* h_out, unsigned long long* d_out,
                               unsigned int* d_ticket, cudaStream_t s, unsigned long long& seq) {
    A a{};
    a.out = d_out;
    a.ticket = d_ticket;
    auto once = [&]() {
        a.seq = ++seq;
        launch_probe_kernel<A><<<grid, 256, 0, s>>>(a);
        volatile unsigned long long* w = h_out;
        while (*w != a.seq) __builtin_ia32_pause();
    };
    for (int i = 0; i < 20; ++i) once();
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < iters; ++i) once();
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return ((double) (t1.tv_sec - t0.tv_sec) * 1e6 + 1e-3 * (double) (t1.tv_nsec - t0.tv_nsec)) / iters;
}
}  // namespace pf

extern "C" PF_API int pf_debug_launch_probe(int device, double* out4) {
    if (!out4 || cudaSetDevice(device) != cudaSuccess) return PF_ERR_INVALID;
    cudaStream_t s;
    unsigned long long *h_out = nullptr, *d_out = nullptr;
    unsigned int* d_ticket = nullptr;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess ||
        cudaHostAlloc(&h_out, 64, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&d_out, h_out, 0) != cudaSuccess || cudaMalloc(&d_ticket, 64) != cudaSuccess ||
        cudaMemset(d_ticket, 0, 64) != cudaSuccess)
        return PF_ERR_CUDA;
    *h_out = 0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const unsigned full = 2u * (unsigned) prop.multiProcessorCount;
    unsigned long long seq = 0;
    out4[0] = pf::launch_probe_run<pf::ProbeSmall>(1, 2000, h_out, d_out, d_ticket, s, seq);
    out4[1] = pf::launch_probe_run<pf::ProbeBig>(1, 2000, h_out, d_out, d_ticket, s, seq);
    out4[2] = pf::launch_probe_run<pf::ProbeSmall>(full, 2000, h_out, d_out, d_ticket, s, seq);
    out4[3] = pf::launch_probe_run<pf::ProbeBig>(full, 2000, h_out, d_out, d_ticket, s, seq);
    cudaStreamSynchronize(s);
    cudaFree(d_ticket);
    cudaFreeHost(h_out);
    cudaStreamDestroy(s);
    return PF_OK;
}
