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
