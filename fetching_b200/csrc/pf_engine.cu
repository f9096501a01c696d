// pf_engine.cu -- the C ABI declared in include/pfetch.h: engine lifetime, data residency,
// host- and device-pointer evaluation paths, the multi-GPU set-up (peer exchange buffers over CUDA IPC for the
// in-kernel exchange of the per-node partials; NCCL all-reduce / all-gather as the alternative).
//
// There is NO CPU fallback anywhere in this file: a missing sm_100 device is PF_ERR_NO_DEVICE.
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <vector>

#include "pf_engine.h"

#if __has_include(<nccl.h>)
#include <nccl.h>
#define PF_HAVE_NCCL_H 1
#else
#define PF_HAVE_NCCL_H 0
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
#endif

namespace pf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

#define PF_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t err__ = (call);                                                          \
        if (err__ != cudaSuccess) {                                                          \
            set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return PF_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

// ------------------------------------------------------------------------------------ NCCL (dlopen)
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
            api.GetUniqueId = (decltype(api.GetUniqueId)) dlsym(api.handle, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank)) dlsym(api.handle, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy)) dlsym(api.handle, "ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce)) dlsym(api.handle, "ncclAllReduce");
            api.AllGather = (decltype(api.AllGather)) dlsym(api.handle, "ncclAllGather");
            api.GetErrorString = (decltype(api.GetErrorString)) dlsym(api.handle, "ncclGetErrorString");
        }
    }
    if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce) return nullptr;
    return &api;
}

PeerXchg Engine::next_xchg() {
    PeerXchg x{};
    if (peer_fused && nranks > 1 && (!node_split || gather_per)) {
        x.gather_per = node_split ? gather_per : 0;
        x.gather_total = node_split ? gather_total : 0;
        x.peers = d_peers;
        x.status = reinterpret_cast<unsigned int*>(d_hout + kHostOutWords + 1);
        x.seq = ++xchg_seq;
        x.rank = (unsigned) rank;
        x.nranks = (unsigned) nranks;
    }
    return x;
}

static bool peer_active(const Engine& e) { return e.peer_fused && e.nranks > 1 && !e.node_split; }

// ------------------------------------------------------------------------------------ resident kernel (host side)
// One resident frontier kernel (pf_resident.cu) can be on a device's SMs at a time -- it fills them.  Whoever launches
// anything else on that device from this process (another engine, or this engine on a path the resident kernel does
// not serve) asks it to leave first; work this process knows nothing about simply waits for the idle timeout.
static Engine* g_res_owner[64] = {};

static bool resident_exited(const Resident& r) {
    const volatile unsigned long long* st = r.h_words + kResStateWord;
    return (*st >> 32) == r.epoch && (*st & 1ull);
}

static void resident_stop(Engine& e) {
    Resident& r = e.res;
    if (!r.alive) return;
    if (!resident_exited(r)) {
        *(volatile unsigned long long*) (r.h_words + kResCtlWord) = r.epoch;
        timespec t0, now;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (unsigned long long spins = 1; !resident_exited(r); ++spins) {
            __builtin_ia32_pause();
            if ((spins & 0xFFFF) == 0) {
                if (cudaStreamQuery(e.stream) != cudaErrorNotReady) break;  // gone (or failed: the next call reports it)
                clock_gettime(CLOCK_MONOTONIC, &now);
                if ((double) (now.tv_sec - t0.tv_sec) + 1e-9 * (double) (now.tv_nsec - t0.tv_nsec) > 10.0) {
                    e.wedged = true;  // it will still leave by its idle timeout; nothing more to do from here
                    break;
                }
            }
        }
    }
    r.alive = false;
    if (g_res_owner[e.device & 63] == &e) g_res_owner[e.device & 63] = nullptr;
}

static void resident_stop_device(int device) {
    Engine* o = g_res_owner[device & 63];
    if (o) resident_stop(*o);
}

// The engine has one set of scratch buffers (records, ticket, panel, exchange sequence): a launch on a stream other
// than the previous launch's is ordered behind it.  Costs nothing while the caller stays on one stream.
static int order_streams(Engine& e, cudaStream_t s) {
    resident_stop_device(e.device);  // a launched kernel needs the SMs the resident one occupies
    if (e.last_stream != nullptr && e.last_stream != s) {
        if (!e.order_event && cudaEventCreateWithFlags(&e.order_event, cudaEventDisableTiming) != cudaSuccess) {
            set_error("cudaEventCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
            return PF_ERR_CUDA;
        }
        if (cudaEventRecord(e.order_event, e.last_stream) != cudaSuccess ||
            cudaStreamWaitEvent(s, e.order_event, 0) != cudaSuccess) {
            cudaGetLastError();  // e.g. the caller destroyed the previous stream: fall back to a full barrier
            if (cudaDeviceSynchronize() != cudaSuccess) {
                set_error("cudaDeviceSynchronize failed: %s", cudaGetErrorString(cudaGetLastError()));
                return PF_ERR_CUDA;
            }
        }
    }
    e.last_stream = s;
    return PF_OK;
}

static int allreduce_out(Engine& e, double* d_out, uint32_t B, cudaStream_t s) {
    if (peer_active(e)) return PF_OK;  // already reduced inside the frontier kernel (peer_allreduce)
    if (!e.nccl_comm || e.nranks <= 1 || e.node_split) return PF_OK;
    NcclApi* n = nccl_api();
    if (!n) {
        set_error("NCCL library not loadable");
        return PF_ERR_NCCL;
    }
    // 2*B doubles: latency bound (<= 1 KB); NCCL picks its LL protocol over NVLink/NVSwitch
    ncclResult_t r = n->AllReduce(d_out, d_out, (size_t) 2 * B, /*ncclDouble*/ 8, /*ncclSum*/ 0,
                                  (ncclComm_t) e.nccl_comm, s);
    if (r != 0) {
        set_error("ncclAllReduce: %s", n->GetErrorString ? n->GetErrorString(r) : "error");
        return PF_ERR_NCCL;
    }
    return PF_OK;
}

// ------------------------------------------------------------------------------------ helpers
static bool is_matvec(const Engine& e) {
    return e.desc.model == PF_MODEL_BLASSO || e.desc.model == PF_MODEL_BOLTZMANN_DENSE;
}

static int launch_model(Engine& e, uint32_t B, const double* d_theta, double* d_out, uint64_t first_item,
                        uint64_t n_items, cudaStream_t s) {
    int r = order_streams(e, s);
    if (r != PF_OK) return r;
    switch (e.desc.model) {
        case PF_MODEL_NORMAL:
        case PF_MODEL_MIXTURE: r = launch_pointwise(e, B, d_theta, d_out, first_item, n_items, s); break;
        case PF_MODEL_BLASSO:
        case PF_MODEL_BOLTZMANN_DENSE: r = launch_matvec(e, B, d_theta, d_out, first_item, n_items, s); break;
        case PF_MODEL_BOLTZMANN_SCALAR: r = launch_boltzmann_scalar(e, B, d_theta, d_out, n_items, s); break;
        default: set_error("unknown model"); return PF_ERR_INVALID;
    }
    if (r < 0) return PF_ERR_CUDA;
    e.launches += (uint64_t) r;
    return PF_OK;
}

static bool all_identity_cov(const Engine& e, const double* thetas, uint32_t B) {
    const uint32_t D = e.dim;
    for (uint32_t b = 0; b < B; ++b) {
        const double* th = thetas + (size_t) b * e.theta_len;
        for (uint32_t i = 0, x = 0; i < D; ++i)
            for (uint32_t j = 0; j <= i; ++j, ++x)
                if (th[D + x] != (j == i ? 1.0 : 0.0)) return false;  // normaltheta.hh:116-125
    }
    return true;
}

static void destroy(Engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    resident_stop(*e);
    if (g_res_owner[e->device & 63] == e) g_res_owner[e->device & 63] = nullptr;
    if (e->nccl_comm) {
        NcclApi* n = nccl_api();
        if (n && n->CommDestroy) n->CommDestroy((ncclComm_t) e->nccl_comm);
    }
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->d_data);
    cudaFree(e->d_resp);
    cudaFree(e->d_theta);
    cudaFree(e->d_out);
    cudaFree(e->d_part);
    cudaFree(e->d_betaT);
    cudaFree(e->d_order);
    cudaFree(e->d_ns_send);
    cudaFree(e->d_ns_recv);
    cudaFree(e->d_ticket);
    for (int r = 0; r < kXchgMaxRanks; ++r)
        if (e->peer_mapped[r]) cudaIpcCloseMemHandle(e->peer_mapped[r]);
    cudaFree(e->d_xchg);
    cudaFree(e->d_peers);
    if (e->h_theta) cudaFreeHost(e->h_theta);
    if (e->h_out) cudaFreeHost(e->h_out);
    if (e->h_trace) cudaFreeHost(e->h_trace);
    if (e->res.h_words) cudaFreeHost(e->res.h_words);
    cudaFree(e->res.d_relay);
    cudaFree(e->res.d_ticket);
    if (e->order_event) cudaEventDestroy(e->order_event);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

static int create(const pf_model_desc* d, Engine** out) {
    if (!d || !out) {
        set_error("null argument");
        return PF_ERR_INVALID;
    }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || d->device < 0 || d->device >= ndev) {
        set_error("no CUDA device %d (found %d); this engine has no CPU fallback", d->device, ndev);
        return PF_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    PF_CUDA(cudaGetDeviceProperties(&prop, d->device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; the kernels are built for sm_100a only", d->device, prop.major, prop.minor);
        return PF_ERR_NO_DEVICE;
    }
    PF_CUDA(cudaSetDevice(d->device));

    Engine* e = new Engine();
    e->desc = *d;
    e->desc.data = nullptr;
    e->desc.response = nullptr;
    e->device = d->device;
    e->sm_count = prop.multiProcessorCount;
    e->f32 = d->dtype == PF_F32;
    e->dim = d->dim;
    e->K = d->components;
    e->n_rows = d->n_rows;
    e->temperature = d->param != 0.0 ? d->param : 1.0;
    for (double& v : e->xmax) v = HUGE_VAL;

    auto fail = [&](int code) {
        destroy(e);
        return code;
    };
    if (d->dtype != PF_F64 && d->dtype != PF_F32) {
        set_error("unknown dtype %d", d->dtype);
        return fail(PF_ERR_INVALID);
    }

    switch (d->model) {
        case PF_MODEL_NORMAL:
            if (!d->data || d->n_rows == 0 || d->dim == 0) { set_error("normal: data/n_rows/dim required"); return fail(PF_ERR_SHAPE); }
            e->item_rows = 1;
            e->n_items = d->n_items ? d->n_items : d->n_rows;
            e->theta_len = d->dim + d->dim * (d->dim + 1) / 2;
            break;
        case PF_MODEL_MIXTURE:
            if (!d->data || d->n_rows == 0 || d->dim == 0 || d->components == 0 || d->components > 32) {
                set_error("mixture: data, n_rows, dim and 1 <= components <= 32 required");
                return fail(PF_ERR_SHAPE);
            }
            e->item_rows = d->item_rows ? d->item_rows : (d->n_rows >= 100 ? d->n_rows / 100 : 1);  // pymixture.py:195
            if (d->n_items)
                e->n_items = d->n_items;
            else {
                e->n_items = (d->n_rows + 1) / e->item_rows;  // pymixture.py:64
                // the reference cannot score a short last item (pymixture.py:88-91): whole items only
                if (e->n_items * e->item_rows > d->n_rows) e->n_items = d->n_rows / e->item_rows;
            }
            e->theta_len = d->components * d->dim;
            break;
        case PF_MODEL_BLASSO:
            if (!d->data || !d->response || d->n_rows == 0 || d->dim == 0) {
                set_error("lasso: data, response, n_rows, dim required");
                return fail(PF_ERR_SHAPE);
            }
            e->item_rows = d->item_rows ? d->item_rows : 18000;  // pyblasso.py:136-137
            e->n_items = d->n_items ? d->n_items : d->n_rows / e->item_rows;  // pyblasso.py:69
            e->theta_len = d->dim + 1;
            break;
        case PF_MODEL_BOLTZMANN_DENSE:
            if (!d->data || d->dim == 0 || d->n_rows != d->dim) {
                set_error("dense Boltzmann: W must be dim x dim (n_rows == dim)");
                return fail(PF_ERR_SHAPE);
            }
            e->item_rows = d->n_rows;
            e->n_items = 1;
            e->theta_len = d->dim;
            break;
        case PF_MODEL_BOLTZMANN_SCALAR:
            e->item_rows = 1;
            e->n_items = d->n_items ? d->n_items : (d->n_rows ? d->n_rows : 1);
            e->n_rows = e->n_items;
            e->theta_len = 1;
            break;
        default:
            set_error("unknown model %d", d->model);
            return fail(PF_ERR_INVALID);
    }
    if (d->model != PF_MODEL_BOLTZMANN_SCALAR && d->model != PF_MODEL_BOLTZMANN_DENSE &&
        (e->n_items == 0 || e->n_items * e->item_rows > e->n_rows)) {
        set_error("n_items (%llu) x item_rows (%llu) inconsistent with n_rows (%llu)",
                  (unsigned long long) e->n_items, (unsigned long long) e->item_rows, (unsigned long long) e->n_rows);
        return fail(PF_ERR_SHAPE);
    }
    if ((d->model == PF_MODEL_NORMAL || d->model == PF_MODEL_MIXTURE) && !pointwise_supported(*e)) {
        set_error("dimension %u not built for this model (1,2,3,4,8)", d->dim);
        return fail(PF_ERR_UNSUPPORTED);
    }
    e->max_frontier = is_matvec(*e) || d->model == PF_MODEL_BOLTZMANN_SCALAR ? 64 : pointwise_max_frontier(*e);

#define PF_CUDA_F(call)                                                                       \
    do {                                                                                      \
        cudaError_t err__ = (call);                                                           \
        if (err__ != cudaSuccess) {                                                           \
            set_error("%s failed: %s", #call, cudaGetErrorString(err__));                    \
            return fail(err__ == cudaErrorMemoryAllocation ? PF_ERR_NOMEM : PF_ERR_CUDA);    \
        }                                                                                     \
    } while (0)

    PF_CUDA_F(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    const size_t esz = e->f32 ? 4 : 8;

    // ---- data set -> device, once
    if (d->model == PF_MODEL_NORMAL || d->model == PF_MODEL_MIXTURE) {
        const size_t elems = (size_t) e->n_rows * e->dim;
        const size_t bytes = ((elems * esz + 15) & ~size_t(15)) + 64;  // tail slack for the aligned copy window
        PF_CUDA_F(cudaMalloc(&e->d_data, bytes));
        PF_CUDA_F(cudaMemsetAsync(e->d_data, 0, bytes, e->stream));
        if (!e->f32) {
            PF_CUDA_F(cudaMemcpyAsync(e->d_data, d->data, elems * 8, cudaMemcpyHostToDevice, e->stream));
            // while the copy runs: the data set's bounding box (max |x_j| per column; a NaN makes it +inf), which the
            // mixture kernel tests theta against before it uses the factored form of the exponent (pf_math.cuh)
            if (d->model == PF_MODEL_MIXTURE && e->dim <= 8) {
                double mx[8] = {};
                bool nan = false;
                const uint32_t D = e->dim;
                for (size_t r = 0; r < e->n_rows; ++r)
                    for (uint32_t j = 0; j < D; ++j) {
                        const double v = fabs(d->data[r * D + j]);
                        nan |= v != v;
                        mx[j] = v > mx[j] ? v : mx[j];
                    }
                for (uint32_t j = 0; j < D; ++j) e->xmax[j] = nan ? HUGE_VAL : mx[j];
            }
            PF_CUDA_F(cudaStreamSynchronize(e->stream));
        } else {
            const size_t chunk = size_t(1) << 24;
            std::vector<float> tmp(elems < chunk ? elems : chunk);
            for (size_t o = 0; o < elems; o += chunk) {
                const size_t n = elems - o < chunk ? elems - o : chunk;
                for (size_t i = 0; i < n; ++i) tmp[i] = (float) d->data[o + i];
                PF_CUDA_F(cudaMemcpyAsync((float*) e->d_data + o, tmp.data(), n * 4, cudaMemcpyHostToDevice, e->stream));
                PF_CUDA_F(cudaStreamSynchronize(e->stream));
            }
        }
    } else if (is_matvec(*e)) {
        const uint32_t kch = (uint32_t) (128 / esz);
        e->cols_pad = (e->dim + kch - 1) / kch * kch;  // columns rounded up to whole 128-byte chunks
        // Row pitch: whole chunks, but never a multiple of 2 KB.  A box is <= 256 rows x one 128-byte chunk, i.e. that many
        // requests exactly one pitch apart; with a power-of-two pitch (p = 1000 floats -> 4096 B) they all land on the same
        // HBM channel and queue up behind each other (ncu: 30 % DRAM throughput, 6 us per box) -- one chunk of padding per
        // row spreads them over the channels.
        size_t pitch = (size_t) e->cols_pad * esz;
        if (pitch % 2048 == 0) pitch += 128;
        e->pitch_bytes = pitch;
        const size_t bytes = pitch * e->n_rows + 1024;
        PF_CUDA_F(cudaMalloc(&e->d_data, bytes));
        PF_CUDA_F(cudaMemsetAsync(e->d_data, 0, bytes, e->stream));
        if (!e->f32) {
            PF_CUDA_F(cudaMemcpy2DAsync(e->d_data, pitch, d->data, (size_t) e->dim * 8, (size_t) e->dim * 8, e->n_rows,
                                        cudaMemcpyHostToDevice, e->stream));
            PF_CUDA_F(cudaStreamSynchronize(e->stream));
        } else {
            const size_t rows_chunk = ((size_t(1) << 24) / e->dim) ? ((size_t(1) << 24) / e->dim) : 1;
            std::vector<float> tmp(rows_chunk * e->dim);
            for (size_t r = 0; r < e->n_rows; r += rows_chunk) {
                const size_t nr = e->n_rows - r < rows_chunk ? e->n_rows - r : rows_chunk;
                for (size_t i = 0; i < nr * e->dim; ++i) tmp[i] = (float) d->data[r * e->dim + i];
                PF_CUDA_F(cudaMemcpy2DAsync((char*) e->d_data + r * pitch, pitch, tmp.data(), (size_t) e->dim * 4,
                                            (size_t) e->dim * 4, nr, cudaMemcpyHostToDevice, e->stream));
                PF_CUDA_F(cudaStreamSynchronize(e->stream));
            }
        }
        if (d->model == PF_MODEL_BLASSO) {
            PF_CUDA_F(cudaMalloc(&e->d_resp, e->n_rows * 8));
            PF_CUDA_F(cudaMemcpyAsync(e->d_resp, d->response, e->n_rows * 8, cudaMemcpyHostToDevice, e->stream));
            PF_CUDA_F(cudaStreamSynchronize(e->stream));
        }
        // coefficient panel: [cols_pad][64] of T, or hi + lo float parts for the tcgen05 path (both fit 8 bytes per entry)
        PF_CUDA_F(cudaMalloc(&e->d_betaT, (size_t) e->cols_pad * kMaxNodesPerPass * 8));
    }

    // ---- per-call buffers
    const size_t tbytes = (size_t) e->max_frontier * e->theta_len * 8;
    PF_CUDA_F(cudaMalloc(&e->d_theta, tbytes));
    PF_CUDA_F(cudaMalloc(&e->d_out, 2 * kMaxNodesPerPass * 8));
    PF_CUDA_F(cudaMalloc(&e->d_part, (size_t) kMaxGrid * kPartSlots * kMaxNodesPerPass * 8));
    PF_CUDA_F(cudaMalloc(&e->d_order, (size_t) kMaxNodesPerPass * 2 * sizeof(unsigned long long)));
    PF_CUDA_F(cudaMalloc(&e->d_ticket, 64));
    PF_CUDA_F(cudaMemsetAsync(e->d_ticket, 0, 64, e->stream));
    PF_CUDA_F(cudaHostAlloc(&e->h_theta, tbytes, cudaHostAllocMapped));
    PF_CUDA_F(cudaHostGetDevicePointer(&e->d_htheta, e->h_theta, 0));
    PF_CUDA_F(cudaHostAlloc(&e->h_out, (kHostOutWords + 8 + 2 * kMaxNodesPerPass) * 8, cudaHostAllocMapped));
    memset(e->h_out, 0, (kHostOutWords + 8 + 2 * kMaxNodesPerPass) * 8);
    PF_CUDA_F(cudaHostGetDevicePointer(&e->d_hout, e->h_out, 0));
    if (const char* tr = getenv("PF_TRACE"); tr && tr[0] == '1') {  // diagnostics: in-kernel time marks
        PF_CUDA_F(cudaHostAlloc(&e->h_trace, (32 + 2 * kMaxGrid) * sizeof(unsigned long long), cudaHostAllocMapped));
        memset(e->h_trace, 0, (32 + 2 * kMaxGrid) * sizeof(unsigned long long));
        PF_CUDA_F(cudaHostGetDevicePointer(&e->d_trace, e->h_trace, 0));
    }
    PF_CUDA_F(cudaStreamSynchronize(e->stream));
#undef PF_CUDA_F
    *out = e;
    return PF_OK;
}

// Collect the n result doubles the kernel's last CTA delivers as self-validating words in mapped host memory
// (PointwiseArgs::done_flag): polling beats copy + synchronise by ~15 us per call and needs no fence on the device
// side; falls back to the stream's status so that a failed kernel cannot hang the caller.
// Bounded: a kernel that neither delivers nor fails within PF_HOST_TIMEOUT_S seconds (default 120; far beyond any
// frontier pass, and beyond the 20 s in-kernel peer timeout) is reported as PF_ERR_CUDA and the engine is marked
// wedged -- every later call fails at once instead of hanging its caller behind a stuck stream.
static double host_timeout_s() {
    static double t = -1.0;
    if (t < 0.0) {
        const char* v = getenv("PF_HOST_TIMEOUT_S");
        t = v && atof(v) > 0.0 ? atof(v) : 120.0;
    }
    return t;
}

static int wait_words(Engine& e, unsigned n, unsigned long long seq, double* vals) {
    volatile unsigned long long* w = reinterpret_cast<volatile unsigned long long*>(e.h_out);
    const unsigned long long tag = seq & 0xFFFFFFFFull;
    unsigned long long spins = 0;
    timespec t_start = {0, 0};
    for (unsigned i = 0; i < n; ++i) {
        for (;;) {
            const unsigned long long lo = w[2 * i], hi = w[2 * i + 1];
            if ((lo >> 32) == tag && (hi >> 32) == tag) {
                const unsigned long long bits = (hi << 32) | (lo & 0xFFFFFFFFull);
                memcpy(&vals[i], &bits, 8);
                break;
            }
            __builtin_ia32_pause();
            if ((++spins & 0x3FFFF) == 0) {
                const cudaError_t q = cudaStreamQuery(e.stream);
                if (q == cudaSuccess) {
                    const unsigned long long lo2 = w[2 * i], hi2 = w[2 * i + 1];
                    if ((lo2 >> 32) == tag && (hi2 >> 32) == tag) continue;
                    set_error("kernel finished without delivering its results");
                    return PF_ERR_CUDA;
                }
                if (q != cudaErrorNotReady) {
                    set_error("stream error while waiting for the frontier kernel: %s", cudaGetErrorString(q));
                    return PF_ERR_CUDA;
                }
                timespec now;
                clock_gettime(CLOCK_MONOTONIC, &now);
                if (t_start.tv_sec == 0 && t_start.tv_nsec == 0) t_start = now;
                if ((double) (now.tv_sec - t_start.tv_sec) + 1e-9 * (double) (now.tv_nsec - t_start.tv_nsec) > host_timeout_s()) {
                    e.wedged = true;
                    set_error("the frontier kernel delivered nothing within %.0f s (PF_HOST_TIMEOUT_S): engine marked unusable",
                              host_timeout_s());
                    return PF_ERR_CUDA;
                }
            }
        }
    }
    return PF_OK;
}

// a call number whose low 32 bits are never 0 (the tag of the result words; the buffer starts zeroed)
static unsigned long long next_host_seq(Engine& e) {
    if ((++e.host_seq & 0xFFFFFFFFull) == 0) ++e.host_seq;
    return e.host_seq;
}

static int eval_host(Engine& e, uint32_t B, const double* thetas, uint64_t first_item, uint64_t n_items,
                     const uint64_t* order, double* sum, double* sum_sq);

// the word after the completion flag in mapped host memory: set by peer_allreduce() on a timeout
static int peer_status(Engine& e) {
    volatile unsigned int* st = reinterpret_cast<unsigned int*>(e.h_out + kHostOutWords + 1);
    if (*st) {
        set_error("peer exchange timed out: a rank of the group did not reach this frontier call");
        return PF_ERR_NCCL;
    }
    return PF_OK;
}

// Node-split (small data sets): the data set is replicated on every GPU, rank r scores nodes
// [r*per, (r+1)*per) with the ordinary single-GPU kernel, and one ncclAllGather of 2*per doubles per rank hands
// every rank every node's sums.  No data-path reduction: a node's sums come from exactly one GPU.
// node split through the in-kernel gather: peers attached, one record holds the frontier, nobody is idle, and the
// model runs a frontier kernel with a finalizing CTA (not the data-free scalar Boltzmann)
static bool node_split_fusable(const Engine& e, uint32_t B) {
    const uint32_t G = (uint32_t) e.nranks, per = (B + G - 1) / G;
    return e.peer_fused && e.peer_attached && B <= (uint32_t) kMaxNodesPerPass && (G - 1) * per < B &&
           per <= e.max_frontier && e.desc.model != PF_MODEL_BOLTZMANN_SCALAR;
}

static int eval_host_node_split(Engine& e, uint32_t B, const double* thetas, uint64_t first_item, uint64_t n_items,
                                double* sum, double* sum_sq) {
    PF_CUDA(cudaSetDevice(e.device));
    resident_stop_device(e.device);
    const uint32_t G = (uint32_t) e.nranks, per = (B + G - 1) / G;
    if (per > kMaxNodesPerPass) {
        set_error("node split: %u nodes per rank exceed one pass (%d)", per, kMaxNodesPerPass);
        return PF_ERR_SHAPE;
    }
    const uint32_t lo = (uint32_t) e.rank * per;
    const uint32_t nloc = lo >= B ? 0 : (B - lo < per ? B - lo : per);
    if (node_split_fusable(e, B)) {  // (the data set is replicated: an empty range is empty on every rank)
        // every rank has at least one node and the whole frontier fits one record: the frontier kernel of my slice
        // stores its sums into every peer's exchange buffer and gathers everybody's itself (PeerXchg gather mode):
        // one kernel, results + completion flag in mapped host memory
        if (e.desc.model == PF_MODEL_NORMAL) e.normal_general = !all_identity_cov(e, thetas + (size_t) lo * e.theta_len, nloc);
        const size_t tbytes = (size_t) nloc * e.theta_len * 8;
        const double* src = thetas + (size_t) lo * e.theta_len;
        if (!(src >= e.h_theta && src + (size_t) nloc * e.theta_len <= e.h_theta + (size_t) e.max_frontier * e.theta_len)) {
            memcpy(e.h_theta, src, tbytes);
            src = e.h_theta;
        }
        PF_CUDA(cudaMemcpyAsync(e.d_theta, src, tbytes, cudaMemcpyHostToDevice, e.stream));
        e.done_seq = next_host_seq(e);
        e.done_flag = reinterpret_cast<unsigned long long*>(e.d_hout);
        e.gather_per = per;
        e.gather_total = B;
        int r = launch_model(e, nloc, e.d_theta, e.d_hout, first_item, n_items, e.stream);
        e.gather_per = e.gather_total = 0;
        e.done_flag = nullptr;
        if (r != PF_OK) return r;
        double vals[2 * kMaxNodesPerPass];
        r = wait_words(e, 2 * B, e.done_seq, vals);
        if (r != PF_OK) return r;
        if (peer_status(e) != PF_OK) return PF_ERR_NCCL;
        memcpy(sum, vals, (size_t) B * 8);
        memcpy(sum_sq, vals + B, (size_t) B * 8);
        return PF_OK;
    }
    NcclApi* n = nccl_api();
    if (!n || !n->AllGather || !e.nccl_comm) {
        set_error("node split of %u nodes over %d ranks needs the NCCL communicator (the in-kernel gather covers "
                  "nranks <= B <= %d)", B, e.nranks, kMaxNodesPerPass);
        return PF_ERR_NCCL;
    }
    std::vector<double> s(per, 0.0), q(per, 0.0);
    if (nloc) {
        const bool was = e.node_split;
        e.node_split = false;  // plain single-GPU evaluation of my slice (no collective inside)
        const int saved = e.nranks;
        e.nranks = 1;
        const int r = eval_host(e, nloc, thetas + (size_t) lo * e.theta_len, first_item, n_items, nullptr, s.data(), q.data());
        e.nranks = saved;
        e.node_split = was;
        if (r != PF_OK) return r;
    }
    // pack [sum | sum_sq] of my slice, gather everybody's
    std::vector<double> pack(2 * (size_t) per);
    memcpy(pack.data(), s.data(), per * 8);
    memcpy(pack.data() + per, q.data(), per * 8);
    PF_CUDA(cudaMemcpyAsync(e.d_ns_send, pack.data(), 2 * (size_t) per * 8, cudaMemcpyHostToDevice, e.stream));
    ncclResult_t rc = n->AllGather(e.d_ns_send, e.d_ns_recv, 2 * (size_t) per, /*ncclDouble*/ 8, (ncclComm_t) e.nccl_comm,
                                   e.stream);
    if (rc != 0) {
        set_error("ncclAllGather: %s", n->GetErrorString ? n->GetErrorString(rc) : "error");
        return PF_ERR_NCCL;
    }
    std::vector<double> all(2 * (size_t) per * G);
    PF_CUDA(cudaMemcpyAsync(all.data(), e.d_ns_recv, all.size() * 8, cudaMemcpyDeviceToHost, e.stream));
    PF_CUDA(cudaStreamSynchronize(e.stream));
    for (uint32_t b = 0; b < B; ++b) {
        const uint32_t r = b / per, i = b % per;
        sum[b] = all[(size_t) r * 2 * per + i];
        sum_sq[b] = all[(size_t) r * 2 * per + per + i];
    }
    return PF_OK;
}

// PF_TRACE=1: ns since CTA 0 entered the kernel (launched kernels) or saw the command (resident kernel)
static void print_trace(Engine& e) {
    if (!e.res.alive)
        cudaStreamSynchronize(e.stream);
    else {  // a resident kernel stays on the stream; its last marks land within microseconds of the result words
        timespec ts = {0, 100000};
        nanosleep(&ts, nullptr);
    }
    const unsigned long long* t = e.h_trace;
    fprintf(stderr, "[pf trace] cta0: prologue %lld first-data %lld scored %lld units-done %lld published %lld | "
            "last cta: enter %lld folded %lld end %lld (ns)\n", (long long) (t[1] - t[0]), (long long) (t[2] - t[0]),
            (long long) (t[3] - t[0]), (long long) (t[4] - t[0]), (long long) (t[5] ? t[5] - t[0] : 0),
            (long long) (t[16] - t[0]), (long long) (t[17] - t[0]), (long long) (t[18] - t[0]));
    if (t[19]) fprintf(stderr, "[pf trace] finalize: partials staged %lld, first round of warp 0 %lld, warp 0 done %lld (ns after entering)\n",
                       (long long) (t[19] - t[16]), (long long) (t[20] - t[16]), (long long) (t[21] - t[16]));
    if (t[22]) fprintf(stderr, "[pf trace] finalize, first round of warp 0: prefix sums %lld, records summed %lld, cut items stitched %lld (ns after entering)\n",
                       (long long) (t[22] - t[16]), (long long) (t[23] - t[16]), (long long) (t[24] - t[16]));
    if (t[6]) fprintf(stderr, "[pf trace] mixture exponent form: %s\n", t[6] == 1 ? "factored" : "clamped distance");
    if (t[7]) fprintf(stderr, "[pf trace] resident kernel: command relayed %lld ns after CTA 0 saw it\n", (long long) (t[7] - t[0]));
    std::vector<long long> in, out;
    for (int c = 0; c < kMaxGrid; ++c)
        if (t[32 + 2 * c]) {
            in.push_back((long long) (t[32 + 2 * c] - t[0]));
            out.push_back((long long) (t[33 + 2 * c] - t[0]));
        }
    if (!in.empty()) {
        std::sort(in.begin(), in.end());
        std::sort(out.begin(), out.end());
        fprintf(stderr, "[pf trace] %zu CTAs: entry min/med/max %lld/%lld/%lld, published min/med/max %lld/%lld/%lld (ns)\n",
                in.size(), in.front(), in[in.size() / 2], in.back(), out.front(), out[out.size() / 2], out.back());
    }
    if (const char* tr = getenv("PF_TRACE"); tr && tr[0] == '1' && tr[1] == '1') {  // PF_TRACE=11: every CTA
#ifdef PF_TRACE_SMID
        fprintf(stderr, "[pf trace] SM of every CTA:");
        for (int c = 0; c < kMaxGrid; ++c)
            if (t[32 + 2 * c]) fprintf(stderr, " %u", (unsigned) (t[32 + 2 * c] & 0xFFull));
        fprintf(stderr, "\n");
#endif
        fprintf(stderr, "[pf trace] published(us) by CTA:");
        for (int c = 0; c < kMaxGrid; ++c)
            if (t[32 + 2 * c]) fprintf(stderr, " %.1f", (double) (long long) (t[33 + 2 * c] - t[0]) * 1e-3);
        fprintf(stderr, "\n");
    }
#ifdef PF_TC_TIMELINE
    if (t[400]) {
        fprintf(stderr, "[pf timeline] step: tma-issued data-landed(split starts) mma-may-issue mma-issued (ns since CTA 0 entry)\n");
        for (int q = 0; q < 200 && t[400 + q]; ++q)
            fprintf(stderr, "[pf timeline] %3d: %7lld %7lld %7lld %7lld\n", q, (long long) (t[400 + q] - t[0]), (long long) (t[600 + q] - t[0]),
                    (long long) (t[800 + q] - t[0]), (long long) (t[1000 + q] - t[0]));
    }
#endif
    memset(e.h_trace, 0, (32 + 2 * kMaxGrid) * sizeof(unsigned long long));
}

// ---- resident path: the whole data set, pointwise models, one GPU, small thetas (see pf_resident.cu)
static bool resident_ready(Engine& e) {
    Resident& r = e.res;
    if (!r.enabled || r.state == 0) return false;
    if (r.state < 0) {
        r.state = 0;
        if (const char* v = getenv("PF_RESIDENT"); v && v[0] == '0') {
            r.enabled = false;
            return false;
        }
        if (const char* v = getenv("PF_RESIDENT_IDLE_US"); v && atof(v) > 0.0) r.idle_ns = (unsigned long long) (atof(v) * 1e3);
        if (cudaSetDevice(e.device) != cudaSuccess) return false;
        if (e.desc.model == PF_MODEL_MIXTURE && !e.f32 && pointwise_table_ptrs(e, &r.exp2_rep, &r.log_mid) != 0) return false;
        if (resident_configure(e) != 1) return false;
        if (cudaHostAlloc(&r.h_words, kResHostWords * 8, cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer(&r.d_hcmd, r.h_words, 0) != cudaSuccess ||
            cudaMalloc(&r.d_relay, (kResRelayDoubles) * 8) != cudaSuccess || cudaMalloc(&r.d_ticket, 64) != cudaSuccess ||
            cudaMemset(r.d_relay, 0, kResRelayDoubles * 8) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        memset(r.h_words, 0, kResHostWords * 8);
        r.state = 1;
    }
    return true;
}

static int resident_start(Engine& e, unsigned long long first_seq) {
    Resident& r = e.res;
    Engine* o = g_res_owner[e.device & 63];
    if (o && o != &e) resident_stop(*o);
    if (r.alive && !resident_exited(r)) resident_stop(e);
    ++r.epoch;
    if (resident_launch(e, first_seq) < 0) return PF_ERR_CUDA;
    r.alive = true;
    ++r.launches;
    ++e.launches;
    e.last_stream = e.stream;
    g_res_owner[e.device & 63] = &e;
    return PF_OK;
}

// one command of nb <= kResMaxNodes nodes
static int eval_resident(Engine& e, uint32_t nb, const double* th, double* sum, double* sum_sq) {
    Resident& r = e.res;
    const unsigned long long seq = next_host_seq(e), tag = (seq & 0xFFFFFFFFull) << 32;
    if (!r.alive || resident_exited(r) || g_res_owner[e.device & 63] != &e) {
        const int rc = resident_start(e, seq);
        if (rc != PF_OK) return rc;
    }
    volatile unsigned long long* w = r.h_words;
    const unsigned n = nb * e.theta_len;
    for (unsigned i = 0; i < n; ++i) {
        unsigned long long bits;
        memcpy(&bits, &th[i], 8);
        w[1 + 2 * i] = tag | (bits & 0xFFFFFFFFull);
        w[2 + 2 * i] = tag | (bits >> 32);
    }
    unsigned hdr = 1u | (nb << 4);
    if (e.desc.model == PF_MODEL_NORMAL && !all_identity_cov(e, th, nb)) hdr |= 1u << 12;
    __atomic_store_n(&r.h_words[0], tag | hdr, __ATOMIC_RELEASE);  // the header last: one poll round finds everything
    ++r.calls;

    // results: the same self-validating words as the launched host path
    volatile unsigned long long* o = reinterpret_cast<volatile unsigned long long*>(e.h_out);
    const unsigned long long want = seq & 0xFFFFFFFFull;
    double vals[2 * kMaxNodesPerPass];
    unsigned long long spins = 0;
    int relaunches = 0;
    timespec t_start = {0, 0};
    for (unsigned i = 0; i < 2 * nb; ++i) {
        for (;;) {
            const unsigned long long lo = o[2 * i], hi = o[2 * i + 1];
            if ((lo >> 32) == want && (hi >> 32) == want) {
                const unsigned long long bits = (hi << 32) | (lo & 0xFFFFFFFFull);
                memcpy(&vals[i], &bits, 8);
                break;
            }
            __builtin_ia32_pause();
            if ((++spins & 0x3FF) != 0) continue;
            if (resident_exited(r)) {
                // it left (idle timeout, or stopped) -- before it saw this command, unless the words are here by now
                const unsigned long long lo2 = o[2 * (2 * nb - 1)], hi2 = o[2 * (2 * nb - 1) + 1];
                if ((lo2 >> 32) == want && (hi2 >> 32) == want) continue;
                const unsigned reason = (unsigned) ((r.h_words[kResStateWord] >> 8) & 0xFF);
                if (reason == 3 || ++relaunches > 3) {
                    set_error("resident kernel left without serving the call (reason %u)", reason);
                    r.alive = false;
                    return PF_ERR_CUDA;
                }
                const int rc = resident_start(e, seq);  // the command words are still in place: it picks them up
                if (rc != PF_OK) return rc;
                continue;
            }
            if ((spins & 0x3FFFF) == 0) {
                const cudaError_t q = cudaStreamQuery(e.stream);
                if (q != cudaErrorNotReady && q != cudaSuccess) {
                    set_error("stream error while waiting for the resident kernel: %s", cudaGetErrorString(q));
                    r.alive = false;
                    return PF_ERR_CUDA;
                }
                timespec now;
                clock_gettime(CLOCK_MONOTONIC, &now);
                if (t_start.tv_sec == 0 && t_start.tv_nsec == 0) t_start = now;
                if ((double) (now.tv_sec - t_start.tv_sec) + 1e-9 * (double) (now.tv_nsec - t_start.tv_nsec) > host_timeout_s()) {
                    e.wedged = true;
                    set_error("the resident kernel delivered nothing within %.0f s (PF_HOST_TIMEOUT_S): engine marked unusable",
                              host_timeout_s());
                    return PF_ERR_CUDA;
                }
            }
        }
    }
    memcpy(sum, vals, (size_t) nb * 8);
    memcpy(sum_sq, vals + nb, (size_t) nb * 8);
    return PF_OK;
}

static int eval_host(Engine& e, uint32_t B, const double* thetas, uint64_t first_item, uint64_t n_items,
                     const uint64_t* order, double* sum, double* sum_sq) {
    if (e.wedged) {
        set_error("engine unusable: an earlier frontier kernel never completed");
        return PF_ERR_CUDA;
    }
    if (e.node_split && e.nranks > 1 && !order) return eval_host_node_split(e, B, thetas, first_item, n_items, sum, sum_sq);
    if (e.nranks > 1 && e.desc.model == PF_MODEL_BOLTZMANN_SCALAR) {
        set_error("the scalar Boltzmann model has no data to shard: not available on an engine group");
        return PF_ERR_UNSUPPORTED;
    }
    const bool empty = n_items == 0 || first_item * e.item_rows >= e.n_rows;
    if (empty && e.nranks <= 1) {  // nothing to score and nobody to exchange with
        for (uint32_t b = 0; b < B; ++b) sum[b] = sum_sq[b] = 0.0;
        return PF_OK;
    }
    PF_CUDA(cudaSetDevice(e.device));
    if (!order && e.nranks <= 1 && first_item == 0 && n_items == e.n_items && e.theta_len <= (uint32_t) kResThetaDoubles &&
        (e.desc.model == PF_MODEL_NORMAL || e.desc.model == PF_MODEL_MIXTURE) && resident_ready(e)) {
        uint32_t per = (uint32_t) kResThetaDoubles / e.theta_len;  // nodes per command
        if (per > e.res.max_nodes) per = e.res.max_nodes;
        for (uint32_t b0 = 0; b0 < B; b0 += per) {
            const uint32_t nb = B - b0 < per ? B - b0 : per;
            const int r = eval_resident(e, nb, thetas + (size_t) b0 * e.theta_len, sum + b0, sum_sq + b0);
            if (r != PF_OK) return r;
        }
        if (e.h_trace) print_trace(e);
        return PF_OK;
    }
    resident_stop_device(e.device);  // everything below launches kernels
    for (uint32_t b0 = 0; b0 < B; b0 += e.max_frontier) {
        const uint32_t nb = B - b0 < e.max_frontier ? B - b0 : e.max_frontier;
        const double* th = thetas + (size_t) b0 * e.theta_len;
        if (e.desc.model == PF_MODEL_NORMAL) e.normal_general = !all_identity_cov(e, th, nb);
        const size_t tbytes = (size_t) nb * e.theta_len * 8;
        // Fast path: one frontier kernel whose last CTA writes the 2*nb results into mapped host memory and then
        // publishes a sequence number.  Row-sharded groups reduce inside the same kernel (peer_allreduce).  Not for NCCL-reduced results, ordered ranges, the scalar model or empty ranges.
        // Whether a rank is on this path must not depend on its LOCAL range: a rank of a peer group whose range is
        // empty still launches (launch_exchange_only) so that its peers' kernels find its record.
        const bool fast = !order && (e.nranks <= 1 || peer_active(e)) && e.desc.model != PF_MODEL_BOLTZMANN_SCALAR;
        // Small frontiers of the pointwise models ride in the kernel parameters (PointwiseArgs::theta_inline); anything
        // else goes through a DMA copy (letting ~300 CTAs read theta from mapped host memory was measured SLOWER,
        // 56 vs 37 us per call for the normal target -- hundreds of small PCIe reads in front of the compute).
        const bool inline_theta = fast && e.allow_inline_theta && (e.desc.model == PF_MODEL_NORMAL || e.desc.model == PF_MODEL_MIXTURE) &&
                                  (size_t) nb * e.theta_len <= (size_t) kInlineThetaDoubles;
        if (!inline_theta) {
            // through the engine's pinned staging buffer -- unless the caller already wrote the frontier there
            // (pf_engine_theta_buffer): then the DMA reads it in place and the host-side copy (25-40 us for the 512 KB of
            // a dense Boltzmann frontier) is saved
            const double* src = th;
            if (!(th >= e.h_theta && th + (size_t) nb * e.theta_len <= e.h_theta + (size_t) e.max_frontier * e.theta_len)) {
                memcpy(e.h_theta, th, tbytes);
                src = e.h_theta;
            }
            PF_CUDA(cudaMemcpyAsync(e.d_theta, src, tbytes, cudaMemcpyHostToDevice, e.stream));
        }
        int r;
        if (fast) {
            e.inline_theta = inline_theta ? th : nullptr;
            e.inline_theta_n = inline_theta ? (unsigned) (nb * e.theta_len) : 0;
            e.done_seq = next_host_seq(e);
            e.done_flag = reinterpret_cast<unsigned long long*>(e.d_hout);
            r = launch_model(e, nb, e.d_theta, e.d_hout, first_item, n_items, e.stream);
            e.done_flag = nullptr;
            e.inline_theta = nullptr;
            e.inline_theta_n = 0;
            if (r != PF_OK) return r;
            double vals[2 * kMaxNodesPerPass];
            r = wait_words(e, 2 * nb, e.done_seq, vals);
            if (r != PF_OK) return r;
            if (peer_status(e) != PF_OK) return PF_ERR_NCCL;
            if (e.h_trace) print_trace(e);
            memcpy(sum + b0, vals, (size_t) nb * 8);
            memcpy(sum_sq + b0, vals + nb, (size_t) nb * 8);
            continue;
        }
        if (order) {  // the reference's per-node visiting order (normal target; checked by the caller)
            PF_CUDA(cudaMemcpyAsync(e.d_order, order + (size_t) 2 * b0, (size_t) nb * 2 * sizeof(uint64_t),
                                    cudaMemcpyHostToDevice, e.stream));
            const int k = launch_normal_ordered(e, nb, e.d_theta, e.d_out, e.d_order, first_item, n_items, e.stream);
            if (k < 0) return PF_ERR_CUDA;
            e.launches += (uint64_t) k;
            r = PF_OK;
        } else
            r = launch_model(e, nb, e.d_theta, e.d_out, first_item, n_items, e.stream);
        if (r != PF_OK) return r;
        r = allreduce_out(e, e.d_out, nb, e.stream);
        if (r != PF_OK) return r;
        double* stage = e.h_out + kHostOutWords + 8;  // behind the result words and the status word
        PF_CUDA(cudaMemcpyAsync(stage, e.d_out, (size_t) 2 * nb * 8, cudaMemcpyDeviceToHost, e.stream));
        PF_CUDA(cudaStreamSynchronize(e.stream));
        memcpy(sum + b0, stage, (size_t) nb * 8);
        memcpy(sum_sq + b0, stage + nb, (size_t) nb * 8);
    }
    return PF_OK;
}

}  // namespace pf

// ============================================================================================
// C ABI
// ============================================================================================
using pf::Engine;

extern "C" {

int pf_abi_version(void) { return PF_ABI_VERSION; }

const char* pf_last_error(void) { return pf::g_err; }

int pf_engine_create(const pf_model_desc* desc, pf_engine** out) {
    Engine* e = nullptr;
    int r = pf::create(desc, &e);
    if (out) *out = reinterpret_cast<pf_engine*>(e);
    return r;
}

void pf_engine_destroy(pf_engine* e) { pf::destroy(reinterpret_cast<Engine*>(e)); }

uint64_t pf_engine_data_size(const pf_engine* e) { return e ? reinterpret_cast<const Engine*>(e)->n_items : 0; }

uint32_t pf_engine_theta_len(const pf_engine* e) { return e ? reinterpret_cast<const Engine*>(e)->theta_len : 0; }

uint32_t pf_engine_max_frontier(const pf_engine* e) {
    return e ? reinterpret_cast<const Engine*>(e)->max_frontier : 0;
}

int pf_engine_model(const pf_engine* e) { return e ? reinterpret_cast<const Engine*>(e)->desc.model : 0; }

uint32_t pf_engine_dim(const pf_engine* e) { return e ? reinterpret_cast<const Engine*>(e)->dim : 0; }

uint32_t pf_engine_components(const pf_engine* e) { return e ? reinterpret_cast<const Engine*>(e)->K : 0; }

int pf_engine_group(const pf_engine* pe, int* rank, int* nranks, int* node_split) {
    if (!pe) return PF_ERR_INVALID;
    const Engine& e = *reinterpret_cast<const Engine*>(pe);
    if (rank) *rank = e.rank;
    if (nranks) *nranks = e.nranks;
    if (node_split) *node_split = e.node_split ? 1 : 0;
    return PF_OK;
}

double* pf_engine_theta_buffer(pf_engine* pe, uint64_t* capacity_doubles) {
    if (!pe) return nullptr;
    Engine& e = *reinterpret_cast<Engine*>(pe);
    if (capacity_doubles) *capacity_doubles = (uint64_t) e.max_frontier * e.theta_len;
    return e.h_theta;
}

uint64_t pf_engine_launch_count(const pf_engine* e) { return e ? reinterpret_cast<const Engine*>(e)->launches : 0; }

int pf_engine_resident_stats(const pf_engine* pe, uint64_t* calls, uint64_t* launches) {
    if (!pe) return 0;
    const Engine& e = *reinterpret_cast<const Engine*>(pe);
    if (calls) *calls = e.res.calls;
    if (launches) *launches = e.res.launches;
    return e.res.state == 1 && e.res.enabled ? 1 : 0;
}

int pf_engine_set_option(pf_engine* pe, int option, int64_t value) {
    if (!pe) return PF_ERR_INVALID;
    Engine& e = *reinterpret_cast<Engine*>(pe);
    switch (option) {
        case PF_OPT_ASSUME_IDENTITY_COV: e.normal_general = value == 0; e.assume_identity = value != 0; return PF_OK;
        case PF_OPT_NODE_SPLIT: e.node_split = value != 0; return PF_OK;
        case PF_OPT_MATVEC_TENSOR: e.mv_tensor = value != 0; e.mv_tensor_force = value == 2; return PF_OK;
        case PF_OPT_INLINE_THETA: e.allow_inline_theta = value != 0; return PF_OK;
        case PF_OPT_RESIDENT:
            e.res.enabled = value != 0;
            if (!e.res.enabled) pf::resident_stop(e);
            return PF_OK;
        case PF_OPT_PEER_EXCHANGE:
            if (value != 0 && !e.peer_attached) {
                pf::set_error("PF_OPT_PEER_EXCHANGE: no peer buffers attached (pf_engine_peer_attach)");
                return PF_ERR_INVALID;
            }
            if (value == 0 && e.nranks > 1 && !e.nccl_comm) {
                pf::set_error("PF_OPT_PEER_EXCHANGE = 0 needs the NCCL communicator (pf_engine_comm_init)");
                return PF_ERR_INVALID;
            }
            e.peer_fused = value != 0;
            return PF_OK;
    }
    pf::set_error("unknown option %d", option);
    return PF_ERR_INVALID;
}

int pf_eval_frontier(pf_engine* pe, uint32_t B, const double* thetas, double* sum, double* sum_sq) {
    if (!pe || !thetas || !sum || !sum_sq || B == 0) {
        pf::set_error("pf_eval_frontier: null argument or B == 0");
        return PF_ERR_INVALID;
    }
    Engine& e = *reinterpret_cast<Engine*>(pe);
    return pf::eval_host(e, B, thetas, 0, e.n_items, nullptr, sum, sum_sq);
}

int pf_eval_frontier_repeat(pf_engine* pe, uint32_t B, const double* thetas, double* sum, double* sum_sq, uint32_t iters,
                            double* seconds) {
    if (!pe || !thetas || !sum || !sum_sq || B == 0 || iters == 0) {
        pf::set_error("pf_eval_frontier_repeat: null argument, B == 0 or iters == 0");
        return PF_ERR_INVALID;
    }
    Engine& e = *reinterpret_cast<Engine*>(pe);
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (uint32_t i = 0; i < iters; ++i) {
        const int r = pf::eval_host(e, B, thetas, 0, e.n_items, nullptr, sum, sum_sq);
        if (r != PF_OK) return r;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
    return PF_OK;
}

int pf_eval_frontier_range(pf_engine* pe, uint32_t B, const double* thetas, uint64_t first_item, uint64_t n_items,
                           const uint64_t* order, double* sum, double* sum_sq) {
    if (!pe || !thetas || !sum || !sum_sq || B == 0) {
        pf::set_error("pf_eval_frontier_range: null argument or B == 0");
        return PF_ERR_INVALID;
    }
    Engine& e = *reinterpret_cast<Engine*>(pe);
    if (first_item > e.n_items || n_items > e.n_items - first_item) {
        pf::set_error("item range [%llu, +%llu) outside [0, %llu)", (unsigned long long) first_item,
                      (unsigned long long) n_items, (unsigned long long) e.n_items);
        return PF_ERR_SHAPE;
    }
    if (order) {
        if (e.desc.model != PF_MODEL_NORMAL) {
            pf::set_error("a visiting order only exists for the normal target (normalsampler.hh:115-119)");
            return PF_ERR_UNSUPPORTED;
        }
        if (e.nranks > 1) {
            pf::set_error("ordered ranges address the whole data set: not available on a row shard");
            return PF_ERR_UNSUPPORTED;
        }
        for (uint32_t b = 0; b < B; ++b)
            if (order[2 * b] == 0 || order[2 * b] >= e.n_rows || order[2 * b + 1] >= e.n_rows) {
                if (!(e.n_rows == 1 && order[2 * b] == 1)) {
                    pf::set_error("order[%u]: stride/start outside [0, n)", b);
                    return PF_ERR_SHAPE;
                }
            }
    }
    if (e.desc.model == PF_MODEL_BOLTZMANN_DENSE && n_items != 1) {
        for (uint32_t b = 0; b < B; ++b) sum[b] = sum_sq[b] = 0.0;
        return PF_OK;
    }
    return pf::eval_host(e, B, thetas, first_item, n_items, order, sum, sum_sq);
}

int pf_eval_frontier_device(pf_engine* pe, uint32_t B, const double* d_thetas, double* d_out, void* cuda_stream) {
    if (!pe || !d_thetas || !d_out || B == 0) {
        pf::set_error("pf_eval_frontier_device: null argument or B == 0");
        return PF_ERR_INVALID;
    }
    Engine& e = *reinterpret_cast<Engine*>(pe);
    if (B > e.max_frontier) {
        pf::set_error("B = %u exceeds max_frontier = %u for the device-pointer path", B, e.max_frontier);
        return PF_ERR_SHAPE;
    }
    if (cudaSetDevice(e.device) != cudaSuccess) return PF_ERR_CUDA;
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e.stream;
    if (e.desc.model == PF_MODEL_NORMAL) e.normal_general = !e.assume_identity;
    if (e.node_split && e.nranks > 1) {
        if (!pf::node_split_fusable(e, B)) {
            pf::set_error("node split on the device-pointer path needs the peer exchange and nranks <= B <= %d",
                          pf::kMaxNodesPerPass);
            return PF_ERR_UNSUPPORTED;
        }
        const uint32_t G = (uint32_t) e.nranks, per = (B + G - 1) / G, lo = (uint32_t) e.rank * per;
        const uint32_t nloc = B - lo < per ? B - lo : per;
        e.gather_per = per;
        e.gather_total = B;
        const int rr = pf::launch_model(e, nloc, d_thetas + (size_t) lo * e.theta_len, d_out, 0, e.n_items, s);
        e.gather_per = e.gather_total = 0;
        return rr;
    }
    int r = pf::launch_model(e, B, d_thetas, d_out, 0, e.n_items, s);
    if (r != PF_OK) return r;
    return pf::allreduce_out(e, d_out, B, s);
}

double pf_blasso_log_prior(const double* theta, uint32_t p, double tau) {
    // pyblasso.py:123-133, same operation order
    const double sigma = theta[0];
    double l1 = 0.0;
    for (uint32_t j = 0; j < p; ++j) l1 += fabs(theta[1 + j]);
    const double log_prior_sigma = -2.0 * log(sigma);
    const double log_laplace = p * log(tau / 2 / sigma) - (tau * l1) / sigma;
    return log_prior_sigma + log_laplace;
}

int pf_nccl_unique_id(void* id128) {
    if (!id128) return PF_ERR_INVALID;
    pf::NcclApi* n = pf::nccl_api();
    if (!n) {
        pf::set_error("NCCL library not loadable");
        return PF_ERR_NCCL;
    }
    ncclUniqueId id;
    if (n->GetUniqueId(&id) != 0) {
        pf::set_error("ncclGetUniqueId failed");
        return PF_ERR_NCCL;
    }
    memcpy(id128, &id, 128);
    return PF_OK;
}

int pf_engine_comm_init(pf_engine* pe, const void* id128, int rank, int nranks) {
    if (!pe || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return PF_ERR_INVALID;
    Engine& e = *reinterpret_cast<Engine*>(pe);
    if (e.nccl_comm) {
        pf::set_error("pf_engine_comm_init: this engine already has a communicator (create a new engine to regroup)");
        return PF_ERR_INVALID;
    }
    if (e.peer_attached && (rank != e.rank || nranks != e.nranks)) {
        pf::set_error("pf_engine_comm_init: rank %d of %d contradicts the attached peer group (rank %d of %d)", rank, nranks,
                      e.rank, e.nranks);
        return PF_ERR_INVALID;
    }
    if (nranks > 1 && e.desc.model == PF_MODEL_BOLTZMANN_SCALAR) {
        pf::set_error("the scalar Boltzmann model has no data to shard: not available on an engine group");
        return PF_ERR_UNSUPPORTED;
    }
    e.rank = rank;
    e.nranks = nranks;
    if (nranks == 1) return PF_OK;
    pf::NcclApi* n = pf::nccl_api();
    if (!n) {
        pf::set_error("NCCL library not loadable");
        return PF_ERR_NCCL;
    }
    if (cudaSetDevice(e.device) != cudaSuccess) return PF_ERR_CUDA;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    ncclResult_t r = n->CommInitRank(&comm, nranks, id, rank);
    if (r != 0) {
        pf::set_error("ncclCommInitRank: %s", n->GetErrorString ? n->GetErrorString(r) : "error");
        return PF_ERR_NCCL;
    }
    e.nccl_comm = comm;
    if (cudaMalloc(&e.d_ns_send, 2 * pf::kMaxNodesPerPass * 8) != cudaSuccess ||
        cudaMalloc(&e.d_ns_recv, (size_t) nranks * 2 * pf::kMaxNodesPerPass * 8) != cudaSuccess)
        return PF_ERR_NOMEM;
    return PF_OK;
}

static int peer_alloc(Engine& e) {
    if (e.d_xchg) return PF_OK;
    if (cudaSetDevice(e.device) != cudaSuccess) return PF_ERR_CUDA;
    const size_t bytes = (size_t) pf::kXchgSlots * pf::kXchgMaxRanks * pf::kXchgRec * 8;
    // a plain cudaMalloc allocation of its own: cudaIpcGetMemHandle exports whole allocations
    if (cudaMalloc(&e.d_xchg, bytes) != cudaSuccess || cudaMemset(e.d_xchg, 0, bytes) != cudaSuccess ||
        cudaMalloc(&e.d_peers, pf::kXchgMaxRanks * sizeof(double*)) != cudaSuccess) {
        pf::set_error("peer exchange: allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        return PF_ERR_NOMEM;
    }
    return PF_OK;
}

static int peer_finish(Engine& e, double* const* table, int rank, int nranks) {
    if (cudaMemcpy(e.d_peers, table, (size_t) nranks * sizeof(double*), cudaMemcpyHostToDevice) != cudaSuccess) {
        pf::set_error("peer exchange: pointer table upload failed");
        return PF_ERR_CUDA;
    }
    e.rank = rank;
    e.nranks = nranks;
    e.xchg_seq = 0;
    e.peer_attached = true;
    e.peer_fused = true;
    return PF_OK;
}

int pf_engine_peer_export(pf_engine* pe, void* handle64) {
    if (!pe || !handle64) return PF_ERR_INVALID;
    Engine& e = *reinterpret_cast<Engine*>(pe);
    int r = peer_alloc(e);
    if (r != PF_OK) return r;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, e.d_xchg) != cudaSuccess) {
        pf::set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(cudaGetLastError()));
        return PF_ERR_CUDA;
    }
    memcpy(handle64, &h, 64);
    return PF_OK;
}

int pf_engine_peer_attach(pf_engine* pe, const void* handles, int rank, int nranks) {
    if (!pe || !handles || nranks < 1 || nranks > pf::kXchgMaxRanks || rank < 0 || rank >= nranks) {
        pf::set_error("pf_engine_peer_attach: 1 <= nranks <= %d, 0 <= rank < nranks", pf::kXchgMaxRanks);
        return PF_ERR_INVALID;
    }
    Engine& e = *reinterpret_cast<Engine*>(pe);
    if (e.peer_attached) {
        // sequence numbers and the flags already sitting in the peers' buffers belong to the first attachment
        pf::set_error("pf_engine_peer_attach: this engine already has its peers (create a new engine to regroup)");
        return PF_ERR_INVALID;
    }
    int r = peer_alloc(e);
    if (r != PF_OK) return r;
    if (cudaSetDevice(e.device) != cudaSuccess) return PF_ERR_CUDA;
    double* table[pf::kXchgMaxRanks] = {};
    for (int k = 0; k < nranks; ++k) {
        if (k == rank) {
            table[k] = e.d_xchg;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + (size_t) k * 64, 64);
        void* p = nullptr;
        // maps the peer's allocation into this process and enables peer access to its device
        cudaError_t ce = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (ce != cudaSuccess) {
            pf::set_error("cudaIpcOpenMemHandle(rank %d): %s", k, cudaGetErrorString(ce));
            cudaGetLastError();
            for (int j = 0; j < k; ++j)  // leave nothing half-attached
                if (e.peer_mapped[j]) {
                    cudaIpcCloseMemHandle(e.peer_mapped[j]);
                    e.peer_mapped[j] = nullptr;
                }
            return PF_ERR_CUDA;
        }
        e.peer_mapped[k] = p;
        table[k] = static_cast<double*>(p);
    }
    return peer_finish(e, table, rank, nranks);
}

int pf_engine_peer_attach_local(pf_engine* const* engines, int nranks) {
    if (!engines || nranks < 1 || nranks > pf::kXchgMaxRanks) return PF_ERR_INVALID;
    double* table[pf::kXchgMaxRanks] = {};
    for (int k = 0; k < nranks; ++k) {
        if (!engines[k]) return PF_ERR_INVALID;
        Engine& e = *reinterpret_cast<Engine*>(engines[k]);
        if (e.peer_attached) {
            pf::set_error("pf_engine_peer_attach_local: engine %d already has its peers", k);
            return PF_ERR_INVALID;
        }
        int r = peer_alloc(e);
        if (r != PF_OK) return r;
        table[k] = e.d_xchg;
    }
    for (int k = 0; k < nranks; ++k) {
        Engine& e = *reinterpret_cast<Engine*>(engines[k]);
        if (cudaSetDevice(e.device) != cudaSuccess) return PF_ERR_CUDA;
        for (int j = 0; j < nranks; ++j) {
            const int dj = reinterpret_cast<Engine*>(engines[j])->device;
            if (dj == e.device) continue;
            cudaError_t ce = cudaDeviceEnablePeerAccess(dj, 0);
            if (ce != cudaSuccess && ce != cudaErrorPeerAccessAlreadyEnabled) {
                pf::set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", e.device, dj, cudaGetErrorString(ce));
                cudaGetLastError();
                return PF_ERR_CUDA;
            }
            cudaGetLastError();
        }
        int r = peer_finish(e, table, k, nranks);
        if (r != PF_OK) return r;
    }
    return PF_OK;
}

int pf_engine_peer_status(pf_engine* pe) {
    if (!pe) return PF_ERR_INVALID;
    return pf::peer_status(*reinterpret_cast<Engine*>(pe));
}

int pf_device_sm_count(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    return prop.multiProcessorCount;
}

}  // extern "C"
