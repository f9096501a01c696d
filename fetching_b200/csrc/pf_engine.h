// pf_engine.h -- internal (not part of the C ABI): engine state and kernel launchers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pfetch.h"

namespace pf {

constexpr int kMaxNodesPerPass = 64;  // frontier nodes scored by one kernel launch
constexpr int kPartSlots = 5;        // per-CTA record: S, Q, head, tail, stitch source (see ItemOwner)
constexpr int kMaxGrid = 148 * 4;    // upper bound on CTAs per launch (sizes the partial buffer)
constexpr int kInlineThetaDoubles = 432;  // keeps PointwiseArgs under the classic 4 KB kernel-parameter limit
constexpr int kHostOutWords = 4 * kMaxNodesPerPass;  // [2][64] doubles as 2 words each (see PointwiseArgs::done_flag)
constexpr int kXchgMaxRanks = 16;    // ranks in one peer-exchange group (one NVSwitch domain)
constexpr int kXchgRec = 2 * kMaxNodesPerPass + 16;  // doubles per (slot, rank) record: S[64], Q[64], flag, pad
constexpr int kXchgSlots = 2;        // records are double-buffered by call parity

// In-kernel exchange of the per-node partials between the GPUs of a row-sharded engine group:
// the finalizing CTA of every rank stores its [2][B] partials straight into every peer's exchange buffer
// (NVLink / NVSwitch peer stores), publishes a sequence flag, waits for the other ranks' flags in its own
// buffer and sums the records in rank order -- so every rank gets bit-identical totals and no collective
// kernel, device->host copy or stream synchronisation follows the frontier kernel.  nranks <= 1: disabled.
struct PeerXchg {
    double* const* peers;      // device array [nranks]: base of every rank's exchange buffer (peer-mapped)
    unsigned int* status;      // mapped host word, set to 1 when a peer did not show up in time
    unsigned long long seq;    // call number, identical on every rank
    unsigned int rank, nranks;
    // NODE SPLIT (replicated data): gather instead of sum -- rank r scored nodes [r * gather_per, ...) of a
    // frontier of gather_total nodes; output node b is record (b / gather_per), column (b % gather_per).  0: sum.
    unsigned int gather_per, gather_total;
};

// Geometry of one launch of the pointwise (normal / mixture) kernels.
struct PointwiseArgs {
    const void* data;       // device, row-major [rows][dim] of T, 256-byte aligned, padded
    const double* theta;    // device [B][theta_len]
    double* part;           // device [kPartSlots][64][kMaxGrid]: per-CTA records, transposed (ItemOwner::store)
    double* out;            // device [2][B]
    unsigned int* ticket;   // device, zero between launches
    unsigned long long row_begin, row_end;  // rows to score (whole items)
    unsigned long long item_rows;           // rows per item (1: every row is an item)
    unsigned int rows_per_cta;
    unsigned int B, theta_len, K;
    // host path (done_flag != nullptr): `out` is mapped pinned host memory and the results are delivered as
    // self-validating words -- every double travels as two 8-byte stores {seq32 | low half}, {seq32 | high half}
    // (aligned 8-byte stores are atomic over PCIe) -- so the kernel needs neither a system-scope fence nor a
    // completion flag after them, and the host, polling the words, has each value as soon as it lands
    unsigned long long* done_flag;
    unsigned long long done_seq;
    PeerXchg xchg;
    // host path, small frontiers: the thetas travel INSIDE the kernel parameters (constant bank, broadcast to every
    // CTA) instead of through a host->device copy in front of the kernel (~8 us of a ~27 us call)
    unsigned int out_off, out_stride;     // results go to out[out_off + b] and out[out_stride + out_off + b]
                                          // (out_stride 0: the launch's own node count) -- a frontier scored in two launches
    unsigned long long* trace;            // diagnostics (PF_TRACE=1): %globaltimer marks of CTA 0 [0..15] and of the
                                          // finalizing CTA [16..31]; nullptr in normal operation
    double xmax[8];                       // max |x_j| over the engine's rows per column (+inf: unknown / non-finite data):
                                          // the bounding box the mixture kernel tests theta against (factored form)
    unsigned int theta_inline_n;          // doubles valid in theta_inline; 0: read `theta`
    double theta_inline[kInlineThetaDoubles];
};

static_assert(sizeof(PointwiseArgs) <= 4096, "PointwiseArgs must fit the classic kernel-parameter limit");

// Geometry of one launch of the streamed-matrix (lasso / dense Boltzmann) kernel.
struct MatvecArgs {
    CUtensorMap tmap;       // [rows][cols_pad] matrix, box = 16 doubles (or 32 floats) x ROWS, 128B swizzle
    const double* theta;    // device [B][theta_len]
    const void* betaT;      // device [cols_pad][Bpad] of T: coefficient panel, transposed by the prep kernel
    const double* y;        // device [rows] (lasso) or nullptr
    double* part;
    double* out;
    unsigned int* ticket;
    unsigned long long row_begin, row_end;
    unsigned long long item_rows;
    unsigned int cols;      // p or D
    unsigned int cols_pad;  // multiple of the k-chunk
    unsigned int B, Bpad, theta_len;
    unsigned int n_units;   // work units = row tiles x k splits
    unsigned int k_splits;  // 1 for lasso; >1 only for the (linear) Boltzmann epilogue
    unsigned int chunks_per_split;
    unsigned int tile_rows;
    unsigned long long rows_per_cta;  // lasso: contiguous rows per CTA (multiple of tile_rows)
    double inv_temperature;
    unsigned long long* done_flag;    // see PointwiseArgs
    unsigned long long done_seq;
    PeerXchg xchg;
    unsigned long long* trace;        // see PointwiseArgs
};

// ---- resident frontier kernel (pf_resident.cu): launched once, driven through a doorbell in mapped host memory
constexpr int kResMaxNodes = 32;                              // nodes per command (larger frontiers: several commands)
constexpr int kResThetaDoubles = 512;                         // theta doubles per command (32 nodes of the K = 8, d = 2 mixture)
constexpr int kResCmdWords = 1 + 2 * kResThetaDoubles;        // header + two words per theta double
constexpr int kResCtlWord = 1040;                             // stop request: the epoch of the launch to stop
constexpr int kResStateWord = 1048;                           // written by the kernel when it leaves: epoch << 32 | reason << 8 | 1
constexpr int kResHostWords = 1056;
constexpr int kResRelayDoubles = kResCmdWords + 7;            // relay block in device memory: the command words as CTA 0 forwards them

struct ResidentArgs {
    const void* data;
    double* part;
    double* out;                       // device alias of the mapped result words (Engine::d_hout)
    unsigned long long* ticket;        // device, zero at launch, monotonic afterwards
    const unsigned long long* h_cmd;   // device alias of the mapped command words {tag << 32 | payload}
    const unsigned long long* h_ctl;   // mapped: stop request
    unsigned long long* h_state;       // mapped: exit notice
    unsigned long long* d_relay;       // device: the command words of the current call, forwarded by CTA 0 (same format)
    const double* exp2_rep;            // device tables of the FP64 mixture loop (pf_math.cuh), or nullptr
    const double* log_mid;
    unsigned long long row_end, item_rows;
    unsigned int rows_per_cta, theta_len, K, cap_bytes;
    unsigned long long first_seq, epoch, idle_ns;
    unsigned long long* trace;
    double xmax[8];
};

struct Resident {
    int state = -1;              // -1: not looked at yet, 0: not eligible, 1: eligible (geometry below is valid)
    bool enabled = true;         // PF_OPT_RESIDENT / PF_RESIDENT
    bool alive = false;          // a launch is (as far as the host knows) on the SMs
    const void* kern = nullptr;
    unsigned grid = 0, rows_per_cta = 0, cap_bytes = 0, max_nodes = 32;  // max_nodes: nodes one command may carry
    size_t smem = 0;
    unsigned long long epoch = 0, idle_ns = 2000000ull;
    unsigned long long* h_words = nullptr;  // pinned + mapped: command words, control word, state word
    unsigned long long* d_hcmd = nullptr;   // device alias
    unsigned long long* d_relay = nullptr;
    unsigned long long* d_ticket = nullptr;
    const double* exp2_rep = nullptr;
    const double* log_mid = nullptr;
    uint64_t calls = 0, launches = 0;
};

struct Engine {
    pf_model_desc desc;
    int device = 0;
    int sm_count = 148;
    bool f32 = false;
    uint64_t n_rows = 0, n_items = 0, item_rows = 1;
    uint32_t dim = 0, K = 0, theta_len = 0, max_frontier = 64;
    uint32_t cols_pad = 0;
    size_t pitch_bytes = 0;      // matrix models: bytes between consecutive rows of the resident matrix (>= cols_pad * esz)
    double temperature = 1.0;
    double xmax[8];              // pointwise models: max |x_j| per column of the resident rows (see PointwiseArgs)
    bool normal_general = true;  // per call: some node has a non-identity covariance (or unknown)
    bool assume_identity = false;  // PF_OPT_ASSUME_IDENTITY_COV
    bool mv_tensor = true;         // PF_OPT_MATVEC_TENSOR: tensor-core contraction (FP64: DMMA; PF_F32 lasso: tcgen05 3xTF32)
    bool mv_tensor_force = false;  // PF_OPT_MATVEC_TENSOR = 2: the tcgen05 path for every frontier size (default: B > 32)

    void* d_data = nullptr;      // data set (T), padded
    double* d_resp = nullptr;    // lasso response (always double)
    double* d_theta = nullptr;   // [max_frontier][theta_len]
    double* d_out = nullptr;     // [2][max_frontier]
    double* d_part = nullptr;    // [kPartSlots][64][kMaxGrid]
    void* d_betaT = nullptr;     // matvec coefficient panel
    unsigned long long* d_order = nullptr;  // [max_frontier][2] StrideIndex (stride, start) per node
    unsigned int* d_ticket = nullptr;
    double* h_theta = nullptr;   // pinned + mapped staging
    double* d_htheta = nullptr;  // device alias of h_theta
    double* h_out = nullptr;     // pinned + mapped: kHostOutWords self-validating result words, then the status word
    double* d_hout = nullptr;    // device alias of h_out
    unsigned long long host_seq = 0;
    unsigned long long* done_flag = nullptr;  // per launch: device alias of the flag, or nullptr
    unsigned long long done_seq = 0;
    cudaStream_t stream = nullptr;
    CUtensorMap tmaps[8];        // tensor maps by box height (cached)
    unsigned tmap_rows[8] = {};
    int n_tmaps = 0;

    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;
    bool node_split = false;      // PF_OPT_NODE_SPLIT: replicated data, nodes divided over ranks, all-gather
    double* d_ns_send = nullptr;  // [2][per]
    double* d_ns_recv = nullptr;  // [nranks][2][per]
    // peer exchange (pf_engine_peer_export / _attach): fused all-reduce inside the frontier kernels
    double* d_xchg = nullptr;              // my exchange buffer [kXchgSlots][kXchgMaxRanks][kXchgRec]
    double** d_peers = nullptr;            // device table of every rank's buffer
    void* peer_mapped[kXchgMaxRanks] = {}; // IPC mappings to close (nullptr for my own / same-process buffers)
    bool peer_fused = false;               // the exchange is attached and selected (PF_OPT_PEER_EXCHANGE)
    bool peer_attached = false;
    unsigned long long xchg_seq = 0;
    unsigned gather_per = 0, gather_total = 0;  // set around a node-split launch (see PeerXchg)
    const double* inline_theta = nullptr;        // set around a host-path launch: HOST thetas to pass as parameters
    unsigned inline_theta_n = 0;
    bool allow_inline_theta = true;              // PF_OPT_INLINE_THETA
    unsigned long long* h_trace = nullptr;       // PF_TRACE=1: mapped pinned [32] (+ device alias)
    unsigned long long* d_trace = nullptr;
    PeerXchg next_xchg();                  // arguments for the next launch (advances the call number)
    uint64_t launches = 0;
    // One set of scratch buffers => launches are serialised across streams: `order_event` is recorded behind every
    // launch and waited on by the next launch that goes to a different stream (pf_eval_frontier_device).
    cudaEvent_t order_event = nullptr;
    cudaStream_t last_stream = nullptr;
    bool wedged = false;  // a frontier kernel never completed (wait_words timed out): every later call fails
    Resident res;
};

// launchers (defined in the .cu files); return the number of kernels launched or <0 on error
int launch_pointwise(Engine& e, uint32_t B, const double* d_theta, double* d_out, uint64_t first_item,
                     uint64_t n_items, cudaStream_t s);
int launch_matvec(Engine& e, uint32_t B, const double* d_theta, double* d_out, uint64_t first_item,
                  uint64_t n_items, cudaStream_t s);
int launch_normal_ordered(Engine& e, uint32_t B, const double* d_theta, double* d_out,
                          const unsigned long long* d_order, uint64_t first_item, uint64_t n_items, cudaStream_t s);
int launch_boltzmann_scalar(Engine& e, uint32_t B, const double* d_theta, double* d_out, uint64_t n_items,
                            cudaStream_t s);
// a rank of a peer-exchange group with NO rows in the requested range: takes part in the exchange with zeros
int launch_exchange_only(Engine& e, uint32_t B, double* d_out, cudaStream_t s, unsigned out_off = 0,
                         unsigned out_stride = 0);
// PF_F32 lasso on tcgen05 (pf_matvec_tc.cu): fills the tile geometry of `a`, launches the panel prep + frontier kernels
int launch_matvec_tc32(Engine& e, MatvecArgs& a, const double* d_theta, cudaStream_t s);
int make_tmap(Engine& e, unsigned tile_rows, CUtensorMap* out);  // cached 2-D tensor map of the matrix (pf_matvec.cu)
int resident_configure(Engine& e);                                  // pf_resident.cu: 1 = the data set fits shared memory
int resident_launch(Engine& e, unsigned long long first_seq);       // pf_resident.cu
int pointwise_table_ptrs(Engine& e, const double** exp2_rep, const double** log_mid);  // pf_pointwise.cu
uint32_t pointwise_max_frontier(const Engine& e);
int ensure_exp_table_for(Engine& e);  // uploads the mixture loop's exp table to e.device (once per device)
bool pointwise_supported(const Engine& e);

void set_error(const char* fmt, ...);

}  // namespace pf
