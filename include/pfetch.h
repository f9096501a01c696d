/* include/pfetch.h -- C ABI of the B200 frontier log-likelihood engine ("pfetch").
 *
 * This is the drop-in boundary for the evaluation path of elaine84/fetching.  In the
 * reference every speculative tree node is shipped to an MPI worker, which walks the data
 * set one item at a time through the Executor concept and accumulates two numbers:
 *
 *     worker.hh:148-155        log_prob = executor.evaluate(position);
 *                              metric_sum += log_prob;  metric_sum_sq += log_prob * log_prob;
 *
 * and reports them in a MetricChange (protocol.hh:43-53).  Here the master hands the WHOLE
 * speculative frontier (every JobInfo its scheduling burst produced, masterloop.cc:108-117)
 * to pf_eval_frontier(), which scores all B nodes over all items in one pass of hand-written
 * sm_100a kernels and returns, per node, exactly those two numbers.
 *
 * Plain C: opaque handle, plain pointers and sizes, int status codes, no exceptions across
 * the boundary, one calling thread per handle (every reference process is single threaded).
 * Caller owns theta/result buffers; the engine owns its device copy of the data set.
 *
 * What each entry point replaces in the reference:
 *   pf_engine_create       the data set every worker builds/holds for itself:
 *                          samplers/normal.cc:38-48 (NormalSampler<2> over NDATA points),
 *                          pysamplers/pymixture.py:57-64, pysamplers/pyblasso.py:44-69,
 *                          samplers/python.cc:96-139 (set_arguments -> sampler(args))
 *   pf_engine_data_size    the size_t a worker REGISTERs (worker.hh:66; python.cc:196-204;
 *                          pymixture.py:66-67; pyblasso.py:79-80)
 *   pf_eval_frontier       Executor::evaluate / evaluate_many over [0, data_size) for every
 *                          frontier node: normalsampler.hh:128-165, boltzmannsampler.hh:71-82,
 *                          python.cc:182-192 -> pymixture.py:86-92, pyblasso.py:135-141
 *   pf_eval_frontier_range the same over items [first, first+count) -- the partial updates of
 *                          worker.hh:141-162 (batch_size items per MetricChange)
 *   pf_blasso_log_prior    python.cc:168-176 -> pyblasso.py:123-133
 *   pf_engine_peer_attach  nothing (the reference has no data parallelism): row
 *   pf_engine_comm_init    shards of one data set on several GPUs, partial (sum, sum_sq) exchanged
 *                          inside the frontier kernel over NVLink peer memory, or all-reduced with NCCL
 *
 * THETA LAYOUT (doubles per node = pf_engine_theta_len()):
 *   PF_MODEL_NORMAL           D means, then the packed lower-triangular covariance,
 *                             D + D(D+1)/2 values (normaltheta.hh:24,49,81-87)
 *   PF_MODEL_BOLTZMANN_SCALAR 1 value (boltzmannsampler.hh:31)
 *   PF_MODEL_BOLTZMANN_DENSE  D unit states x (new workload; log p = -x'Wx/T, one item)
 *   PF_MODEL_MIXTURE          K*d component means, row-major [K][d] (pymixture.py:72-73)
 *   PF_MODEL_BLASSO           sigma, then p coefficients beta (pyblasso.py:138-139)
 *
 * ITEMS.  metric_sum_sq is a sum of squares of ITEM log-probabilities.  An item is one point
 * for the normal target (normalsampler.hh:128-129), one batch of `item_rows` rows for the
 * Python models (pymixture.py:86-92, pyblasso.py:135-141), the whole energy for Boltzmann.
 */
#ifndef PFETCH_H
#define PFETCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_ABI_VERSION 1

#if defined(__GNUC__)
#define PF_API __attribute__((visibility("default")))
#else
#define PF_API
#endif

typedef struct pf_engine pf_engine;

enum pf_status {
    PF_OK = 0,
    PF_ERR_INVALID = -1,      /* bad argument (null pointer, B == 0, unknown model ...) */
    PF_ERR_CUDA = -2,         /* a CUDA call failed; see pf_last_error() */
    PF_ERR_SHAPE = -3,        /* sizes inconsistent with the model */
    PF_ERR_NOMEM = -4,
    PF_ERR_UNSUPPORTED = -5,  /* e.g. dimension outside what the kernels are built for */
    PF_ERR_NCCL = -6,         /* the exchange between GPUs failed (NCCL error, or a peer never arrived) */
    PF_ERR_NO_DEVICE = -7     /* no sm_100 device: there is NO CPU fallback */
};

enum pf_model {
    PF_MODEL_NORMAL = 1,            /* samplers/normal.cc, normalsampler.hh */
    PF_MODEL_BOLTZMANN_SCALAR = 2,  /* boltzmannsampler.hh */
    PF_MODEL_BOLTZMANN_DENSE = 3,   /* BASELINE config 2 (no reference implementation) */
    PF_MODEL_MIXTURE = 4,           /* pysamplers/pymixture.py */
    PF_MODEL_BLASSO = 5             /* pysamplers/pyblasso.py */
};

enum pf_dtype {
    PF_F64 = 0, /* data stored and computed in IEEE double, like the reference (1e-10 rel) */
    PF_F32 = 1  /* data stored/computed in float, FP64 accumulation (1e-5 rel) */
};

typedef struct pf_model_desc {
    int32_t model;       /* enum pf_model */
    int32_t dtype;       /* enum pf_dtype */
    int32_t device;      /* CUDA device ordinal */
    int32_t reserved;
    uint64_t n_rows;     /* rows of `data` held by THIS engine (points / X rows / W rows);
                            for BOLTZMANN_SCALAR: the item count (testmaster -d) */
    uint32_t dim;        /* D (normal), d (mixture), p (lasso), D (dense Boltzmann) */
    uint32_t components; /* K (mixture), else 0 */
    uint64_t item_rows;  /* rows per item: 1 (normal), batch (mixture), 18000 (lasso), 0 = default */
    uint64_t n_items;    /* items held by this engine; 0 = derive it the reference's way:
                            n_rows (normal); (n_rows+1)/item_rows but whole items only (mixture,
                            pymixture.py:64); n_rows/item_rows (lasso, pyblasso.py:69); 1 (dense) */
    double param;        /* temperature T (Boltzmann, boltzmannsampler.hh:36-38; 0 -> 1.0) */
    const double* data;  /* HOST, row-major [n_rows][dim] doubles; copied to the device once */
    const double* response; /* HOST, lasso y[n_rows]; else NULL */
} pf_model_desc;

/* Create an engine; the data set is copied to device memory (converted to float for PF_F32)
 * and stays resident.  For a row-sharded data set each rank passes only its own rows, cut at
 * item boundaries.  Returns PF_OK or a negative pf_status. */
PF_API int pf_engine_create(const pf_model_desc* desc, pf_engine** out);
PF_API void pf_engine_destroy(pf_engine* e);

PF_API uint64_t pf_engine_data_size(const pf_engine* e); /* items held by this engine */
PF_API uint32_t pf_engine_theta_len(const pf_engine* e); /* doubles per node */
PF_API uint32_t pf_engine_max_frontier(const pf_engine* e); /* nodes scored per kernel pass; larger B is chunked */
PF_API int pf_engine_model(const pf_engine* e);           /* enum pf_model */
PF_API uint32_t pf_engine_dim(const pf_engine* e);
PF_API uint32_t pf_engine_components(const pf_engine* e);
/* Multi-GPU group of this engine: *rank / *nranks as set by pf_engine_peer_attach / pf_engine_comm_init (0 / 1 for
 * a single engine), *node_split = PF_OPT_NODE_SPLIT.  Any pointer may be NULL. */
PF_API int pf_engine_group(const pf_engine* e, int* rank, int* nranks, int* node_split);

/* Score B frontier nodes over ALL items.  thetas: HOST [B][theta_len]; sum, sum_sq: HOST [B]
 * (overwritten).  The thetas reach the device in the kernel parameters (small frontiers of the pointwise
 * models) or by a host->device copy on the engine's stream; the frontier kernel(s) -- including the exchange
 * between GPUs when peers are attached -- deliver the 2*B results as self-validating words in mapped pinned
 * memory, which this call polls: no device->host copy, no stream synchronise.  Returns when the results are
 * in `sum` / `sum_sq`. */
PF_API int pf_eval_frontier(pf_engine* e, uint32_t B, const double* thetas, double* sum, double* sum_sq);

/* Optional zero-copy staging: a PINNED host buffer owned by the engine, room for pf_engine_max_frontier() nodes
 * (*capacity_doubles, if non-NULL).  A caller that assembles its frontier there ([B][theta_len], from the start of the
 * buffer) and passes that pointer as `thetas` to pf_eval_frontier / _range / _repeat saves the engine's own copy into
 * pinned memory in front of the host->device DMA -- what a master that gathers the nodes' proposals (protocol.hh:200-206
 * strings) into one buffer anyway should do.  Any other pointer works as before.  Valid until pf_engine_destroy. */
PF_API double* pf_engine_theta_buffer(pf_engine* e, uint64_t* capacity_doubles);

/* The same call `iters` times back to back (every iteration a complete pf_eval_frontier: theta from the host
 * buffer, kernel, results back in `sum` / `sum_sq`), timed with the host's monotonic clock into *seconds.  What a
 * C++ master's loop costs per burst, without a scripting language's per-call overhead in between (bench.py). */
PF_API int pf_eval_frontier_repeat(pf_engine* e, uint32_t B, const double* thetas, double* sum, double* sum_sq,
                            uint32_t iters, double* seconds);

/* Same over items [first_item, first_item + n_items) only.  For PF_MODEL_NORMAL the items are
 * visited in the reference's per-node pseudo-random order when `order` is non-NULL:
 * order[2*b] = stride, order[2*b+1] = start of node b's StrideIndex (strideindex.hh:46-48,
 * normalsampler.hh:115-119); NULL means natural order.  Results are the sums over the range
 * (the caller adds them to its running MetricChange, worker.hh:130-134). */
PF_API int pf_eval_frontier_range(pf_engine* e, uint32_t B, const double* thetas, uint64_t first_item,
                           uint64_t n_items, const uint64_t* order, double* sum, double* sum_sq);

/* Device-resident variant: d_thetas [B][theta_len] and d_out [2][B] (sums, then sums of
 * squares) are DEVICE pointers; work is enqueued on `cuda_stream` (a cudaStream_t; NULL = the
 * engine's own stream) and the call returns without synchronising.  B <= pf_engine_max_frontier.
 * An engine has ONE set of scratch buffers (per-CTA records, ticket, coefficient panel, exchange sequence):
 * it is not reentrant.  Calls on different streams are ordered by the engine itself (an event recorded behind
 * each launch, waited on by the next launch that uses another stream), so they never overlap; for concurrency
 * create one engine per stream. */
PF_API int pf_eval_frontier_device(pf_engine* e, uint32_t B, const double* d_thetas, double* d_out,
                            void* cuda_stream);

/* BLasso.log_prior (pyblasso.py:123-133): host-side, O(p); tau is the Laplace rate (5.0). */
PF_API double pf_blasso_log_prior(const double* theta, uint32_t p, double tau);

/* Row-sharded multi-GPU: one engine per rank over its shard; after pf_engine_comm_init the
 * eval calls all-reduce (sum) the 2*B partials with NCCL so every rank gets the totals.
 * pf_nccl_unique_id fills a 128-byte id on one rank; ship it to the others out of band. */
PF_API int pf_nccl_unique_id(void* id128);
PF_API int pf_engine_comm_init(pf_engine* e, const void* id128, int rank, int nranks);

/* The same reduction WITHOUT a collective kernel: the frontier kernel's finalizing CTA stores this rank's
 * [2][B] partials straight into every peer's exchange buffer over NVLink / NVSwitch (peer-mapped device
 * memory), waits for the other ranks' records and sums them in rank order, so all ranks get bit-identical
 * totals and the host path keeps its copy-free completion flag.  Replaces nothing in the reference (it has
 * no data parallelism); replaces the ncclAllReduce above and is preferred once attached.
 *   pf_engine_peer_export        64-byte CUDA IPC handle of this engine's exchange buffer (one process per GPU:
 *                                all-gather the handles out of band, e.g. torch.distributed)
 *   pf_engine_peer_attach        handles[nranks][64] of all ranks in rank order (own entry ignored); sets
 *                                rank / nranks like pf_engine_comm_init and selects the fused exchange
 *   pf_engine_peer_attach_local  all engines of the group live in THIS process (any devices): no IPC
 *   pf_engine_peer_status        PF_ERR_NCCL after a call in which some rank never arrived (20 s timeout)
 * Every rank must make the same sequence of evaluation calls (as with any collective), with the same B and, for
 * pf_eval_frontier_range, any LOCAL item range it likes -- a rank whose range is empty (n_items == 0) still takes
 * part in the exchange and contributes zeros.  nranks <= 16.  An engine is attached once: regrouping means new
 * engines.  The scalar Boltzmann model has no data to shard and refuses to be grouped (PF_ERR_UNSUPPORTED).
 * pf_chain_run on a row-sharded group runs the same deterministic chain on every rank: every rank must hold the
 * SAME number of items (the chain sees nranks x that many) and pass the same configuration. */
PF_API int pf_engine_peer_export(pf_engine* e, void* handle64);
PF_API int pf_engine_peer_attach(pf_engine* e, const void* handles, int rank, int nranks);
PF_API int pf_engine_peer_attach_local(pf_engine* const* engines, int nranks);
PF_API int pf_engine_peer_status(pf_engine* e);

/* ---- a whole chain through the engine (host side: speculative tree + sampler + frontier hand-off) ----
 *
 * Equivalent of `mpirun -np frontier+1 ./fetching -S <sampler> -l chain_length -s policy -c correlation
 * -d dimension -r reassign_ratio [-t] [-p]` (tree.cc:43-170): the master's speculative tree keeps
 * drawing uniforms and proposals from the same replayable streams, but every scheduling burst
 * (masterloop.cc:108-117) is scored by ONE pf_eval_frontier call instead of one MPI worker per node.
 * The sampler follows the engine's model.  `row` receives every chain level the reference would
 * print (stype.cc:67-84); `job` (optional) receives every JobInfo handed out and the proposal it
 * produced, for tree-indexing parity tests. */
typedef struct pf_chain_config {
    uint32_t frontier;           /* virtual workers == nodes per engine call (mpirun -np minus 1) */
    uint32_t scheduling_policy;  /* -s: 0 naive, 1 constant, 2 predictive (master.hh:29-33) */
    uint64_t chain_length;       /* -l */
    double correlation;          /* -c (0 -> 0.99) */
    double reassign_ratio;       /* -r */
    uint32_t dimension;          /* -d: initial proposal scale log(2.38^2 / d) (master.cc:25) */
    int32_t threshold;           /* -t */
    int32_t verbose;             /* -p: per-level table on stderr (master.cc:324-345) */
    int32_t print_tsv;           /* print `time depth accept metric_sum theta` rows on stdout */
    uint32_t seed;               /* mixture: seed of first_proposal (pymixture.py:71) */
    uint32_t deferred_proposals; /* 0: the tree learns a proposal as soon as it is made (MasterTester-like);
                                    1: only when the master reads SETPROPOSAL, so jobs carry replay paths */
    uint64_t batch_items;        /* -b: 0 = score nodes whole; > 0 = partial updates every batch_items items, which
                                    feed the predictive scheduler and allow abandoning (worker.hh:141-175) */
    uint32_t update_style;       /* -y: 0 fast/slow, 1 incremental, 2 logarithmic (protocol.hh:26-30) */
    uint32_t reserved2;
    double tau;                  /* lasso prior rate (0 -> 5.0, pyblasso.py:32) */
    const double* initial_theta; /* lasso: optional initial (sigma, beta); NULL -> (1, 0, ..., 0) */
} pf_chain_config;

typedef struct pf_chain_stats {
    uint64_t levels;        /* completed chain steps */
    uint64_t rows;          /* rows delivered to `row` */
    uint64_t engine_calls;  /* frontier batches == kernel passes */
    uint64_t nodes_scored;  /* log-likelihood evaluations */
    uint64_t max_frontier;
    uint64_t allocated_nodes, recycled_nodes;
    double seconds, eval_seconds;
    uint64_t items_scored;  /* sum over engine calls of nodes x items: the work paid for */
    uint64_t abandoned;     /* nodes dropped before completion (ABANDON, or found dead) */
    uint64_t accepted;      /* accepted proposals among levels 1.. (the chain's acceptance count) */
    double min_margin;      /* min over decided levels of |delta_log_prior + sum - parent_sum - log u| (master.cc:89):
                               an error in the sums smaller than this flips no accept / reject decision */
    double min_margin_rel;  /* min of that margin / max(|sum|, |parent_sum|): compare with the 1e-10 (FP64) or
                               1e-5 (FP32) relative tolerance of the sums; +inf before the first decision */
} pf_chain_stats;

typedef void (*pf_chain_row_fn)(void* user, uint64_t depth, int accept, double metric_sum, const double* theta,
                                uint32_t theta_len);
typedef void (*pf_chain_job_fn)(void* user, uint64_t id, uint64_t generation, int state, const char* path,
                                uint64_t random_position, uint64_t proposal_random_position,
                                uint64_t next_random_position, double log_prior, const double* theta,
                                uint32_t theta_len);

PF_API int pf_chain_run(pf_engine* e, const pf_chain_config* cfg, pf_chain_row_fn row, pf_chain_job_fn job,
                        void* user, pf_chain_stats* stats);

/* The same host machinery (tree, sampler, frontier hand-off) with a caller-supplied frontier
 * evaluator instead of the GPU engine: eval(user, B, thetas[B][theta_len], sum[B], sum_sq[B]) -> 0.
 * Needs no device; it exists so the host logic can be exercised anywhere (the CPU test-suite
 * plugs the oracle in here) and so another evaluator can sit behind the same tree. */
typedef int (*pf_chain_eval_fn)(void* user, uint32_t B, const double* thetas, double* sum, double* sum_sq);
/* items [first, first + count); order[2b], order[2b+1] = stride, start of node b (normal target), else (1, 0) */
typedef int (*pf_chain_range_fn)(void* user, uint32_t B, const double* thetas, uint64_t first, uint64_t count,
                                 const uint64_t* order, double* sum, double* sum_sq);
PF_API int pf_chain_run_with_evaluator(int model, uint32_t dim, uint32_t components, uint64_t data_size,
                                       const pf_chain_config* cfg, pf_chain_eval_fn eval,
                                       pf_chain_range_fn eval_range, void* eval_user, pf_chain_row_fn row,
                                       pf_chain_job_fn job, void* user, pf_chain_stats* stats);

/* Options. */
enum pf_option {
    /* PF_MODEL_NORMAL, device-pointer path only: promise that every node's covariance is the
     * identity (what the reference's proposals always produce, normalsampler.hh:73-75), which
     * selects the cheaper kernel.  The host-pointer path inspects the thetas itself. */
    PF_OPT_ASSUME_IDENTITY_COV = 1,
    /* After pf_engine_comm_init: 0 (default) = ROW SHARDS -- every rank holds different rows, per-node partial
     * sums are all-reduced; 1 = NODE SPLIT -- every rank holds the WHOLE (small) data set, rank r scores nodes
     * [r*ceil(B/G), ...) and the results are all-gathered (host-pointer calls only). */
    PF_OPT_NODE_SPLIT = 2,
    /* Row shards: 1 = reduce inside the frontier kernel over peer memory (default once pf_engine_peer_attach
     * succeeded), 0 = ncclAllReduce after the kernel (needs pf_engine_comm_init). */
    PF_OPT_PEER_EXCHANGE = 3,
    /* Lasso / dense Boltzmann: 1 (default) = contraction on the tensor cores, 0 = FMA on the CUDA cores.
     * FP64: mma.sync.m8n8k4.f64 (DMMA) -- same IEEE arithmetic, different summation order over the columns.
     * PF_F32 lasso: tcgen05.mma kind::tf32 with every operand split into two tf32 pieces (hi.hi + lo.hi + hi.lo,
     * FP32 accumulation in TMEM): FP32-faithful to ~1e-6 relative, inside the 1e-5 contract of PF_F32; used for
     * every frontier size (it is faster than the CUDA-core variant from B = 1 on); 0 selects the CUDA cores. */
    PF_OPT_MATVEC_TENSOR = 4,
    /* Host-pointer calls of the pointwise models: 1 (default) = frontiers of up to 432 doubles of theta travel in
     * the kernel parameters, 0 = always through a host->device copy in front of the kernel. */
    PF_OPT_INLINE_THETA = 5,
    /* Host-pointer calls over ALL items of a pointwise model (normal target, mixture) whose data set fits the shared
     * memory of the GPU (about 1.1e6 rows of two doubles): 1 (default) = served by the RESIDENT frontier kernel --
     * launched once, rows / tables / barriers kept in shared memory between calls, driven through a doorbell in mapped
     * host memory (thetas in, sums out as self-validating words; no launch, no copy, no synchronise per call).  It
     * leaves the SMs when this process launches anything else on the device, when the engine is destroyed, or after
     * 2 ms without a call (PF_RESIDENT_IDLE_US), and comes back with the next call.  0 = every call launches a kernel.
     * Results are bit-identical either way.  The environment variable PF_RESIDENT=0 changes the default. */
    PF_OPT_RESIDENT = 6
};
PF_API int pf_engine_set_option(pf_engine* e, int option, int64_t value);

/* Introspection. */
PF_API int pf_abi_version(void);
PF_API const char* pf_last_error(void);                 /* thread-local text of the last failure */
PF_API uint64_t pf_engine_launch_count(const pf_engine* e); /* kernels launched by this engine so far */
/* Calls served by the resident frontier kernel so far (*calls) and how often it was launched (*launches); either
 * pointer may be NULL.  Returns 1 if the engine's data set qualifies for it (known after the first eligible call). */
PF_API int pf_engine_resident_stats(const pf_engine* e, uint64_t* calls, uint64_t* launches);
PF_API int pf_device_sm_count(int device);
/* Diagnostics: apply the engine's device exp(-x/2) (which = 0: double, 2: float; 5: the mixture loop's
 * 2048-entry degree-3 version; 9: its degree-2 successor, the clamped loop's) or log(x) (1: double, 3: float, 4: table
 * version, 6: the mixture loop's degree-6 version, 8: the degree-4 version with the exponent accumulated apart), or
 * 2^x (7: the factored mixture loop's) elementwise to a HOST array, for accuracy tests of the math kernels. */
PF_API int pf_debug_math(int which, const double* in, double* out, size_t n, int device);
/* Diagnostics: measured FP64 FMA throughput (1e12 FMA/s) for warps/SM in {4,8,16,32} x
 * independent chains per thread in {1,2,4,8}; out16[4*w + i].  The largest entry is the FP64
 * roofline denominator (MEASURED_PEAKS.json has no FP64 figure). */
PF_API int pf_debug_fp64_probe(int device, double* out16);
/* Diagnostics: FMA rate (1e12/s) through the FP64 tensor-core instruction mma.sync.m8n8k4.f64 (DMMA) for
 * warps/SM in {8,16} x independent accumulator tiles per warp in {1,2,4,8}; out8[4*w + i]. */
PF_API int pf_debug_dmma_probe(int device, double* out8);
/* Diagnostics: DFMA rate (1e12/s) with k = 0..3 integer multiply-adds interleaved per DFMA (issue-slot model). */
PF_API int pf_debug_issue_probe(int device, double* out4);
/* Diagnostics: DFMA rate (1e12/s) with 1, 2, 3 distinct register operands per DFMA, then a
 * 2-register DFMA with 2 and 4 one-register ALU instructions between (register-file bandwidth); 5 values. */
PF_API int pf_debug_rf_probe(int device, double* out5);
/* Diagnostics: do DFMA (FP64 pipe) and DMMA m8n8k4 (tensor pipe) overlap?  SM clocks per block per sub-partition for
 * blocks of (DFMAs, DMMAs) = (16,0) (0,2) (0,4) (16,1) (16,2) (16,4) (32,4) (8,4), every DMMA tile consumed by 2 DFMAs. */
PF_API int pf_debug_dfma_dmma_probe(int device, double* out8);

/* Diagnostics: the floor of a host-buffer call -- microseconds per {kernel launch, one word delivered to mapped host
 * memory, host polls it} for {24-byte parameters, 1 CTA}, {3.6 KB parameters, 1 CTA}, {24 B, 2 x SMs CTAs + ticket},
 * {3.6 KB, 2 x SMs CTAs + ticket}: what pf_eval_frontier costs before any scoring. */
PF_API int pf_debug_launch_probe(int device, double* out4);

#ifdef __cplusplus
}
#endif
#endif /* PFETCH_H */
