/* libzoo_sm100 -- C ABI of the B200-native (sm_100a) value-based learner update.
 *
 * Drop-in boundary for ONE hot path of ReinforcementLearningZoo.jl v0.3.7: `RLBase.update!` of
 * BasicDQNLearner / DQNLearner / PrioritizedDQNLearner / RainbowLearner / IQNLearner on a replay
 * minibatch, plus the prioritized-replay SumTree sample / priority update.
 *
 * The reference has no FFI of its own (pure Julia, multiple dispatch); each entry point below
 * cites the reference method (file:line under /root/reference) whose body it replaces, and
 * INTEGRATION.md shows the `ccall` glue a maintainer adds on the Julia side.
 *
 * Conventions
 *  - every function returns zoo_status (0 = OK); zoo_last_error() gives a thread-local message.
 *  - handles are opaque, library-owned, freed by *_destroy; NOT thread-safe: serialise calls per
 *    handle (the reference drives this path from one Julia task).
 *  - array arguments are caller-owned and borrowed for the duration of the call only.  Host
 *    pointers unless a `*_dev` variant / `on_device` flag says otherwise.
 *  - layouts are the bytes of the Julia arrays:
 *        frames / conv inputs   (W,H,C,B) column-major  == row-major [B][C][H][W]
 *        CrossCor weight        (kw,kh,Cin,Cout)        == row-major [Cout][Cin][kh][kw]
 *        Dense weight           (out,in) column-major   == row-major [in][out]
 *        Q / logits / quantiles (A,B) / (atoms,A,B) / (A,N,B) column-major
 *                               == row-major [B][A] / [B][A][atoms] / [B][N][A]
 *    actions are 0-based int32 here (the Julia glue subtracts 1); SumTree indices are 1-based
 *    int64 logical indices exactly as Julia returns them.
 *  - there is no CPU fallback: on a device that is not compute capability 10.x every compute
 *    entry point returns ZOO_UNSUPPORTED_ARCH.
 */
#ifndef ZOO_SM100_H
#define ZOO_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ZOO_API __attribute__((visibility("default")))
#else
#define ZOO_API
#endif

typedef int32_t zoo_status;
enum { ZOO_OK = 0, ZOO_BAD_ARG = 1, ZOO_CUDA_ERROR = 2, ZOO_NCCL_ERROR = 3, ZOO_UNSUPPORTED_ARCH = 4 };

typedef struct zoo_net zoo_net;         /* NeuralNetworkApproximator.model  (RLCore; used at dqn.jl:44-80) */
typedef struct zoo_opt zoo_opt;         /* NeuralNetworkApproximator.optimizer (Dopamine_DQN_Atari.jl:42)   */
typedef struct zoo_learner zoo_learner; /* the *Learner structs: dqn.jl:3-21, rainbow.jl:28-53, iqn.jl:61-79 */
typedef struct zoo_sumtree zoo_sumtree; /* RLCore SumTree behind t[:priority] (common.jl:18,22,36)          */
typedef struct zoo_dp zoo_dp;           /* data-parallel communicator (new; no counterpart in the reference) */

ZOO_API const char *zoo_last_error(void);
ZOO_API int32_t zoo_version(void);
/* compute capability major*10+minor of the current device, or <0 if no device */
ZOO_API int32_t zoo_device_arch(void);
ZOO_API zoo_status zoo_set_device(int32_t device);

/* ------------------------------------------------------------------ networks ------------ */
enum { ZOO_LAYER_SCALE255 = 0, /* x ./ 255            Dopamine_DQN_Atari.jl:28    */
       ZOO_LAYER_CONV = 1,     /* CrossCor            Dopamine_DQN_Atari.jl:29-31 */
       ZOO_LAYER_FLATTEN = 2,  /* reshape(x, :, B)    Dopamine_DQN_Atari.jl:32    */
       ZOO_LAYER_DENSE = 3 };  /* Dense               Dopamine_DQN_Atari.jl:33-34 */

typedef struct {
    int32_t kind;   /* ZOO_LAYER_* */
    int32_t n_in;   /* conv: Cin;  dense: in  */
    int32_t n_out;  /* conv: Cout; dense: out */
    int32_t ksize, stride, pad; /* conv only (square kernels, symmetric zero pad) */
    int32_t relu;   /* activation: 1 = relu, 0 = identity */
} zoo_layer;

enum { ZOO_NET_CHAIN = 0,   /* Flux Chain                                         */
       ZOO_NET_IQN = 1,     /* ImplicitQuantileNet(psi, phi, header)  iqn.jl:17-33 */
       ZOO_NET_DUELING = 2  /* DuelingNetwork(base, val, adv)     common.jl:44-56 */ };

enum { ZOO_PREC_FP32 = 0,      /* fp32 storage; contractions on tcgen05 as 3xTF32 (error-compensated hi/lo split,
                                   fp32-grade) with CUDA-core fallback for tiny/misaligned shapes; parity <= 1e-4 rel */
       ZOO_PREC_BF16 = 1,      /* bf16 operands on tcgen05 (kind::f16), fp32 TMEM accumulate; fastest, see DESIGN.md */
       ZOO_PREC_TF32 = 2,      /* fp32 storage, single-pass tcgen05 kind::tf32                                       */
       ZOO_PREC_FP32_SIMT = 3  /* fp32 storage, CUDA-core FFMA contractions only (second, independent fp32 path)      */ };

#define ZOO_MAX_LAYERS 12
typedef struct {
    int32_t kind;                 /* ZOO_NET_* */
    int32_t precision;            /* ZOO_PREC_* */
    int32_t in_c, in_h, in_w;     /* observation shape; vector observations: in_c = n, in_h = in_w = 1 */
    int32_t max_batch;            /* largest B (per call, before the x N quantile expansion) */
    int32_t max_quantiles;        /* IQN: max(N, N', K); else 0 */
    /* three sub-chains, in Flux `params` order:
       CHAIN: a only.  IQN: a = psi, b = phi, c = header.  DUELING: a = base, b = val, c = adv. */
    int32_t n_a, n_b, n_c;
    zoo_layer a[ZOO_MAX_LAYERS], b[ZOO_MAX_LAYERS], c[ZOO_MAX_LAYERS];
} zoo_net_desc;

ZOO_API zoo_status zoo_net_create(const zoo_net_desc *desc, zoo_net **out);
ZOO_API zoo_status zoo_net_destroy(zoo_net *net);
/* params(app): tensors in Flux order W1,b1,W2,b2,... ; sizes in elements */
ZOO_API zoo_status zoo_net_num_tensors(const zoo_net *net, int32_t *n);
ZOO_API zoo_status zoo_net_tensor_size(const zoo_net *net, int32_t tensor_id, int64_t *n_elem);
ZOO_API zoo_status zoo_net_num_params(const zoo_net *net, int64_t *n_elem);
/* Flux.loadparams! / cpu(params): host arrays in the Julia layouts above */
ZOO_API zoo_status zoo_net_set_params(zoo_net *net, int32_t tensor_id, const float *host, int64_t n_elem);
ZOO_API zoo_status zoo_net_get_params(const zoo_net *net, int32_t tensor_id, float *host, int64_t n_elem);
/* copyto!(dst, src)  -- common.jl:12-14 (target sync), dqn.jl:60 / rainbow.jl:88 / iqn.jl:113 (ctor force-sync) */
ZOO_API zoo_status zoo_net_copy(zoo_net *dst, const zoo_net *src);

enum { ZOO_OBS_U8 = 0, ZOO_OBS_F32 = 1 };
/* app(x): batch forward for acting -- dqn.jl:90-100, rainbow.jl:118-124 (caller applies support/softmax),
 * iqn.jl:148-155 (tau: [B][n_tau] host floats, required for IQN nets, else NULL/0).
 * out_host: [B][n_out] (chain/dueling) or [B][n_tau][n_out] (IQN). */
ZOO_API zoo_status zoo_net_forward(zoo_net *net, const void *obs_host, int32_t obs_dtype, int32_t B,
                                   const float *tau_host, int32_t n_tau, float *out_host);
ZOO_API zoo_status zoo_net_out_dim(const zoo_net *net, int32_t *n_out);

/* ------------------------------------------------------------------ optimisers ---------- */
enum { ZOO_OPT_RMSPROP = 0, /* Flux RMSProp(eta, rho)   Dopamine_DQN_Atari.jl:42            */
       ZOO_OPT_ADAM = 1 };  /* Flux ADAM(eta, (b1,b2))  Dopamine_Rainbow_Atari.jl:41, Dopamine_IQN_Atari.jl:51 */
/* update!(app, gs) -- call sites dqn.jl:144, rainbow.jl:193, iqn.jl:234, basic_dqn.jl:89.
 * RMSProp: a = rho, b ignored.  ADAM: a = beta1, b = beta2.  eps is Flux's constant 1e-8.
 * Arithmetic: the kernels evaluate the update in fp32 with scalars (1 - rho, beta^t corrections) derived in fp64 on the host; Flux
 * evaluates the broadcast with Float64 scalars and rounds once per element.  The difference is far below the 1e-4 parity gate
 * (tests/test_witnesses.py holds both against torch.optim over 10 steps). */
ZOO_API zoo_status zoo_opt_create(zoo_net *net, int32_t kind, double eta, double a, double b, double eps, zoo_opt **out);
ZOO_API zoo_status zoo_opt_destroy(zoo_opt *opt);
/* checkpointing (BSON policy save, Dopamine_DQN_Atari.jl:123-126): slot 0 = acc | m, slot 1 = v.
 * Per-tensor host arrays in the same layouts as zoo_net_get_params. */
ZOO_API zoo_status zoo_opt_get_state(const zoo_opt *opt, int32_t slot, int32_t tensor_id, float *host, int64_t n_elem);
ZOO_API zoo_status zoo_opt_set_state(zoo_opt *opt, int32_t slot, int32_t tensor_id, const float *host, int64_t n_elem);
ZOO_API zoo_status zoo_opt_get_step(const zoo_opt *opt, int64_t *t);
ZOO_API zoo_status zoo_opt_set_step(zoo_opt *opt, int64_t t);
/* counter of the on-device Philox tau stream (IQN updates that draw their own quantiles tick it once); part of a checkpoint */
ZOO_API zoo_status zoo_opt_get_rng(const zoo_opt *opt, uint64_t *counter);
ZOO_API zoo_status zoo_opt_set_rng(zoo_opt *opt, uint64_t counter);

/* ------------------------------------------------------------------ learners ------------ */
enum { ZOO_LEARNER_BASIC_DQN = 0,       /* basic_dqn.jl:69-90        */
       ZOO_LEARNER_DQN = 1,             /* dqn.jl:102-145            */
       ZOO_LEARNER_PRIORITIZED_DQN = 2, /* prioritized_dqn.jl:111-151 */
       ZOO_LEARNER_RAINBOW = 3,         /* rainbow.jl:126-221        */
       ZOO_LEARNER_IQN = 4,             /* iqn.jl:159-237            */
       ZOO_LEARNER_QR_DQN = 5,          /* qr_dqn.jl:106-138: cfg.N = n_quantile, cfg.kappa; net outputs n_quantile*n_actions   */
       ZOO_LEARNER_REM_DQN = 6 };       /* rem_dqn.jl:100-145: cfg.N = ensemble_num; batch.tau = the E convex weights (required);
                                           net outputs n_actions*ensemble_num */

/* The reference's learners take a pluggable `loss_func` (dqn.jl:137, prioritized_dqn.jl:140, rainbow.jl:180).  This library
 * implements the ones the reference's experiment files pass: Flux huber_loss (delta = 1; `agg = identity` for the PER learners,
 * Dopamine_DQN_Atari.jl:51) for the DQN family, logitcrossentropy(...; agg = identity) for Rainbow (Dopamine_Rainbow_Atari.jl:54),
 * the quantile Huber loss with `kappa` for IQN / QR-DQN.  Any other loss_func is out of scope and has no entry point. */
typedef struct {
    int32_t kind;            /* ZOO_LEARNER_* */
    int32_t batch_size;      /* sampler.batch_size */
    int32_t n_actions;
    float gamma;             /* sampler.gamma (Float32) */
    int32_t update_horizon;  /* sampler.n; target uses gamma^n (dqn.jl:133) -- IQN uses gamma^1 (iqn.jl:189) */
    int32_t double_dqn;      /* DQNLearner.is_enable_double_DQN (dqn.jl:58: default true) */
    float beta_priority;     /* PER exponent (prioritized_dqn.jl:58) */
    /* Rainbow */
    int32_t n_atoms;
    float v_min, v_max;      /* clamp bounds (rainbow.jl:164-165) */
    const float *support;    /* n_atoms host floats (rainbow.jl:72); copied at create */
    float delta_z;           /* Float32(support[2]-support[1]) (rainbow.jl:74) */
    /* IQN */
    float kappa;             /* iqn.jl:66 */
    int32_t N, N_prime, N_em;
    uint64_t seed;           /* device RNG seed for tau when the batch carries none (iqn.jl:109,173,191) */
} zoo_learner_cfg;

typedef struct {
    int32_t B;                 /* must equal cfg.batch_size */
    int32_t obs_dtype;         /* ZOO_OBS_U8 | ZOO_OBS_F32 */
    const void *state;         /* batch.state       [B][C][H][W]                       */
    const void *next_state;    /* batch.next_state                                     */
    const int32_t *action;     /* batch.action, 0-based                                */
    const float *reward;       /* batch.reward (already the n-step discounted sum)     */
    const uint8_t *terminal;   /* batch.terminal (Bool)                                */
    const float *priority;     /* batch.priority or NULL (PER iff present, rainbow.jl:168) */
    const uint8_t *next_mask;  /* batch.next_legal_actions_mask [B][A] or NULL (dqn.jl:121-124) */
    const float *tau;          /* IQN: [B][N]   or NULL => on-device Philox              */
    const float *tau_prime;    /* IQN: [B][N'] or NULL                                   */
    int32_t on_device;         /* 1: every pointer above is a device pointer             */
} zoo_batch;

/* target may be NULL for BASIC_DQN.  The learner borrows net/opt handles (must outlive it). */
ZOO_API zoo_status zoo_learner_create(const zoo_learner_cfg *cfg, zoo_net *online, zoo_net *target, zoo_opt *opt,
                                      zoo_dp *dp_or_null, zoo_learner **out);
ZOO_API zoo_status zoo_learner_destroy(zoo_learner *lrn);
/* update!(learner, batch): loss_out -> learner.loss (NULL => fully asynchronous enqueue, no sync);
 * new_priority_out: [B] floats (batch_losses .+ 1f-10).^beta, or NULL. */
ZOO_API zoo_status zoo_learner_update(zoo_learner *lrn, const zoo_batch *batch, float *loss_out, float *new_priority_out);
/* update! without waiting for the result (the reference's update! is just as asynchronous under CUDA.jl: dqn.jl:138-140 only stores
 * the loss): submit enqueues the update and a device->host copy of its loss (and, if want_priorities, of its new priorities) into
 * library-owned pinned memory; collect waits for the OLDEST submitted update and hands its results out.  At most 4 updates may be
 * in flight.  Host batches (on_device = 0) are copied on a second stream into one of two staging slots, so the H2D copy of update
 * i + 1 overlaps the kernels of update i; the arrays a batch points to must stay valid until the matching collect returns. */
ZOO_API zoo_status zoo_learner_submit(zoo_learner *lrn, const zoo_batch *batch, int32_t want_priorities);
ZOO_API zoo_status zoo_learner_collect(zoo_learner *lrn, float *loss_out, float *new_priority_out);
/* The two halves of update!, for parity tests and external all-reduce:
 * compute_grads = everything up to and including Zygote.gradient; apply = update!(Q, gs). */
ZOO_API zoo_status zoo_learner_compute_grads(zoo_learner *lrn, const zoo_batch *batch, float *loss_out, float *new_priority_out);
ZOO_API zoo_status zoo_learner_apply(zoo_learner *lrn);
/* gs[p] for one parameter tensor, Julia layout */
ZOO_API zoo_status zoo_learner_get_grad(const zoo_learner *lrn, int32_t tensor_id, float *host, int64_t n_elem);
/* flat fp32 gradient buffer on the device (internal layout), e.g. for an external NCCL all-reduce */
ZOO_API zoo_status zoo_learner_grad_buffer(zoo_learner *lrn, void **dev_ptr, int64_t *n_elem);
/* diagnostics for parity tests: per-sample loss [B] of the last update */
ZOO_API zoo_status zoo_learner_get_per_sample_loss(const zoo_learner *lrn, float *host, int32_t n);
/* CUDA stream (cudaStream_t) all work of this learner is enqueued on; zoo_learner_sync waits for it */
ZOO_API zoo_status zoo_learner_stream(zoo_learner *lrn, void **stream);
ZOO_API zoo_status zoo_learner_sync(zoo_learner *lrn);
/* number of kernel launches enqueued by the last update (a claim the bench reports as gpu_launches) */
ZOO_API zoo_status zoo_learner_last_launches(const zoo_learner *lrn, int32_t *n);
/* capture the device part of update() into a CUDA graph and replay it on later calls (0 = off) */
ZOO_API zoo_status zoo_learner_use_graph(zoo_learner *lrn, int32_t enable);

/* ------------------------------------------------------------------ SumTree ------------- */
enum { ZOO_DRAW_F32 = 0, ZOO_DRAW_F64 = 1 };
enum { ZOO_SAMPLE_INDEPENDENT = 0, /* v = u * total                      */
       ZOO_SAMPLE_STRATIFIED = 1 };/* v = (i-1+u_i)/n * total, i = 1..n */

ZOO_API zoo_status zoo_sumtree_create(int64_t capacity, zoo_sumtree **out);
ZOO_API zoo_status zoo_sumtree_destroy(zoo_sumtree *t);
/* push!(t[:priority], p)  -- common.jl:36 */
ZOO_API zoo_status zoo_sumtree_push(zoo_sumtree *t, float p);
ZOO_API zoo_status zoo_sumtree_push_many(zoo_sumtree *t, const float *p, int64_t n);
/* sample(rng, t, sampler) -- common.jl:18: the caller supplies the n uniform draws u in [0,1)
 * (dtype f32|f64; the descent runs in that dtype) so results are bit-exact and reproducible.
 * idx_out: 1-based logical indices (int64); prio_out: leaf priorities. */
ZOO_API zoo_status zoo_sumtree_sample(zoo_sumtree *t, const void *u, int32_t u_dtype, int64_t n, int32_t mode,
                                      int64_t *idx_out, float *prio_out);
/* t[:priority][inds] .= priorities -- common.jl:22: sequential, in batch order, fp32, duplicates allowed */
ZOO_API zoo_status zoo_sumtree_update(zoo_sumtree *t, const int64_t *idx, const float *p, int64_t n);
ZOO_API zoo_status zoo_sumtree_total(zoo_sumtree *t, float *total);
ZOO_API zoo_status zoo_sumtree_get(zoo_sumtree *t, const int64_t *idx, int64_t n, float *prio_out);
ZOO_API zoo_status zoo_sumtree_length(const zoo_sumtree *t, int64_t *length, int64_t *first);
/* whole tree (nparents + capacity floats, Julia order tree[1..]) for checkpoints and parity tests */
ZOO_API zoo_status zoo_sumtree_tree_len(const zoo_sumtree *t, int64_t *n);
ZOO_API zoo_status zoo_sumtree_export(zoo_sumtree *t, float *host_tree, int64_t n);
ZOO_API zoo_status zoo_sumtree_import(zoo_sumtree *t, const float *host_tree, int64_t n, int64_t length, int64_t first);
/* device-resident variants (pointers are device memory on the current device; results stay on the
 * device, enqueued on the tree's stream -- used by the bench's HBM-resident leg) */
ZOO_API zoo_status zoo_sumtree_sample_dev(zoo_sumtree *t, const void *u_dev, int32_t u_dtype, int64_t n, int32_t mode,
                                          int64_t *idx_dev, float *prio_dev);
ZOO_API zoo_status zoo_sumtree_update_dev(zoo_sumtree *t, const int64_t *idx_dev, const float *p_dev, int64_t n);
ZOO_API zoo_status zoo_sumtree_stream(zoo_sumtree *t, void **stream);
ZOO_API zoo_status zoo_sumtree_sync(zoo_sumtree *t);

/* ------------------------------------------------------------------ device-resident replay ---- */
/* The batch assembly behind `sample(rng, t, sampler)` (common.jl:18; trajectory + NStepBatchSampler built at
 * Dopamine_DQN_Atari.jl:55-66, implemented in RLCore 0.7.2) on the device: uint8 single-frame ring, 4-frame stacking, n-step
 * reward fold with terminal truncation, uniform or SumTree-prioritized indices, priority write-back -- no host round trip
 * between sampling and the update.  Semantics: csrc/replay.cu header. */
typedef struct zoo_replay zoo_replay;
enum { ZOO_REPLAY_UNIFORM = 2 /* besides ZOO_SAMPLE_INDEPENDENT / ZOO_SAMPLE_STRATIFIED for a prioritized shard */ };
ZOO_API zoo_status zoo_replay_create(int64_t capacity, int32_t frame_h, int32_t frame_w, int32_t stack_size, int32_t n_step, float gamma,
                                     int32_t max_batch, int32_t prioritized, zoo_replay **out);
ZOO_API zoo_status zoo_replay_destroy(zoo_replay *r);
/* push!(trajectory, ...) of one transition: its (single) state frame [h][w] uint8, action (0-based), reward, terminal,
 * and -- prioritized shard -- push!(t[:priority], priority) (common.jl:28-37) */
ZOO_API zoo_status zoo_replay_push(zoo_replay *r, const uint8_t *frame_host, int32_t action, float reward, uint8_t terminal, float priority);
/* bulk push of n transitions: frames [n][h][w]; priority may be NULL (default_priority 100) */
ZOO_API zoo_status zoo_replay_push_many(zoo_replay *r, const uint8_t *frames_host, const int32_t *action, const float *reward,
                                        const uint8_t *terminal, const float *priority, int64_t n);
ZOO_API zoo_status zoo_replay_length(const zoo_replay *r, int64_t *length, int64_t *first);
/* u_host: B uniform draws (f32 | f64).  Fills `batch` with device pointers (on_device = 1; valid until the next sample),
 * enqueued on `stream` (pass zoo_learner_stream).  inds_out: optional host array of the 1-based logical indices (synchronises). */
ZOO_API zoo_status zoo_replay_sample(zoo_replay *r, void *stream, const void *u_host, int32_t u_dtype, int32_t B, int32_t mode, zoo_batch *batch,
                                     int64_t *inds_out);
/* t[:priority][inds] .= priorities (common.jl:22) for the last sample; prio_dev = device pointer, e.g. zoo_learner_priorities_dev */
ZOO_API zoo_status zoo_replay_update_priorities(zoo_replay *r, void *stream, const float *prio_dev);
/* RLCore's prioritized sampler REDRAWS an index that falls outside the valid range stack .. length - n (call site
 * src/algorithms/dqns/common.jl:18).  The redraws need fresh uniforms: spare_host holds [B][n_per_sample] of them (dtype as the
 * draws passed to zoo_replay_sample; sample b uses row b, each an independent single draw).  They are used by every following
 * zoo_replay_sample until replaced; an index still invalid after n_per_sample redraws (or with none supplied) is clamped to the
 * range edge, counted (zoo_replay_invalid_count) and given the priority of the transition actually fetched. */
ZOO_API zoo_status zoo_replay_set_spare_draws(zoo_replay *r, void *stream, const void *spare_host, int32_t u_dtype, int32_t B, int32_t n_per_sample);
/* PER draws that fell outside the valid range (the reference redraws them; here they are clamped and counted) */
ZOO_API zoo_status zoo_replay_invalid_count(zoo_replay *r, uint32_t *n);
ZOO_API zoo_status zoo_replay_sumtree(zoo_replay *r, zoo_sumtree **tree);
/* device buffer [B] of the learner's new priorities (batch_losses .+ 1f-10).^beta of the last update */
ZOO_API zoo_status zoo_learner_priorities_dev(zoo_learner *lrn, float **dev_ptr);

/* ------------------------------------------------------------------ data parallel ------- */
/* One process per GPU.  Rank 0 calls zoo_dp_unique_id, the host shares the 128 bytes with the other
 * ranks (torch.distributed / MPI / a file), every rank calls zoo_dp_init.  A learner created with a
 * zoo_dp all-reduces (sum) its flat gradient buffer with NCCL between backward and the optimiser step;
 * the per-rank batch is the rank's shard, the loss is scaled by 1/(B*world). */
ZOO_API zoo_status zoo_dp_unique_id(uint8_t id_out[128]);
ZOO_API zoo_status zoo_dp_init(const uint8_t id[128], int32_t rank, int32_t world, zoo_dp **out);
ZOO_API zoo_status zoo_dp_destroy(zoo_dp *dp);

/* ------------------------------------------------------------------ profiling ----------- */
/* In-situ launch profiler: between begin and end an event is recorded on `stream` after every kernel this
 * thread enqueues through the library (eager mode only, not graph replay).  delay_us > 0 first parks a sleeping
 * kernel on the stream so the launches queue up behind it and the gaps are device time, not host launch rate.
 * zoo_prof_get returns launch i's
 * label ("<phase>:<layer>|<launcher>:<line>") and its device time in ms (time since the previous launch ended). */
ZOO_API zoo_status zoo_prof_begin(void *stream, int32_t delay_us);
ZOO_API zoo_status zoo_prof_end(int32_t *n_launches);
ZOO_API zoo_status zoo_prof_get(int32_t i, char *name, int32_t name_cap, float *ms);
/* Timeline mode: zoo_prof_timeline(1, NULL); run one eager update (its lanes stay concurrent); zoo_prof_timeline(0, &n).  Record i =
 * one kernel launch: its phase label, the stream it ran on, start / end in microseconds since the first launch. */
ZOO_API zoo_status zoo_prof_timeline(int32_t on, int32_t *n_records);
ZOO_API zoo_status zoo_prof_timeline_get(int32_t i, char *name, int32_t name_cap, uint64_t *stream, float *start_us, float *end_us);

/* ------------------------------------------------------------------ test hooks ---------- */
/* Stand-alone GEMM of the two contraction engines, for parity tests of the kernels themselves:
 * D[M][N] = sum_k A(m,k) B(n,k).  a_kmajor: A stored [M][K] (1) or [K][M] (0); same for B ([N][K] / [K][N]).
 * engine: 0 = CUDA-core fp32, 1 = tcgen05 bf16 (inputs rounded to bf16), 2 = tcgen05 tf32, 3 = tcgen05 3xTF32,
 * 4 = the persistent tcgen05 bf16 engine (large shapes only).  Host pointers. */
ZOO_API zoo_status zoo_test_gemm(int32_t engine, int32_t M, int32_t N, int32_t K, const float *A, int32_t a_kmajor,
                                 const float *B, int32_t b_kmajor, float *D, int32_t split_k);

/* Tuning hook: force the tcgen05 GEMM's shared-memory ring depth (0 = chosen per launch). */
ZOO_API zoo_status zoo_test_set_tc_stages(int32_t stages);
/* device-time benchmark of one GEMM shape: mean microseconds over `iters` back-to-back GEMMs (operands resident, zeros).
 * mode: 0 plain store, 1 atomic accumulate (wgrad), 2 bias + relu (forward); + 8 = cold operands (each launch reads its own
 * copy of A and B, the copies together larger than L2). */
ZOO_API zoo_status zoo_test_gemm_bench(int32_t engine, int32_t M, int32_t N, int32_t K, int32_t a_kmajor, int32_t b_kmajor,
                                       int32_t split_k, int32_t mode, int32_t iters, float *us_out);

/* %globaltimer stamps (ns) of CTA (0,0,0) of one tcgen05 GEMM launch at its phase boundaries (see core.cu) */
ZOO_API zoo_status zoo_test_gemm_trace(int32_t engine, int32_t M, int32_t N, int32_t K, int32_t a_kmajor, int32_t b_kmajor,
                                       int32_t split_k, uint64_t *stamps9);

/* Per-launch phase trace of the tcgen05 engine: begin(cap) arms it; every engine launch from then on (also those captured into a
 * CUDA graph, on every replay) stamps CTA (0,0,0)'s phases into its own 16-slot record; stop() returns the number of records and
 * stops handing out new ones; get(i) reads record i (name = kind[MxNxK]<grid>, stamps = ns of %globaltimer, see core.cu). */
ZOO_API zoo_status zoo_test_trace_begin(int32_t cap);
ZOO_API zoo_status zoo_test_trace_stop(int32_t *n);
ZOO_API zoo_status zoo_test_trace_get(int32_t i, char *name, int32_t name_cap, uint64_t *stamps16);

/* Parity hook (tests/test_gpu_full_size.py): raw bytes of the output buffer of parameter layer `layer` of chain `chain`
 * (0 = Chain / psi / base, 1 = phi / val, 2 = header / adv) as the last forward left it: post-activation values, internal
 * layout [rows][out_h][out_w][out_c] in the net's activation type.  *is_bf16 tells the element type (the network's output
 * layer is always fp32).  The caller must have synchronised the learner (an update with loss_out does).  The test derives the
 * relu masks the device used from it, so the CPU twin can be re-run with exactly those masks. */
ZOO_API zoo_status zoo_test_net_activation(zoo_net *net, int32_t chain, int32_t layer, int64_t rows, void *host, int64_t host_bytes,
                                           int32_t *is_bf16);

#ifdef __cplusplus
}
#endif
#endif /* ZOO_SM100_H */
