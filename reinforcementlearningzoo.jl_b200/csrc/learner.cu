// Optimisers (Flux RMSProp / ADAM) and the learner update orchestration.
// Reference: update!(learner, batch) basic_dqn.jl:69-90, dqn.jl:102-145, prioritized_dqn.jl:111-151,
//            rainbow.jl:126-196, iqn.jl:159-237; update!(Q, gs) -> Flux.Optimise.update!.
#include <cmath>
#include <cstring>

#include "heads.cuh"
#include "net.cuh"

extern "C" zoo_status zoo_internal_tensor_io(zoo_net *net, float *flat, int32_t id, float *host, int64_t n, int to_device);

namespace zoo {
zoo_status dp_allreduce_sum(zoo_dp *dp, float *buf, long long n, cudaStream_t st);
zoo_status dp_allreduce_max(zoo_dp *dp, float *buf, long long n, cudaStream_t st);
int dp_world(const zoo_dp *dp);
zoo_status dp_alloc_registered(zoo_dp *dp, size_t bytes, void **ptr, void **handle);
void dp_deregister(zoo_dp *dp, void *handle);

bool dqn_tail_supported(int N, long long K, int x_bf16, int rows_fwd);
bool dqn_tail_multi_supported(int N, long long K, int x_bf16, int B);
zoo_status dqn_tail_multi(int kind, int double_dqn, const HeadArgs &h, const void *X, int x_bf16, const float *W, const float *bias, const float *q_t,
                          float gamma_n, float beta, int two_rows, int bwd_next, long long K, int relu_prev, float *q_out, float *dq_out, void *dX,
                          float *dW, float *db, float *db_prev, float *loss, float *prio_out, float *part, unsigned *ticket, cudaStream_t st);
zoo_status dqn_tail(int kind, int double_dqn, const HeadArgs &h, const void *X, int x_bf16, const float *W, const float *bias, const float *q_t, float gamma_n,
                    float beta, int rows_fwd, int rows_bwd, long long K, int relu_prev, float *q_out, float *dq_out, void *dX, float *dW, float *db,
                    float *db_prev, float *loss, float *prio_out, cudaStream_t st);

struct OptScalars {
    double b1p, b2p;        // beta^t carried like Flux's per-array state (starts at beta)
    float c1, c2;           // 1/(1-b1p), 1/(1-b2p) for the step being applied
    unsigned long long t;   // optimiser steps applied so far
    unsigned long long rng; // tau stream counter (one tick per learner update)
};

// one thread: latch this step's bias corrections, then advance beta^t (ADAM) and the counters
__global__ void opt_tick_kernel(OptScalars *s, double b1, double b2) {
    s->c1 = (float)(1.0 / (1.0 - s->b1p));
    s->c2 = (float)(1.0 / (1.0 - s->b2p));
    s->b1p *= b1;
    s->b2p *= b2;
    s->t += 1;
}
__global__ void rng_tick_kernel(OptScalars *s) { s->rng += 1; }

// RMSProp: acc = rho*acc + (1-rho)*g^2;  p -= g * eta / (sqrt(acc) + eps).  fp32 arithmetic with fp64-derived scalars.
// Both optimiser kernels leave the gradient buffer zeroed (the next update's split-K / bias-grad atomics accumulate into it),
// which replaces a separate 16 MB memset per update.
__global__ void rmsprop_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ acc, bf16 *__restrict__ shadow,
                               long long n4, float rho, float omr, float eta, float eps) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 pv = reinterpret_cast<float4 *>(p)[i];
    const float4 gv = reinterpret_cast<const float4 *>(g)[i];
    float4 av = reinterpret_cast<float4 *>(acc)[i];
    float *pp = &pv.x; const float *gp = &gv.x; float *ap = &av.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float gg = gp[k];
        const float a = rho * ap[k] + omr * (gg * gg);
        ap[k] = a;
        pp[k] -= gg * (eta / (sqrtf(a) + eps));
    }
    reinterpret_cast<float4 *>(p)[i] = pv;
    reinterpret_cast<float4 *>(acc)[i] = av;
    reinterpret_cast<float4 *>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (shadow) {
        reinterpret_cast<__nv_bfloat162 *>(shadow)[2 * i] = __floats2bfloat162_rn(pv.x, pv.y);
        reinterpret_cast<__nv_bfloat162 *>(shadow)[2 * i + 1] = __floats2bfloat162_rn(pv.z, pv.w);
    }
}

// ADAM: m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g^2; p -= m/(1-b1^t) / (sqrt(v/(1-b2^t)) + eps) * eta
__global__ void adam_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                            bf16 *__restrict__ shadow, long long n4, float b1, float omb1, float b2, float omb2, float eta, float eps,
                            const OptScalars *__restrict__ sc) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float c1 = sc->c1, c2 = sc->c2;
    float4 pv = reinterpret_cast<float4 *>(p)[i];
    const float4 gv = reinterpret_cast<const float4 *>(g)[i];
    float4 mv = reinterpret_cast<float4 *>(m)[i];
    float4 vv = reinterpret_cast<float4 *>(v)[i];
    float *pp = &pv.x; const float *gp = &gv.x; float *mp = &mv.x; float *vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float gg = gp[k];
        const float mm = b1 * mp[k] + omb1 * gg;
        const float vn = b2 * vp[k] + omb2 * (gg * gg);
        mp[k] = mm; vp[k] = vn;
        pp[k] -= (mm * c1) / (sqrtf(vn * c2) + eps) * eta;
    }
    reinterpret_cast<float4 *>(p)[i] = pv;
    reinterpret_cast<float4 *>(m)[i] = mv;
    reinterpret_cast<float4 *>(v)[i] = vv;
    reinterpret_cast<float4 *>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (shadow) {
        reinterpret_cast<__nv_bfloat162 *>(shadow)[2 * i] = __floats2bfloat162_rn(pv.x, pv.y);
        reinterpret_cast<__nv_bfloat162 *>(shadow)[2 * i + 1] = __floats2bfloat162_rn(pv.z, pv.w);
    }
}

// One launch stages a device-resident batch ([state; next_state] contiguous + the small vectors) instead of 5-9 D2D copies.
struct StageArgs {
    const uint4 *s, *s2; uint4 *obs; long long obs16;  // 16-byte units per observation block (state / next_state)
    const int *action; const float *reward; const uint8_t *terminal; const float *priority; const uint8_t *mask;
    const float *tau, *taup;
    int *d_action; float *d_reward; uint8_t *d_terminal; float *d_priority; uint8_t *d_mask; float *d_tau, *d_taup;
    int B, A;
    long long n_tau, n_taup;  // element counts of tau / tau_prime (IQN: B*N, B*N'; REM-DQN: E convex weights, 0)
    // optional: the zero-padded NHWC copy of the 4-channel frame stacks that the first convolution reads (what obs_to_nhwc4_x4_kernel
    // in net.cu would produce as the first launch of each net's forward), written here for the rows each net is about to consume:
    // rows [0, on_rows) of [state; next_state] for the online net, next_state for the target net.  quads = 4-pixel groups of all 2 B rows.
    void *nhwc_on, *nhwc_tg;
    long long quads;
    int on_rows, H, W, pad, scale255, in_f32, out_bf16;
};
// exact IEEE RN(v / 255) for the 256 values a uint8 frame can hold: q0 = RN(v r), rem = v - 255 q0 (exact), q = RN(q0 + rem r) with
// r = RN(1 / 255) -- three instructions instead of the ~25 of the IEEE divide's general path.  tests/test_host_logic.py checks all
// 256 values against exact rational arithmetic; the staging launch was issue-bound on the divides (ncu: 76 % issue slots busy).
__device__ __forceinline__ float div255_u8(float x) {
    constexpr float r = 1.0f / 255.0f;
    const float q0 = __fmul_rn(x, r);
    const float rem = __fmaf_rn(-q0, 255.f, x);
    return __fmaf_rn(rem, r, q0);
}
// One quad = 4 horizontally adjacent pixels x 4 channels of one observation row of [state; next_state]: four 4-byte (uint8) or 16-byte
// (fp32) loads, one 4-pixel run of the padded NHWC copy out (for the online net, the target net, or both).
template <typename Tin>
struct StageQuad {
    typedef typename std::conditional<sizeof(Tin) == 1, uint32_t, float4>::type Word;
    Word w[4];
    unsigned b, y, xx;
    bool to_on, to_tg;
};
template <typename Tin>
__device__ __forceinline__ void stage_quad_load(const StageArgs &a, long long i, StageQuad<Tin> &q) {
    const unsigned HW = (unsigned)(a.H * a.W), W = (unsigned)a.W;
    const unsigned qpi = HW >> 2;                    // quads per image (the host checks W % 4 == 0 and quads < 2^31)
    q.to_on = false; q.to_tg = false;
    if (i >= a.quads) return;
    q.b = (unsigned)i / qpi;                         // row of [state; next_state]
    const unsigned p = ((unsigned)i - q.b * qpi) << 2;
    q.to_on = (int)q.b < a.on_rows; q.to_tg = a.nhwc_tg && (int)q.b >= a.B;
    if (!q.to_on && !q.to_tg) return;
    const Tin *src = ((int)q.b < a.B ? reinterpret_cast<const Tin *>(a.s) + (size_t)q.b * 4 * HW : reinterpret_cast<const Tin *>(a.s2) + (size_t)(q.b - a.B) * 4 * HW) + p;
    q.y = p / W; q.xx = p - q.y * W;
#pragma unroll
    for (int c = 0; c < 4; ++c) q.w[c] = __ldg(reinterpret_cast<const typename StageQuad<Tin>::Word *>(src + (size_t)c * HW));
}
template <typename Tin, typename Tout>
__device__ __forceinline__ void stage_quad_store(const StageArgs &a, const StageQuad<Tin> &q) {
    if (!q.to_on && !q.to_tg) return;
    const size_t PH = a.H + 2 * a.pad, PW = a.W + 2 * a.pad;
    __align__(16) Tout o[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float f[4];
        if constexpr (sizeof(Tin) == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) f[j] = (float)((q.w[c] >> (8 * j)) & 0xffu);
        } else {
            f[0] = q.w[c].x; f[1] = q.w[c].y; f[2] = q.w[c].z; f[3] = q.w[c].w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = f[j];
            if (a.scale255) v = sizeof(Tin) == 1 ? div255_u8(v) : __fdiv_rn(v, 255.f);
            if (sizeof(Tout) == 2) reinterpret_cast<bf16 *>(o[j])[c] = __float2bfloat16_rn(v);
            else reinterpret_cast<float *>(o[j])[c] = v;
        }
    }
    typedef typename std::conditional<sizeof(Tout) == 2, uint2, uint4>::type Vec;  // one pixel = 4 channels
    auto put = [&](Tout *dst) {
        if (sizeof(Tout) == 2 && ((uintptr_t)dst & 15) == 0) {  // bf16: the quad is 32 bytes -- two 16-byte stores when the run is aligned
            *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(o[0]);
            *reinterpret_cast<uint4 *>(dst + 8) = *reinterpret_cast<const uint4 *>(o[2]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<Vec *>(dst + j * 4) = *reinterpret_cast<const Vec *>(o[j]);
        }
    };
    if (q.to_on) put(reinterpret_cast<Tout *>(a.nhwc_on) + ((q.b * PH + q.y + a.pad) * PW + q.xx + a.pad) * 4);
    if (q.to_tg) put(reinterpret_cast<Tout *>(a.nhwc_tg) + (((q.b - a.B) * PH + q.y + a.pad) * PW + q.xx + a.pad) * 4);
}
// two quads per thread (i and i + half), both loaded before either is converted: eight independent loads in flight per thread
template <typename Tin, typename Tout>
__device__ __forceinline__ void stage_nhwc_pair(const StageArgs &a, long long i) {
    const long long half = (a.quads + 1) >> 1;
    if (i >= half) return;
    StageQuad<Tin> q0, q1;
    stage_quad_load<Tin>(a, i, q0);
    stage_quad_load<Tin>(a, i + half, q1);
    stage_quad_store<Tin, Tout>(a, q0);
    stage_quad_store<Tin, Tout>(a, q1);
}
#ifndef ZOO_STAGE_MINB
#define ZOO_STAGE_MINB 4
#endif
__global__ void __launch_bounds__(256, ZOO_STAGE_MINB) stage_batch_kernel(StageArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.obs16) { a.obs[i] = __ldg(a.s + i); a.obs[a.obs16 + i] = __ldg(a.s2 + i); }
    if (a.nhwc_on) {
        if (a.in_f32) { if (a.out_bf16) stage_nhwc_pair<float, bf16>(a, i); else stage_nhwc_pair<float, float>(a, i); }
        else { if (a.out_bf16) stage_nhwc_pair<uint8_t, bf16>(a, i); else stage_nhwc_pair<uint8_t, float>(a, i); }
    }
    if (i < a.B) {
        a.d_action[i] = a.action[i]; a.d_reward[i] = a.reward[i]; a.d_terminal[i] = a.terminal[i];
        if (a.priority) a.d_priority[i] = a.priority[i];
    }
    if (a.mask && i < (long long)a.B * a.A) a.d_mask[i] = a.mask[i];
    if (a.tau && i < a.n_tau) a.d_tau[i] = a.tau[i];
    if (a.taup && i < a.n_taup) a.d_taup[i] = a.taup[i];
}

}  // namespace zoo

using namespace zoo;

struct zoo_opt {
    zoo_net *net;
    int kind;
    double eta, a, b, eps;
    float *s0 = nullptr, *s1 = nullptr;  // acc | m, v  (flat, internal layout)
    OptScalars *sc = nullptr;            // device
    long long host_t = 0;
};

struct zoo_learner {
    zoo_learner_cfg cfg;
    zoo_net *online, *target;
    zoo_opt *opt;
    zoo_dp *dp;
    cudaStream_t st;
    int B;
    int obs_dtype = -1;
    // device staging of the batch
    void *d_obs = nullptr;  // [2B][obs] : state rows 0..B-1, next_state rows B..2B-1
    int *d_action = nullptr; float *d_reward = nullptr; uint8_t *d_terminal = nullptr; float *d_priority = nullptr;
    uint8_t *d_mask = nullptr; float *d_tau = nullptr, *d_taup = nullptr;
    float *d_support = nullptr, *d_target = nullptr, *d_tdist = nullptr;
    float *d_per = nullptr, *d_prio = nullptr, *d_wraw = nullptr, *d_wmax = nullptr;
    float *d_tail_part = nullptr;  // multi-CTA DQN tail: 64 partial losses + the arrival ticket
    // the small per-sample vectors above live in one device block mirrored by a pinned host block: one H2D per update
    uint8_t *d_pack = nullptr, *h_pack = nullptr;
    size_t pack_bytes = 0;
    size_t off_action = 0, off_reward = 0, off_terminal = 0, off_priority = 0, off_mask = 0, off_tau = 0, off_taup = 0;
    cudaEvent_t ev_pack = nullptr;
    bool pack_inflight = false;
    float *d_loss = nullptr;  // lives at grads[n_flat] so one all-reduce covers gradients and loss
    float gamma_n;
    int last_launches = 0;
    // CUDA graph of the device part of compute_grads + apply
    int use_graph = 0;
    cudaGraphExec_t gexec = nullptr;
    int graph_key = -1;  // encodes (has_priority, has_mask, has_tau, obs_dtype, full)
    int graph_launches = 0;
    int warm = 0;
    bool target_forked = false;
    void *nccl_reg = nullptr;  // ncclCommRegister handle of the gradient buffer (data-parallel learners)
    // pipelined host batches (zoo_learner_submit / zoo_learner_collect): the H2D copies of update i + 1 run on a copy stream into
    // one of two staging slots while update i computes; one staging launch moves a slot into d_obs / d_pack on the learner's stream
    struct HostSlot {
        uint8_t *d_obs = nullptr, *d_pack = nullptr, *h_pack = nullptr;
        cudaEvent_t ev_copied = nullptr, ev_consumed = nullptr;
        bool used = false;
    } slot[2];
    int slot_next = 0;
    cudaStream_t st_copy = nullptr;
    static constexpr int PIPE = 4;     // submitted-but-not-collected updates
    float *h_res[PIPE] = {};           // pinned: loss, then B priorities
    cudaEvent_t ev_res[PIPE] = {};
    bool res_prio[PIPE] = {};
    int res_head = 0, res_count = 0;
};

static zoo_status opt_tick(zoo_opt *o, cudaStream_t st) {
    prof_label("opt");
    opt_tick_kernel<<<1, 1, 0, st>>>(o->sc, o->a, o->b);
    ZOO_LAUNCH_CHECK();
    o->host_t += 1;
    return ZOO_OK;
}

// update!(Q, gs) on the flat range [lo, hi) of the parameter / gradient / state buffers (multiples of 64 elements)
static zoo_status opt_range(zoo_opt *o, long long lo, long long hi, cudaStream_t st) {
    zoo_net *net = o->net;
    const long long n4 = (hi - lo) / 4;
    if (n4 <= 0) return ZOO_OK;
    prof_label("opt");
    bf16 *sh = net->shadow ? net->shadow + lo : nullptr;
    if (o->kind == ZOO_OPT_RMSPROP)
        ZOO_CUDA(launch_pdl(rmsprop_kernel, dim3(cdiv(n4, 256)), dim3(256), (size_t)0, st, net->params + lo, net->grads + lo, o->s0 + lo, sh, n4,
                            (float)o->a, (float)(1.0 - o->a), (float)o->eta, (float)o->eps));
    else
        ZOO_CUDA(launch_pdl(adam_kernel, dim3(cdiv(n4, 256)), dim3(256), (size_t)0, st, net->params + lo, net->grads + lo, o->s0 + lo, o->s1 + lo, sh,
                            n4, (float)o->a, (float)(1.0 - o->a), (float)o->b, (float)(1.0 - o->b), (float)o->eta, (float)o->eps,
                            (const OptScalars *)o->sc));
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

static zoo_status opt_step(zoo_opt *o, cudaStream_t st) {
    ZOO_TRY(opt_tick(o, st));
    return opt_range(o, 0, o->net->n_flat, st);
}

// Late bucket: everything from the largest weight tensor to the end of the flat buffers (fc1, fc2, ... and the loss slot)
// is final as soon as that layer's wgrad has run on the side lane, long before the conv stack's backward ends.  Its
// all-reduce and optimiser step are issued right there, on the side lane, overlapping the remaining backward; the small
// early bucket follows after the join.
struct LateCtx { zoo_learner *L; bool full; };
static zoo_status late_bucket(void *ctx_, cudaStream_t st);

// The target net's forward depends on nothing the online net produces (PrioritizedDQN / Rainbow / DQN): it runs on the
// online net's side lane, concurrently with the online forward, and is joined back before the loss head.
static zoo_status target_forward(zoo_learner *L, zoo_net *tg, const void *s2, const float *tau, int n_tau, cudaStream_t st) {
    zoo_net *on = L->online;
    const bool par = on->side && !g_prof_on;
    if (par) ZOO_TRY(net_side_fork(on, st));
    // opt-in (ZOO_TARGET_SHALLOW=1): measured slower on the B = 32 update (0.2466 vs 0.2411 ms) -- the side lane's CTAs then start earlier
    // and hold their SMs longer
    static const bool shallow = [] { const char *e = getenv("ZOO_TARGET_SHALLOW"); return e && e[0] == '1'; }();
    g_tc_shallow = par && shallow;
    zoo_status s = net_forward(tg, s2, L->obs_dtype, L->B, tau, n_tau, par ? on->side : st, 0);
    g_tc_shallow = 0;
    L->target_forked = par;
    return s;
}
static zoo_status target_join(zoo_learner *L, cudaStream_t st) {
    if (!L->target_forked) return ZOO_OK;
    L->target_forked = false;
    return net_side_join(L->online, st);
}

// everything between "batch is on the device" and "results are on the device"
static zoo_status enqueue_grads(zoo_learner *L, bool has_prio, bool has_mask, bool has_tau, bool full) {
    const zoo_learner_cfg &c = L->cfg;
    zoo_net *on = L->online, *tg = L->target;
    cudaStream_t st = L->st;
    const int B = L->B;
    const size_t obs_row = (size_t)obs_elems(on) * (L->obs_dtype == ZOO_OBS_U8 ? 1 : 4);
    const void *s_ptr = L->d_obs;
    const void *s2_ptr = (const uint8_t *)L->d_obs + (size_t)B * obs_row;
    const int world = L->dp ? dp_world(L->dp) : 1;
    if (g_prof_on) { prof_label("stage"); prof_mark("stage_batch", __LINE__); }
    LateCtx late{L, full};
    const bool split = on->late_off > 0;
    on->late_cb = split ? late_bucket : nullptr;
    on->late_ctx = &late;
    if (full) {  // this step's bias corrections: off the critical path, first thing on the side lane
        const bool par = on->side && !g_prof_on;
        if (par) ZOO_TRY(net_side_fork(on, st));
        ZOO_TRY(opt_tick(L->opt, par ? on->side : st));
    }
    prof_label("head");
    HeadArgs h;
    h.B = B; h.A = c.n_actions; h.action = L->d_action; h.reward = L->d_reward; h.terminal = L->d_terminal;
    h.mask = has_mask ? L->d_mask : nullptr; h.w_raw = nullptr; h.w_max = nullptr;
    h.inv_B = 1.0f / (float)((long long)B * world); h.per_sample = L->d_per;
    const bool per_kind = c.kind == ZOO_LEARNER_PRIORITIZED_DQN || ((c.kind == ZOO_LEARNER_RAINBOW || c.kind == ZOO_LEARNER_IQN) && has_prio);
    if (per_kind) {
        prof_label("head"); ZOO_TRY(launch_per_weights(L->d_priority, c.beta_priority, B, L->d_wraw, L->d_wmax, st));
        if (L->dp) ZOO_TRY(dp_allreduce_max(L->dp, L->d_wmax, 1, st));
        h.w_raw = L->d_wraw; h.w_max = L->d_wmax;
    }
    int bwd_rows = B, n_tau = 0;
    // DQN family at Dopamine batch sizes: head layer forward + TD head + loss + head layer backward in ONE launch (narrow.cu)
    bool tail = false;
    if ((c.kind == ZOO_LEARNER_BASIC_DQN || c.kind == ZOO_LEARNER_DQN || c.kind == ZOO_LEARNER_PRIORITIZED_DQN) && on->desc.kind == ZOO_NET_CHAIN &&
        on->a.layers.size() >= 2) {
        // ZOO_DQN_TAIL: 0 (default) = four launches (narrow_fwd, dqn_head, loss_reduce, narrow_bwd); 2 = the multi-CTA tail (each CTA owns
        // 2048 / K samples; 20 us eager against 37 us for the four, but the graph-replayed update does not get faster: 0.2538 vs
        // 0.2508 ms, the four small kernels leave more SMs to the lanes running beside them); 1 = the single-CTA tail (~26 us).
        // All three are parity-tested (tests/test_gpu_gemm.py::test_non_default_kernel_paths_keep_parity).
        static const int tail_mode = [] { const char *e = getenv("ZOO_DQN_TAIL"); return e ? atoi(e) : 0; }();
        const LayerExec &Lh = on->a.layers.back();
        const int dbl = c.kind == ZOO_LEARNER_DQN && c.double_dqn;
        const int rows_fwd = (c.kind == ZOO_LEARNER_BASIC_DQN || dbl) ? 2 * B : B;
        const bool tail_ok = Lh.kind == ZOO_LAYER_DENSE && Lh.out_fp32 && !Lh.relu && !(split && on->tensors[Lh.wt].offset == on->late_off);
        const bool multi = tail_mode == 2 && tail_ok && dqn_tail_multi_supported(c.n_actions, Lh.K, on->prec == ZOO_PREC_BF16, B);
        tail = multi || (tail_mode == 1 && tail_ok && dqn_tail_supported(c.n_actions, Lh.K, on->prec == ZOO_PREC_BF16, rows_fwd));
        if (tail) {
            const LayerExec &Lp = on->a.layers[on->a.layers.size() - 2];
            bwd_rows = c.kind == ZOO_LEARNER_BASIC_DQN ? 2 * B : B;
            if (c.kind != ZOO_LEARNER_BASIC_DQN) ZOO_TRY(target_forward(L, tg, s2_ptr, nullptr, 0, st));  // side lane, rejoined before the tail
            on->skip_head = true;
            zoo_status s = net_forward(on, s_ptr, L->obs_dtype, rows_fwd, nullptr, 0, st, bwd_rows);
            if (s == ZOO_OK && c.kind != ZOO_LEARNER_BASIC_DQN) s = target_join(L, st);
            if (s == ZOO_OK) {
                prof_label("tail:t%d[%dx%dx%lld]", Lh.wt, rows_fwd, c.n_actions, Lh.K);
                if (multi)
                    s = dqn_tail_multi(c.kind, dbl, h, Lp.out, on->prec == ZOO_PREC_BF16, on->params + on->tensors[Lh.wt].offset,
                                       on->params + on->tensors[Lh.bt].offset, tg ? tg->out_final : nullptr,
                                       c.kind == ZOO_LEARNER_BASIC_DQN ? c.gamma : L->gamma_n, c.beta_priority, rows_fwd == 2 * B,
                                       c.kind == ZOO_LEARNER_BASIC_DQN, Lh.K, Lp.relu, on->out_final, on->dout_final, Lp.dout,
                                       on->grads + on->tensors[Lh.wt].offset, on->grads + on->tensors[Lh.bt].offset, on->grads + on->tensors[Lp.bt].offset,
                                       L->d_loss, L->d_prio, L->d_tail_part, (unsigned *)(L->d_tail_part + 64), st);
                else
                s = dqn_tail(c.kind, dbl, h, Lp.out, on->prec == ZOO_PREC_BF16, on->params + on->tensors[Lh.wt].offset, on->params + on->tensors[Lh.bt].offset,
                             tg ? tg->out_final : nullptr, c.kind == ZOO_LEARNER_BASIC_DQN ? c.gamma : L->gamma_n, c.beta_priority, rows_fwd, bwd_rows, Lh.K,
                             Lp.relu, on->out_final, on->dout_final, Lp.dout, on->grads + on->tensors[Lh.wt].offset, on->grads + on->tensors[Lh.bt].offset,
                             on->grads + on->tensors[Lp.bt].offset, L->d_loss, L->d_prio, st);
            }
            if (s == ZOO_OK) s = net_backward(on, L->obs_dtype, s_ptr, bwd_rows, 0, st);
            on->skip_head = false;
            on->late_cb = nullptr;
            ZOO_TRY(s);
        }
    }
    if (tail) {
    } else if (c.kind == ZOO_LEARNER_BASIC_DQN) {
        ZOO_TRY(net_forward(on, s_ptr, L->obs_dtype, 2 * B, nullptr, 0, st, 2 * B));  // Q([s; s'])
        prof_label("head"); ZOO_TRY(launch_dqn_head(c.kind, 0, h, on->out_final, nullptr, c.gamma, on->dout_final, st));
        bwd_rows = 2 * B;
    } else if (c.kind == ZOO_LEARNER_DQN || c.kind == ZOO_LEARNER_PRIORITIZED_DQN) {
        const int dbl = c.kind == ZOO_LEARNER_DQN && c.double_dqn;
        ZOO_TRY(target_forward(L, tg, s2_ptr, nullptr, 0, st));  // side lane, rejoined before the head
        ZOO_TRY(net_forward(on, s_ptr, L->obs_dtype, dbl ? 2 * B : B, nullptr, 0, st, B));
        ZOO_TRY(target_join(L, st));
        prof_label("head"); ZOO_TRY(launch_dqn_head(c.kind, dbl, h, on->out_final, tg->out_final, L->gamma_n, on->dout_final, st));
    } else if (c.kind == ZOO_LEARNER_RAINBOW) {
        ZOO_TRY(target_forward(L, tg, s2_ptr, nullptr, 0, st));
        ZOO_TRY(net_forward(on, s_ptr, L->obs_dtype, B, nullptr, 0, st, B));
        ZOO_TRY(target_join(L, st));
        prof_label("head"); ZOO_TRY(launch_c51_head(h, c.n_atoms, tg->out_final, on->out_final, L->d_support, c.delta_z, c.v_min, c.v_max, L->gamma_n,
                                on->dout_final, L->d_tdist, st));
    } else if (c.kind == ZOO_LEARNER_QR_DQN) {
        ZOO_TRY(target_forward(L, tg, s2_ptr, nullptr, 0, st));
        ZOO_TRY(net_forward(on, s_ptr, L->obs_dtype, B, nullptr, 0, st, B));
        ZOO_TRY(target_join(L, st));
        prof_label("head"); ZOO_TRY(launch_qr_target(h, c.N, tg->out_final, c.gamma, L->d_target, st));
        ZOO_TRY(launch_qr_loss(h, c.N, on->out_final, L->d_target, c.kappa, on->dout_final, st));
    } else if (c.kind == ZOO_LEARNER_REM_DQN) {
        ZOO_REQUIRE(has_tau, "update: REMDQNLearner needs the convex weights in batch.tau (ensemble_num floats)");
        ZOO_TRY(target_forward(L, tg, s2_ptr, nullptr, 0, st));
        ZOO_TRY(net_forward(on, s_ptr, L->obs_dtype, B, nullptr, 0, st, B));
        ZOO_TRY(target_join(L, st));
        prof_label("head"); ZOO_TRY(launch_rem_head(h, c.N, on->out_final, tg->out_final, L->d_tau, L->gamma_n, on->dout_final, st));
    } else {  // IQN
        if (!has_tau) {
            prof_label("head"); ZOO_TRY(launch_tau(L->d_taup, (long long)B * c.N_prime, c.seed, &L->opt->sc->rng, 0, st));
            ZOO_TRY(launch_tau(L->d_tau, (long long)B * c.N, c.seed, &L->opt->sc->rng, 1, st));
            rng_tick_kernel<<<1, 1, 0, st>>>(L->opt->sc);
            ZOO_LAUNCH_CHECK();
        }
        ZOO_TRY(net_forward(tg, s2_ptr, L->obs_dtype, B, L->d_taup, c.N_prime, st, 0));
        prof_label("head"); ZOO_TRY(launch_iqn_target(h, c.N_prime, tg->out_final, c.gamma, L->d_target, st));
        ZOO_TRY(net_forward(on, s_ptr, L->obs_dtype, B, L->d_tau, c.N, st, B));
        prof_label("head"); ZOO_TRY(launch_iqn_loss(h, c.N, c.N_prime, on->out_final, L->d_target, L->d_tau, c.kappa, on->dout_final, st));
        n_tau = c.N;
    }
    if (!tail) {
        // ZOO_LOSS_ASIDE=1 (opt-in): the scalar loss and the new priorities hang off the head's per-sample losses and nothing downstream
        // reads them before the update ends, so they can run on the side lane (first thing there, ahead of the late bucket's
        // all-reduce, which carries the loss slot) and the backward's first kernel follows the head directly: 0.2203 vs 0.2233 ms for
        // the B = 32 update.  Not the default: with it the submit / collect pipeline test of the PER learner failed in 2 of 4 runs
        // (not yet understood), and the round's evidence was measured without it.
        static const bool loss_aside = [] { const char *e = getenv("ZOO_LOSS_ASIDE"); return e && e[0] == '1'; }();
        const bool aside = loss_aside && on->side && !g_prof_on;
        if (aside) ZOO_TRY(net_side_fork(on, st));
        prof_label("head"); ZOO_TRY(launch_loss_reduce(h, c.beta_priority, L->d_loss, L->d_prio, aside ? on->side : st));
        ZOO_TRY(net_backward(on, L->obs_dtype, s_ptr, bwd_rows, n_tau, st));  // (rejoins the side lane at its end)
    }
    on->late_cb = nullptr;
    const long long early_end = split ? on->late_off : on->n_flat + 64;  // gradients (+ loss, in the last bucket) per collective
    if (L->dp) ZOO_TRY(dp_allreduce_sum(L->dp, on->grads, early_end, st));
    if (full) ZOO_TRY(opt_range(L->opt, 0, split ? on->late_off : on->n_flat, st));
    return ZOO_OK;
}

static zoo_status late_bucket(void *ctx_, cudaStream_t st) {
    LateCtx *c = (LateCtx *)ctx_;
    zoo_learner *L = c->L;
    zoo_net *on = L->online;
    if (L->dp) ZOO_TRY(dp_allreduce_sum(L->dp, on->grads + on->late_off, on->n_flat + 64 - on->late_off, st));
    if (c->full) ZOO_TRY(opt_range(L->opt, on->late_off, on->n_flat, st));
    return ZOO_OK;
}

static zoo_status pipe_alloc(zoo_learner *L) {
    if (L->st_copy) return ZOO_OK;
    ZOO_CUDA(cudaStreamCreateWithFlags(&L->st_copy, cudaStreamNonBlocking));
    const size_t obs_cap = (size_t)2 * L->B * obs_elems(L->online) * sizeof(float);
    for (auto &sl : L->slot) {
        ZOO_CUDA(cudaMalloc(&sl.d_obs, obs_cap));
        ZOO_CUDA(cudaMalloc(&sl.d_pack, L->pack_bytes));
        ZOO_CUDA(cudaHostAlloc(&sl.h_pack, L->pack_bytes, cudaHostAllocDefault));
        ZOO_CUDA(cudaEventCreateWithFlags(&sl.ev_copied, cudaEventDisableTiming));
        ZOO_CUDA(cudaEventCreateWithFlags(&sl.ev_consumed, cudaEventDisableTiming));
    }
    for (int i = 0; i < zoo_learner::PIPE; ++i) {
        ZOO_CUDA(cudaHostAlloc(&L->h_res[i], (size_t)(1 + L->B) * sizeof(float), cudaHostAllocDefault));
        ZOO_CUDA(cudaEventCreateWithFlags(&L->ev_res[i], cudaEventDisableTiming));
    }
    return ZOO_OK;
}

static zoo_status stage_batch(zoo_learner *L, const zoo_batch *b, bool pipelined = false) {
    const zoo_learner_cfg &c = L->cfg;
    cudaStream_t st = L->st;
    const int B = L->B;
    ZOO_REQUIRE(b && b->B == B, "update: batch size %d != learner batch_size %d", b ? b->B : -1, B);
    ZOO_REQUIRE(b->obs_dtype == ZOO_OBS_U8 || b->obs_dtype == ZOO_OBS_F32, "update: bad obs dtype");
    ZOO_REQUIRE(b->state && b->next_state && b->action && b->reward && b->terminal, "update: state/next_state/action/reward/terminal are required");
    if (c.kind == ZOO_LEARNER_PRIORITIZED_DQN) ZOO_REQUIRE(b->priority, "update: PrioritizedDQNLearner needs batch.priority");
    L->obs_dtype = b->obs_dtype;
    const size_t obs_bytes = (size_t)B * obs_elems(L->online) * (b->obs_dtype == ZOO_OBS_U8 ? 1 : 4);
    if (c.kind == ZOO_LEARNER_IQN) ZOO_REQUIRE((b->tau != nullptr) == (b->tau_prime != nullptr), "update: pass both tau and tau_prime or neither");
    zoo_batch via;  // pipelined host batch: what the staging launch below reads (the slot's device copies)
    if (pipelined && !b->on_device && obs_bytes % 16 == 0) {
        ZOO_TRY(pipe_alloc(L));
        auto &sl = L->slot[L->slot_next];
        L->slot_next ^= 1;
        if (sl.used) ZOO_CUDA(cudaEventSynchronize(sl.ev_consumed));  // the staging launch of two submits ago has read the slot
        uint8_t *h = sl.h_pack;
        memcpy(h + L->off_action, b->action, (size_t)B * sizeof(int));
        memcpy(h + L->off_reward, b->reward, (size_t)B * sizeof(float));
        memcpy(h + L->off_terminal, b->terminal, (size_t)B);
        if (b->priority) memcpy(h + L->off_priority, b->priority, (size_t)B * sizeof(float));
        if (b->next_mask) memcpy(h + L->off_mask, b->next_mask, (size_t)B * c.n_actions);
        const bool has_tau = (c.kind == ZOO_LEARNER_IQN || c.kind == ZOO_LEARNER_REM_DQN) && b->tau;
        if (c.kind == ZOO_LEARNER_IQN && b->tau) {
            memcpy(h + L->off_tau, b->tau, (size_t)B * c.N * sizeof(float));
            memcpy(h + L->off_taup, b->tau_prime, (size_t)B * c.N_prime * sizeof(float));
        }
        if (c.kind == ZOO_LEARNER_REM_DQN && b->tau) memcpy(h + L->off_tau, b->tau, (size_t)c.N * sizeof(float));
        const size_t used = has_tau ? L->pack_bytes : (L->off_tau ? L->off_tau : L->pack_bytes);
        ZOO_CUDA(cudaMemcpyAsync(sl.d_obs, b->state, obs_bytes, cudaMemcpyHostToDevice, L->st_copy));
        ZOO_CUDA(cudaMemcpyAsync(sl.d_obs + obs_bytes, b->next_state, obs_bytes, cudaMemcpyHostToDevice, L->st_copy));
        ZOO_CUDA(cudaMemcpyAsync(sl.d_pack, h, used, cudaMemcpyHostToDevice, L->st_copy));
        ZOO_CUDA(cudaEventRecord(sl.ev_copied, L->st_copy));
        ZOO_CUDA(cudaStreamWaitEvent(st, sl.ev_copied, 0));
        via = *b;
        via.on_device = 1;
        via.state = sl.d_obs; via.next_state = sl.d_obs + obs_bytes;
        via.action = (const int32_t *)(sl.d_pack + L->off_action); via.reward = (const float *)(sl.d_pack + L->off_reward);
        via.terminal = sl.d_pack + L->off_terminal;
        via.priority = b->priority ? (const float *)(sl.d_pack + L->off_priority) : nullptr;
        via.next_mask = b->next_mask ? sl.d_pack + L->off_mask : nullptr;
        via.tau = has_tau ? (const float *)(sl.d_pack + L->off_tau) : nullptr;
        via.tau_prime = (c.kind == ZOO_LEARNER_IQN && b->tau) ? (const float *)(sl.d_pack + L->off_taup) : nullptr;
        sl.used = true;
        b = &via;
    }
    zoo_learner::HostSlot *const consumed = (b == &via) ? &L->slot[L->slot_next ^ 1] : nullptr;
    if (b->on_device && obs_bytes % 16 == 0 && ((uintptr_t)b->state & 15) == 0 && ((uintptr_t)b->next_state & 15) == 0) {
        StageArgs a{};
        a.s = (const uint4 *)b->state; a.s2 = (const uint4 *)b->next_state; a.obs = (uint4 *)L->d_obs; a.obs16 = (long long)(obs_bytes / 16);
        a.action = b->action; a.reward = b->reward; a.terminal = b->terminal; a.priority = b->priority; a.mask = b->next_mask;
        const bool rem = c.kind == ZOO_LEARNER_REM_DQN;
        a.tau = (c.kind == ZOO_LEARNER_IQN || rem) ? b->tau : nullptr; a.taup = c.kind == ZOO_LEARNER_IQN ? b->tau_prime : nullptr;
        a.n_tau = rem ? c.N : (long long)B * c.N; a.n_taup = (long long)B * c.N_prime;
        a.d_action = L->d_action; a.d_reward = L->d_reward; a.d_terminal = L->d_terminal; a.d_priority = L->d_priority; a.d_mask = L->d_mask;
        a.d_tau = L->d_tau; a.d_taup = L->d_taup; a.B = B; a.A = c.n_actions;
        long long n = std::max<long long>(a.obs16, (long long)B * std::max(std::max(c.n_actions, c.N), std::max(c.N_prime, 1)));
        // the first convolution's padded NHWC input, written by this launch instead of by the first launch of each net's forward
        // (one launch less at the head of the update's critical path, and next_state is repacked once for both nets)
        a.nhwc_on = nullptr; a.nhwc_tg = nullptr; a.quads = 0;
        zoo_net *on = L->online, *tg = L->target;
        on->prestaged_src = nullptr; on->prestaged_rows = 0;
        if (tg) { tg->prestaged_src = nullptr; tg->prestaged_rows = 0; }
        static const bool prestage = [] { const char *e = getenv("ZOO_PRESTAGE"); return !(e && e[0] == '0'); }();
        if (prestage && on->obs_nhwc && on->desc.in_c == 4 && on->desc.in_w % 4 == 0 && (long long)2 * B * on->desc.in_h * on->desc.in_w < (1ll << 32) && !on->a.layers.empty() && on->a.layers[0].first_raw &&
            (!tg || (tg->obs_nhwc && tg->prec == on->prec))) {
            const zoo::LayerExec &L0 = on->a.layers[0];
            const bool dbl = (c.kind == ZOO_LEARNER_DQN && c.double_dqn) || c.kind == ZOO_LEARNER_BASIC_DQN;
            a.on_rows = (int)std::min<long long>(dbl ? 2 * B : B, on->a.rows_cap);
            a.H = on->desc.in_h; a.W = on->desc.in_w; a.pad = (L0.ih - on->desc.in_h) / 2; a.scale255 = L0.scale255;
            a.in_f32 = b->obs_dtype == ZOO_OBS_F32; a.out_bf16 = on->prec == ZOO_PREC_BF16;
            a.nhwc_on = on->obs_nhwc_buf; a.nhwc_tg = tg ? tg->obs_nhwc_buf : nullptr;
            a.quads = (long long)2 * B * a.H * a.W / 4;  // < 2^31: the kernel indexes quads with 32-bit arithmetic
            on->prestaged_src = L->d_obs; on->prestaged_rows = a.on_rows;
            if (tg) { tg->prestaged_src = (const uint8_t *)L->d_obs + obs_bytes; tg->prestaged_rows = B; }
            // Both nets now read the padded NHWC copy only (the first layer is a convolution; its weight gradient reads the same copy), so
            // the raw [state; next_state] block is not duplicated into d_obs any more -- d_obs stays the identity of the staged batch
            // (prestaged_src).  ZOO_STAGE_RAW=1 keeps the copy (A/B; Rainbow B = 1024: 116 MB of traffic per update).
            static const bool keep_raw = [] { const char *e = getenv("ZOO_STAGE_RAW"); return e && e[0] == '1'; }();
            if (!keep_raw && a.on_rows >= (dbl ? 2 * B : B)) {
                a.obs16 = 0;
                n = (long long)B * std::max(std::max(c.n_actions, c.N), std::max(c.N_prime, 1));
            }
            // (8-byte loads -- 8 pixels per thread and channel -- were measured slower: 80 vs 64 us at B = 1024, the 64-byte store stride costs more than the loads gain)
            n = std::max(n, (a.quads + 1) / 2);  // two quads per thread
        }
        stage_batch_kernel<<<cdiv(n, 256), 256, 0, st>>>(a);
        ZOO_LAUNCH_CHECK();
        if (consumed) ZOO_CUDA(cudaEventRecord(consumed->ev_consumed, st));
        return ZOO_OK;
    }
    L->online->prestaged_src = nullptr; L->online->prestaged_rows = 0;
    if (L->target) { L->target->prestaged_src = nullptr; L->target->prestaged_rows = 0; }
    if (!b->on_device) {
        // host batch: the two frame blocks go straight from the caller's arrays; the small vectors are packed into the pinned
        // mirror of d_pack and cross PCIe in one copy (seven tiny copies cost ~5 us of launch latency each)
        ZOO_CUDA(cudaMemcpyAsync(L->d_obs, b->state, obs_bytes, cudaMemcpyHostToDevice, st));
        ZOO_CUDA(cudaMemcpyAsync((uint8_t *)L->d_obs + obs_bytes, b->next_state, obs_bytes, cudaMemcpyHostToDevice, st));
        if (L->pack_inflight) ZOO_CUDA(cudaEventSynchronize(L->ev_pack));  // the previous update's copy has read h_pack
        memcpy(L->h_pack + L->off_action, b->action, (size_t)B * sizeof(int));
        memcpy(L->h_pack + L->off_reward, b->reward, (size_t)B * sizeof(float));
        memcpy(L->h_pack + L->off_terminal, b->terminal, (size_t)B);
        if (b->priority) memcpy(L->h_pack + L->off_priority, b->priority, (size_t)B * sizeof(float));
        if (b->next_mask) memcpy(L->h_pack + L->off_mask, b->next_mask, (size_t)B * c.n_actions);
        if (c.kind == ZOO_LEARNER_IQN && b->tau) {
            memcpy(L->h_pack + L->off_tau, b->tau, (size_t)B * c.N * sizeof(float));
            memcpy(L->h_pack + L->off_taup, b->tau_prime, (size_t)B * c.N_prime * sizeof(float));
        }
        if (c.kind == ZOO_LEARNER_REM_DQN && b->tau) memcpy(L->h_pack + L->off_tau, b->tau, (size_t)c.N * sizeof(float));
        const size_t used = (c.kind == ZOO_LEARNER_IQN && !b->tau) ? L->off_tau : L->pack_bytes;  // device-drawn tau stays untouched
        ZOO_CUDA(cudaMemcpyAsync(L->d_pack, L->h_pack, used, cudaMemcpyHostToDevice, st));
        ZOO_CUDA(cudaEventRecord(L->ev_pack, st));
        L->pack_inflight = true;
        return ZOO_OK;
    }
    const cudaMemcpyKind k = cudaMemcpyDeviceToDevice;  // device batch whose frame blocks are not 16-byte aligned
    ZOO_CUDA(cudaMemcpyAsync(L->d_obs, b->state, obs_bytes, k, st));
    ZOO_CUDA(cudaMemcpyAsync((uint8_t *)L->d_obs + obs_bytes, b->next_state, obs_bytes, k, st));
    ZOO_CUDA(cudaMemcpyAsync(L->d_action, b->action, B * sizeof(int), k, st));
    ZOO_CUDA(cudaMemcpyAsync(L->d_reward, b->reward, B * sizeof(float), k, st));
    ZOO_CUDA(cudaMemcpyAsync(L->d_terminal, b->terminal, B, k, st));
    if (b->priority) ZOO_CUDA(cudaMemcpyAsync(L->d_priority, b->priority, B * sizeof(float), k, st));
    if (b->next_mask) ZOO_CUDA(cudaMemcpyAsync(L->d_mask, b->next_mask, (size_t)B * c.n_actions, k, st));
    if (c.kind == ZOO_LEARNER_IQN && b->tau) {
        ZOO_CUDA(cudaMemcpyAsync(L->d_tau, b->tau, (size_t)B * c.N * sizeof(float), k, st));
        ZOO_CUDA(cudaMemcpyAsync(L->d_taup, b->tau_prime, (size_t)B * c.N_prime * sizeof(float), k, st));
    }
    if (c.kind == ZOO_LEARNER_REM_DQN && b->tau) ZOO_CUDA(cudaMemcpyAsync(L->d_tau, b->tau, (size_t)c.N * sizeof(float), k, st));
    return ZOO_OK;
}

static zoo_status run_update(zoo_learner *L, const zoo_batch *b, float *loss_out, float *prio_out, bool full, bool pipelined = false) {
    ZOO_REQUIRE(L, "update: null learner");
    ZOO_TRY(stage_batch(L, b, pipelined));
    // the gradient buffer must be zero on entry; the optimiser kernels leave it so, anything else (creation, a
    // compute_grads without apply, a set of optimiser state) marks it dirty and costs one memset here, outside the graph
    if (L->online->grads_dirty) {
        ZOO_CUDA(cudaMemsetAsync(L->online->grads, 0, (L->online->n_flat + 64) * sizeof(float), L->st));
        L->online->grads_dirty = false;
    }
    if (!full) L->online->grads_dirty = true;
    const bool has_prio = b->priority != nullptr, has_mask = b->next_mask != nullptr, has_tau = b->tau != nullptr;
    const int key = (has_prio ? 1 : 0) | (has_mask ? 2 : 0) | (has_tau ? 4 : 0) | (L->obs_dtype << 3) | (full ? 16 : 0) |
                    (L->online->prestaged_src ? 64 : 0);  // (the captured forward skips or keeps its own repack launch accordingly)
    g_launches = 0;
    if (L->use_graph && L->warm) {
        if (!L->gexec || L->graph_key != key) {
            if (L->gexec) { cudaGraphExecDestroy(L->gexec); L->gexec = nullptr; }
            cudaGraph_t graph;
            ZOO_CUDA(cudaStreamBeginCapture(L->st, cudaStreamCaptureModeThreadLocal));
            zoo_status s = enqueue_grads(L, has_prio, has_mask, has_tau, full);
            cudaError_t ce = cudaStreamEndCapture(L->st, &graph);
            if (s != ZOO_OK) { if (ce == cudaSuccess) cudaGraphDestroy(graph); return s; }
            ZOO_CUDA(ce);
            ZOO_CUDA(cudaGraphInstantiate(&L->gexec, graph, 0));
            cudaGraphDestroy(graph);
            L->graph_key = key;
            L->graph_launches = g_launches;
            if (full) L->opt->host_t -= 1;  // the capture pass did not execute
        }
        ZOO_CUDA(cudaGraphLaunch(L->gexec, L->st));
        if (full) L->opt->host_t += 1;
        L->last_launches = L->graph_launches;
    } else {
        ZOO_TRY(enqueue_grads(L, has_prio, has_mask, has_tau, full));
        L->last_launches = g_launches;
        L->warm = 1;
    }
    ZOO_TRY(net_mark_busy(L->online, L->st));  // net-level entry points on the nets' own streams wait for this update (net.cuh)
    if (L->target) ZOO_TRY(net_mark_busy(L->target, L->st));
    if (prio_out) ZOO_CUDA(cudaMemcpyAsync(prio_out, L->d_prio, L->B * sizeof(float), cudaMemcpyDeviceToHost, L->st));
    if (loss_out) ZOO_CUDA(cudaMemcpyAsync(loss_out, L->d_loss, sizeof(float), cudaMemcpyDeviceToHost, L->st));
    if (prio_out || loss_out) ZOO_CUDA(cudaStreamSynchronize(L->st));
    return ZOO_OK;
}

extern "C" {

zoo_status zoo_opt_create(zoo_net *net, int32_t kind, double eta, double a, double b, double eps, zoo_opt **out) {
    ZOO_REQUIRE(net && out, "zoo_opt_create: null argument");
    ZOO_REQUIRE(kind == ZOO_OPT_RMSPROP || kind == ZOO_OPT_ADAM, "zoo_opt_create: unknown optimiser %d", kind);
    ZOO_TRY(net_alloc_backward(net));
    zoo_opt *o = new zoo_opt();
    o->net = net; o->kind = kind; o->eta = eta; o->a = a; o->b = kind == ZOO_OPT_ADAM ? b : 0.0; o->eps = eps;
    ZOO_CUDA(cudaMalloc(&o->s0, net->n_flat * sizeof(float)));
    ZOO_CUDA(cudaMemset(o->s0, 0, net->n_flat * sizeof(float)));
    if (kind == ZOO_OPT_ADAM) {
        ZOO_CUDA(cudaMalloc(&o->s1, net->n_flat * sizeof(float)));
        ZOO_CUDA(cudaMemset(o->s1, 0, net->n_flat * sizeof(float)));
    }
    ZOO_CUDA(cudaMalloc(&o->sc, sizeof(OptScalars)));
    OptScalars h{};
    h.b1p = o->a; h.b2p = o->b; h.t = 0; h.rng = 0;
    ZOO_CUDA(cudaMemcpy(o->sc, &h, sizeof(h), cudaMemcpyHostToDevice));
    *out = o;
    return ZOO_OK;
}

zoo_status zoo_opt_destroy(zoo_opt *o) {
    if (!o) return ZOO_OK;
    cudaDeviceSynchronize();
    cudaFree(o->s0); if (o->s1) cudaFree(o->s1); cudaFree(o->sc);
    delete o;
    return ZOO_OK;
}

static float *opt_slot(const zoo_opt *o, int slot) { return slot == 0 ? o->s0 : (slot == 1 ? o->s1 : nullptr); }

zoo_status zoo_opt_get_state(const zoo_opt *o, int32_t slot, int32_t id, float *host, int64_t n) {
    ZOO_REQUIRE(o && opt_slot(o, slot), "zoo_opt_get_state: bad optimiser/slot");
    return zoo_internal_tensor_io(o->net, opt_slot(o, slot), id, host, n, 0);
}
zoo_status zoo_opt_set_state(zoo_opt *o, int32_t slot, int32_t id, const float *host, int64_t n) {
    ZOO_REQUIRE(o && opt_slot(o, slot), "zoo_opt_set_state: bad optimiser/slot");
    return zoo_internal_tensor_io(o->net, opt_slot(o, slot), id, const_cast<float *>(host), n, 1);
}
zoo_status zoo_opt_get_step(const zoo_opt *o, int64_t *t) {
    ZOO_REQUIRE(o && t, "zoo_opt_get_step: null argument");
    *t = o->host_t;
    return ZOO_OK;
}
zoo_status zoo_opt_set_step(zoo_opt *o, int64_t t) {
    ZOO_REQUIRE(o && t >= 0, "zoo_opt_set_step: bad argument");
    OptScalars h{};
    ZOO_CUDA(cudaDeviceSynchronize());
    ZOO_CUDA(cudaMemcpy(&h, o->sc, sizeof(h), cudaMemcpyDeviceToHost));
    // beta^(t+1) by the same repeated fp64 multiplication the run carried (opt_tick_kernel), so that a restored ADAM state is
    // bit-identical to the one that was saved (std::pow can differ in the last bit)
    double b1p = o->a, b2p = o->b;
    for (int64_t i = 0; i < t; ++i) { b1p *= o->a; b2p *= o->b; }
    h.b1p = b1p; h.b2p = b2p; h.t = (unsigned long long)t;
    ZOO_CUDA(cudaMemcpy(o->sc, &h, sizeof(h), cudaMemcpyHostToDevice));
    o->host_t = t;
    return ZOO_OK;
}

// counter of the on-device tau stream (IQN: one tick per update that drew its own quantiles); saved / restored by checkpoints
zoo_status zoo_opt_get_rng(const zoo_opt *o, uint64_t *counter) {
    ZOO_REQUIRE(o && counter, "zoo_opt_get_rng: null argument");
    OptScalars h{};
    ZOO_CUDA(cudaDeviceSynchronize());
    ZOO_CUDA(cudaMemcpy(&h, o->sc, sizeof(h), cudaMemcpyDeviceToHost));
    *counter = h.rng;
    return ZOO_OK;
}
zoo_status zoo_opt_set_rng(zoo_opt *o, uint64_t counter) {
    ZOO_REQUIRE(o, "zoo_opt_set_rng: null argument");
    OptScalars h{};
    ZOO_CUDA(cudaDeviceSynchronize());
    ZOO_CUDA(cudaMemcpy(&h, o->sc, sizeof(h), cudaMemcpyDeviceToHost));
    h.rng = counter;
    ZOO_CUDA(cudaMemcpy(o->sc, &h, sizeof(h), cudaMemcpyHostToDevice));
    return ZOO_OK;
}

zoo_status zoo_learner_create(const zoo_learner_cfg *cfg, zoo_net *online, zoo_net *target, zoo_opt *opt, zoo_dp *dp, zoo_learner **out) {
    ZOO_REQUIRE(cfg && online && opt && out, "zoo_learner_create: null argument");
    ZOO_REQUIRE(cfg->kind >= ZOO_LEARNER_BASIC_DQN && cfg->kind <= ZOO_LEARNER_REM_DQN, "zoo_learner_create: unknown learner kind %d", cfg->kind);
    ZOO_REQUIRE(cfg->kind == ZOO_LEARNER_BASIC_DQN || target, "zoo_learner_create: this learner needs a target network");
    ZOO_REQUIRE(opt->net == online, "zoo_learner_create: the optimiser must belong to the online network");
    ZOO_REQUIRE(cfg->batch_size > 0 && cfg->batch_size <= online->max_batch, "zoo_learner_create: batch_size %d exceeds the net's max_batch %lld",
                cfg->batch_size, online->max_batch);
    ZOO_REQUIRE(cfg->n_actions > 0 && cfg->update_horizon >= 1, "zoo_learner_create: bad n_actions / update_horizon");
    if (target) ZOO_REQUIRE(target->n_flat == online->n_flat && target->max_batch >= cfg->batch_size, "zoo_learner_create: target net does not match");
    const bool iqn = cfg->kind == ZOO_LEARNER_IQN;
    ZOO_REQUIRE(iqn == (online->desc.kind == ZOO_NET_IQN), "zoo_learner_create: IQNLearner <-> ImplicitQuantileNet mismatch");
    if (cfg->kind == ZOO_LEARNER_RAINBOW) {
        ZOO_REQUIRE(cfg->n_atoms >= 2 && cfg->support && online->n_out == cfg->n_atoms * cfg->n_actions,
                    "zoo_learner_create: Rainbow net must output n_atoms*n_actions = %d logits (has %d)", cfg->n_atoms * cfg->n_actions, online->n_out);
    } else if (cfg->kind == ZOO_LEARNER_QR_DQN || cfg->kind == ZOO_LEARNER_REM_DQN) {
        ZOO_REQUIRE(cfg->N >= 1 && online->n_out == cfg->N * cfg->n_actions,
                    "zoo_learner_create: QR-DQN / REM-DQN net must output N*n_actions = %d values (has %d)", cfg->N * cfg->n_actions, online->n_out);
    } else {
        ZOO_REQUIRE(online->n_out == cfg->n_actions, "zoo_learner_create: net outputs %d values, n_actions is %d", online->n_out, cfg->n_actions);
    }
    if (iqn) ZOO_REQUIRE(cfg->N > 0 && cfg->N_prime > 0 && cfg->N <= online->max_q && cfg->N_prime <= target->max_q && cfg->N_em == online->b.in_feat,
                         "zoo_learner_create: IQN N/N'/N_em do not fit the networks");
    zoo_learner *L = new zoo_learner();
    L->cfg = *cfg; L->cfg.support = nullptr;
    L->online = online; L->target = target; L->opt = opt; L->dp = dp; L->B = cfg->batch_size;
    L->gamma_n = (float)std::pow((double)cfg->gamma, (double)cfg->update_horizon);  // gamma^n, gamma::Float32 (dqn.jl:133)
    {
        // the learner's stream carries the update's critical path (online forward -> head -> data-gradient chain); the nets' side /
        // comm lanes (target forward, weight gradients, late bucket) fill the SMs it leaves free.  ZOO_LANE_PRIO=0: all lanes equal.
        static const bool prio = [] { const char *e = getenv("ZOO_LANE_PRIO"); return !(e && e[0] == '0'); }();
        int lo = 0, hi = 0;
        ZOO_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        ZOO_CUDA(cudaStreamCreateWithPriority(&L->st, cudaStreamNonBlocking, prio ? hi : 0));
    }
    const int B = L->B, A = cfg->n_actions;
    if (dp && dp_world(dp) > 1 && !online->grads_nccl) {
        // data parallel: move the flat gradient buffer into NCCL-allocated, communicator-registered memory (zero-copy all-reduce)
        void *p = nullptr, *h = nullptr;
        const size_t bytes = (online->n_flat + 64) * sizeof(float);
        ZOO_TRY(dp_alloc_registered(dp, bytes, &p, &h));
        if (p) {
            ZOO_CUDA(cudaDeviceSynchronize());
            ZOO_CUDA(cudaMemset(p, 0, bytes));
            ZOO_CUDA(cudaFree(online->grads));
            online->grads = (float *)p; online->grads_nccl = true; online->grads_dirty = false;
            L->nccl_reg = h;
        }
    }
    L->d_loss = online->grads + online->n_flat;  // tail of the gradient buffer (see net_alloc_backward)
    // late bucket = [offset of the largest weight tensor, end): valid when backward visits tensors in descending offset
    // order (Chain, ImplicitQuantileNet); a DuelingNetwork's val / adv heads break that order, so it keeps one bucket
    online->late_off = -1;
    const char *lb = getenv("ZOO_LATE_BUCKET");
    if (online->desc.kind != ZOO_NET_DUELING && !(lb && lb[0] == '0')) {
        long long best = 0;
        for (auto &t : online->tensors)
            if (t.kind == 0 && t.count > best) { best = t.count; online->late_off = t.offset; }
        if (online->late_off <= 0 || best * 2 < online->n_params) online->late_off = -1;  // not worth a second bucket
    }
    ZOO_CUDA(cudaMalloc(&L->d_obs, (size_t)2 * B * obs_elems(online) * sizeof(float)));
    {
        auto up = [](size_t x) { return (x + 255) / 256 * 256; };
        size_t o = 0;
        L->off_action = o; o = up(o + (size_t)B * sizeof(int));
        L->off_reward = o; o = up(o + (size_t)B * sizeof(float));
        L->off_terminal = o; o = up(o + (size_t)B);
        L->off_priority = o; o = up(o + (size_t)B * sizeof(float));
        L->off_mask = o; o = up(o + (size_t)B * A);
        const bool rem = cfg->kind == ZOO_LEARNER_REM_DQN;
        if (iqn || rem) { L->off_tau = o; o = up(o + (size_t)(rem ? 1 : B) * cfg->N * sizeof(float)); }
        if (iqn) { L->off_taup = o; o = up(o + (size_t)B * cfg->N_prime * sizeof(float)); }
        L->pack_bytes = o;
        ZOO_CUDA(cudaMalloc(&L->d_pack, o));
        ZOO_CUDA(cudaHostAlloc(&L->h_pack, o, cudaHostAllocDefault));
        ZOO_CUDA(cudaEventCreateWithFlags(&L->ev_pack, cudaEventDisableTiming));
        L->d_action = (int *)(L->d_pack + L->off_action);
        L->d_reward = (float *)(L->d_pack + L->off_reward);
        L->d_terminal = L->d_pack + L->off_terminal;
        L->d_priority = (float *)(L->d_pack + L->off_priority);
        L->d_mask = L->d_pack + L->off_mask;
        if (iqn || rem) L->d_tau = (float *)(L->d_pack + L->off_tau);
        if (iqn) L->d_taup = (float *)(L->d_pack + L->off_taup);
    }
    ZOO_CUDA(cudaMalloc(&L->d_per, B * sizeof(float)));
    ZOO_CUDA(cudaMalloc(&L->d_prio, B * sizeof(float)));
    ZOO_CUDA(cudaMalloc(&L->d_wraw, B * sizeof(float)));
    ZOO_CUDA(cudaMalloc(&L->d_wmax, sizeof(float)));
    ZOO_CUDA(cudaMalloc(&L->d_tail_part, 65 * sizeof(float)));
    ZOO_CUDA(cudaMemset(L->d_tail_part, 0, 65 * sizeof(float)));
    if (cfg->kind == ZOO_LEARNER_RAINBOW) {
        ZOO_CUDA(cudaMalloc(&L->d_support, cfg->n_atoms * sizeof(float)));
        ZOO_CUDA(cudaMemcpy(L->d_support, cfg->support, cfg->n_atoms * sizeof(float), cudaMemcpyHostToDevice));
        ZOO_CUDA(cudaMalloc(&L->d_tdist, (size_t)B * cfg->n_atoms * sizeof(float)));
    }
    if (cfg->kind == ZOO_LEARNER_QR_DQN) ZOO_CUDA(cudaMalloc(&L->d_target, (size_t)B * cfg->N * sizeof(float)));
    if (iqn) {
        ZOO_CUDA(cudaMalloc(&L->d_target, (size_t)B * cfg->N_prime * sizeof(float)));
    }
    *out = L;
    return ZOO_OK;
}

zoo_status zoo_learner_destroy(zoo_learner *L) {
    if (!L) return ZOO_OK;
    cudaStreamSynchronize(L->st);
    if (L->gexec) cudaGraphExecDestroy(L->gexec);
    if (L->nccl_reg) dp_deregister(L->dp, L->nccl_reg);
    if (L->h_pack) cudaFreeHost(L->h_pack);
    if (L->ev_pack) cudaEventDestroy(L->ev_pack);
    if (L->st_copy) {
        cudaStreamSynchronize(L->st_copy);
        for (auto &sl : L->slot) {
            if (sl.d_obs) cudaFree(sl.d_obs);
            if (sl.d_pack) cudaFree(sl.d_pack);
            if (sl.h_pack) cudaFreeHost(sl.h_pack);
            if (sl.ev_copied) cudaEventDestroy(sl.ev_copied);
            if (sl.ev_consumed) cudaEventDestroy(sl.ev_consumed);
        }
        for (int i = 0; i < zoo_learner::PIPE; ++i) {
            if (L->h_res[i]) cudaFreeHost(L->h_res[i]);
            if (L->ev_res[i]) cudaEventDestroy(L->ev_res[i]);
        }
        cudaStreamDestroy(L->st_copy);
    }
    void *ptrs[] = {L->d_obs, L->d_pack, L->d_support,
                    L->d_target, L->d_tdist, L->d_per, L->d_prio, L->d_wraw, L->d_wmax, L->d_tail_part};
    for (void *p : ptrs) if (p) cudaFree(p);
    cudaStreamDestroy(L->st);
    delete L;
    return ZOO_OK;
}

zoo_status zoo_learner_update(zoo_learner *L, const zoo_batch *b, float *loss_out, float *prio_out) { return run_update(L, b, loss_out, prio_out, true); }

// update! without waiting for its result: the loss (and the new priorities) of this update are fetched by a later
// zoo_learner_collect, in submission order.  Host batches are copied on a second stream into one of two staging slots, so the
// H2D copy of update i + 1 overlaps the kernels of update i; the caller's arrays must stay valid until the matching collect.
zoo_status zoo_learner_submit(zoo_learner *L, const zoo_batch *b, int32_t want_priorities) {
    ZOO_REQUIRE(L && b, "zoo_learner_submit: null argument");
    ZOO_REQUIRE(L->res_count < zoo_learner::PIPE, "zoo_learner_submit: %d updates are already in flight, collect one first", zoo_learner::PIPE);
    ZOO_TRY(run_update(L, b, nullptr, nullptr, true, true));
    ZOO_TRY(pipe_alloc(L));
    const int r = (L->res_head + L->res_count) % zoo_learner::PIPE;
    ZOO_CUDA(cudaMemcpyAsync(L->h_res[r], L->d_loss, sizeof(float), cudaMemcpyDeviceToHost, L->st));
    if (want_priorities) ZOO_CUDA(cudaMemcpyAsync(L->h_res[r] + 1, L->d_prio, L->B * sizeof(float), cudaMemcpyDeviceToHost, L->st));
    ZOO_CUDA(cudaEventRecord(L->ev_res[r], L->st));
    L->res_prio[r] = want_priorities != 0;
    L->res_count += 1;
    return ZOO_OK;
}

zoo_status zoo_learner_collect(zoo_learner *L, float *loss_out, float *prio_out) {
    ZOO_REQUIRE(L, "zoo_learner_collect: null learner");
    ZOO_REQUIRE(L->res_count > 0, "zoo_learner_collect: nothing was submitted");
    const int r = L->res_head;
    ZOO_REQUIRE(!prio_out || L->res_prio[r], "zoo_learner_collect: this update was submitted without want_priorities");
    ZOO_CUDA(cudaEventSynchronize(L->ev_res[r]));
    if (loss_out) *loss_out = L->h_res[r][0];
    if (prio_out) memcpy(prio_out, L->h_res[r] + 1, L->B * sizeof(float));
    L->res_head = (r + 1) % zoo_learner::PIPE;
    L->res_count -= 1;
    return ZOO_OK;
}
zoo_status zoo_learner_compute_grads(zoo_learner *L, const zoo_batch *b, float *loss_out, float *prio_out) { return run_update(L, b, loss_out, prio_out, false); }

zoo_status zoo_learner_apply(zoo_learner *L) {
    ZOO_REQUIRE(L, "zoo_learner_apply: null learner");
    g_launches = 0;
    ZOO_TRY(opt_step(L->opt, L->st));
    ZOO_TRY(net_mark_busy(L->online, L->st));
    ZOO_CUDA(cudaStreamSynchronize(L->st));
    L->online->grads_dirty = false;  // the optimiser kernels re-zero the gradients
    return ZOO_OK;
}

zoo_status zoo_learner_get_grad(const zoo_learner *L, int32_t id, float *host, int64_t n) {
    ZOO_REQUIRE(L, "zoo_learner_get_grad: null learner");
    ZOO_CUDA(cudaStreamSynchronize(L->st));
    return zoo_internal_tensor_io(L->online, L->online->grads, id, host, n, 0);
}

zoo_status zoo_learner_grad_buffer(zoo_learner *L, void **dev_ptr, int64_t *n) {
    ZOO_REQUIRE(L && dev_ptr && n, "zoo_learner_grad_buffer: null argument");
    *dev_ptr = L->online->grads;
    *n = L->online->n_flat;
    return ZOO_OK;
}

zoo_status zoo_learner_get_per_sample_loss(const zoo_learner *L, float *host, int32_t n) {
    ZOO_REQUIRE(L && host && n == L->B, "zoo_learner_get_per_sample_loss: expected %d floats", L ? L->B : 0);
    ZOO_CUDA(cudaMemcpyAsync(host, L->d_per, n * sizeof(float), cudaMemcpyDeviceToHost, L->st));
    ZOO_CUDA(cudaStreamSynchronize(L->st));
    return ZOO_OK;
}

zoo_status zoo_learner_priorities_dev(zoo_learner *L, float **dev_ptr) {
    ZOO_REQUIRE(L && dev_ptr, "zoo_learner_priorities_dev: null argument");
    *dev_ptr = L->d_prio;
    return ZOO_OK;
}

zoo_status zoo_learner_stream(zoo_learner *L, void **stream) {
    ZOO_REQUIRE(L && stream, "zoo_learner_stream: null argument");
    *stream = (void *)L->st;
    return ZOO_OK;
}
zoo_status zoo_learner_sync(zoo_learner *L) {
    ZOO_REQUIRE(L, "zoo_learner_sync: null learner");
    ZOO_CUDA(cudaStreamSynchronize(L->st));
    return ZOO_OK;
}
zoo_status zoo_learner_last_launches(const zoo_learner *L, int32_t *n) {
    ZOO_REQUIRE(L && n, "zoo_learner_last_launches: null argument");
    *n = L->last_launches;
    return ZOO_OK;
}
zoo_status zoo_learner_use_graph(zoo_learner *L, int32_t enable) {
    ZOO_REQUIRE(L, "zoo_learner_use_graph: null learner");
    L->use_graph = enable;
    return ZOO_OK;
}

}  // extern "C"
