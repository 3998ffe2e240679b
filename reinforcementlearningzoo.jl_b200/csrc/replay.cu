// Device-resident replay shard: the batch assembly the reference does on the host inside `sample(rng, t, sampler)`
// (call site src/algorithms/dqns/common.jl:18; trajectory / sampler built at Dopamine_DQN_Atari.jl:55-66 --
// CircularArraySARTTrajectory + NStepBatchSampler{SARTS}(gamma, n, stack_size, batch_size); the implementation lives in
// ReinforcementLearningCore 0.7.2, not vendored: semantics restated in SURVEY.md Appendix A.2; the tests hold a CPU restatement).
//
//   ring of records [capacity][h*w + 16]: one uint8 frame per transition (the 4-frame stack is formed at sample time) followed
//   by its action (int32), reward (f32) and terminal (u8) -- one record = one push = one H2D copy; a prioritized shard adds the
//   SumTree of csrc/sumtree.cu
//   logical index i (1-based, oldest = 1) -> ring slot (first - 1 + i - 1) mod capacity      (same mapping as the SumTree)
//
//   fetch(i):  s  = frames[i-stack+1 .. i]  stacked as channels          s' = frames[i+n-stack+1 .. i+n]
//              a  = action[i];   m = first k in 1..n with terminal[i+k-1];   t = (m exists)
//              r  = sum_{k=1..(m or n)} gamma^(k-1) reward[i+k-1]   evaluated back to front in fp32: acc = r_k + gamma * acc
//   valid i:   stack <= i <= length - n
// Everything is enqueued on the stream the caller passes (the learner's), so sample -> update -> priority write-back is one
// ordered device-side sequence with no host round trip.
#include "sumtree.cuh"

struct zoo_sumtree;
namespace zoo {
zoo_status sumtree_sample_on(zoo_sumtree *t, cudaStream_t st, const void *u_dev, int u_dtype, long long n, int mode, long long *idx_dev, float *prio_dev);
zoo_status sumtree_update_on(zoo_sumtree *t, cudaStream_t st, const long long *idx_dev, const float *p_dev, long long n);
void sumtree_state(const zoo_sumtree *t, long long *length, long long *first);
void sumtree_view(const zoo_sumtree *t, SumTreeView *v);

struct ReplayView {
    const uint8_t *ring;
    long long capacity, first, length;
    int frame_bytes, rec_bytes, stack, n;
    float gamma;
};

__device__ __forceinline__ long long slot_of(const ReplayView &v, long long i) { return (v.first - 1 + i - 1) % v.capacity; }
__device__ __forceinline__ const uint8_t *rec_of(const ReplayView &v, long long i) { return v.ring + slot_of(v, i) * v.rec_bytes; }
__device__ __forceinline__ int rec_action(const ReplayView &v, long long i) { return *reinterpret_cast<const int *>(rec_of(v, i) + v.frame_bytes); }
__device__ __forceinline__ float rec_reward(const ReplayView &v, long long i) { return *reinterpret_cast<const float *>(rec_of(v, i) + v.frame_bytes + 4); }
__device__ __forceinline__ uint8_t rec_terminal(const ReplayView &v, long long i) { return rec_of(v, i)[v.frame_bytes + 8]; }

// uniform draws -> logical indices in the valid range.  PER indices come from the SumTree; one that falls outside
// stack .. length - n is REDRAWN, as RLCore's sampler does ("each redrawn while ind is not in the valid range"):
// lane b consumes its own spare uniforms spare[b * R + r], r = 0 .. R-1, each an independent single draw get(u * total) in the
// dtype of the draws, and takes the priority of the transition it finally fetches.  Only when all R spares are used up (or
// none were supplied) is the index clamped to the range edge and counted in n_invalid.
template <typename T>
__global__ void replay_indices_kernel(ReplayView v, const T *__restrict__ u, long long *__restrict__ idx, int B, int from_tree,
                                      unsigned *__restrict__ n_invalid, SumTreeView tv, const T *__restrict__ spare, int R,
                                      float *__restrict__ prio) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long lo = v.stack, hi = v.length - v.n;
    long long i;
    if (from_tree) {
        i = idx[b];
        for (int r = 0; (i < lo || i > hi) && r < R; ++r) {
            float p;
            i = sumtree_descend<T>(tv, Rn<T>::mul(spare[(long long)b * R + r], (T)__ldg(tv.tree + 1)), &p);
            prio[b] = p;
        }
        if (i < lo || i > hi) {
            atomicAdd(n_invalid, 1u);
            i = i < lo ? lo : hi;
            long long k = i + tv.first - 1;
            if (k > tv.capacity) k -= tv.capacity;
            prio[b] = tv.tree[tv.nparents + k];  // the importance weight must belong to the transition that is fetched
        }
    } else {
        i = lo + (long long)(u[b] * (T)(hi - lo + 1));
        if (i > hi) i = hi;
    }
    idx[b] = i;
}

// one CTA per (sample, stacked frame, state | next_state): copies one frame in 16-byte units
__global__ void replay_fetch_frames_kernel(ReplayView v, const long long *__restrict__ idx, uint4 *__restrict__ state, uint4 *__restrict__ next_state) {
    const int b = blockIdx.x, c = blockIdx.y, which = blockIdx.z;
    const long long i = idx[b] + (which ? v.n : 0) - v.stack + 1 + c;
    const int units = v.frame_bytes / 16;
    const uint4 *src = reinterpret_cast<const uint4 *>(rec_of(v, i));
    uint4 *dst = (which ? next_state : state) + ((long long)b * v.stack + c) * units;
    for (int k = threadIdx.x; k < units; k += blockDim.x) dst[k] = __ldg(src + k);
}

__global__ void replay_fetch_scalars_kernel(ReplayView v, const long long *__restrict__ idx, int B, int *__restrict__ action,
                                            float *__restrict__ reward, uint8_t *__restrict__ terminal) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long i = idx[b];
    action[b] = rec_action(v, i);
    int m = v.n;
    uint8_t t = 0;
    for (int k = 1; k <= v.n; ++k)
        if (rec_terminal(v, i + k - 1)) { m = k; t = 1; break; }
    float acc = 0.f;
    for (int k = m; k >= 1; --k) acc = __fadd_rn(rec_reward(v, i + k - 1), __fmul_rn(v.gamma, acc));
    reward[b] = acc;
    terminal[b] = t;
}
}  // namespace zoo

using namespace zoo;

extern "C" {
zoo_status zoo_sumtree_create(int64_t capacity, zoo_sumtree **out);
zoo_status zoo_sumtree_destroy(zoo_sumtree *t);
zoo_status zoo_sumtree_push(zoo_sumtree *t, float p);
zoo_status zoo_sumtree_push_many(zoo_sumtree *t, const float *p, int64_t n);
}

struct zoo_replay {
    long long capacity = 0, first = 1, length = 0;
    int h = 0, w = 0, stack = 1, n = 1, max_batch = 0;
    float gamma = 0.99f;
    uint8_t *ring = nullptr;      // [capacity][rec_bytes]
    size_t rec_bytes = 0;
    cudaEvent_t ev_push = nullptr;  // last push has landed
    // the ring is read (frame gather) on the caller's stream: ev_use marks the last such read, and the next push -- which may
    // overwrite the oldest slot -- waits for it on the push stream
    cudaEvent_t ev_use = nullptr;
    bool use_pending = false;
    zoo_sumtree *tree = nullptr;  // prioritized only
    // last sampled batch (device)
    uint8_t *b_state = nullptr, *b_next = nullptr, *b_terminal = nullptr;
    int *b_action = nullptr;
    float *b_reward = nullptr, *b_priority = nullptr;
    long long *b_idx = nullptr;
    void *b_u = nullptr;
    void *b_spare = nullptr;  // spare uniforms for PER redraws: [max_batch][spare_R] in the dtype of the draws
    int spare_R = 0, spare_dtype = -1;
    unsigned *n_invalid = nullptr;
    int last_B = 0;
    uint8_t *h_frame = nullptr;  // pinned staging ring (8 slots) of pushed frames + scalars
    unsigned long long n_pushed = 0;
    cudaStream_t st = nullptr;   // pushes
};

static ReplayView view_of(const zoo_replay *r) {
    ReplayView v;
    v.ring = r->ring;
    v.capacity = r->capacity; v.first = r->first; v.length = r->length;
    v.frame_bytes = r->h * r->w; v.rec_bytes = (int)r->rec_bytes; v.stack = r->stack; v.n = r->n; v.gamma = r->gamma;
    return v;
}

extern "C" {

zoo_status zoo_replay_create(int64_t capacity, int32_t frame_h, int32_t frame_w, int32_t stack_size, int32_t n_step, float gamma,
                             int32_t max_batch, int32_t prioritized, zoo_replay **out) {
    ZOO_REQUIRE(out && capacity >= 2 && frame_h > 0 && frame_w > 0 && stack_size >= 1 && n_step >= 1 && max_batch >= 1, "zoo_replay_create: bad argument");
    ZOO_REQUIRE((frame_h * frame_w) % 16 == 0, "zoo_replay_create: frame size %d x %d is not a multiple of 16 bytes", frame_h, frame_w);
    ZOO_TRY(require_sm100());
    zoo_replay *r = new zoo_replay();
    r->capacity = capacity; r->h = frame_h; r->w = frame_w; r->stack = stack_size; r->n = n_step; r->gamma = gamma; r->max_batch = max_batch;
    const size_t fb = (size_t)frame_h * frame_w;
    ZOO_CUDA(cudaStreamCreateWithFlags(&r->st, cudaStreamNonBlocking));
    r->rec_bytes = fb + 16;
    ZOO_CUDA(cudaMalloc(&r->ring, (size_t)capacity * r->rec_bytes));
    ZOO_CUDA(cudaEventCreateWithFlags(&r->ev_push, cudaEventDisableTiming));
    ZOO_CUDA(cudaEventCreateWithFlags(&r->ev_use, cudaEventDisableTiming));
    ZOO_CUDA(cudaMalloc(&r->b_state, (size_t)max_batch * stack_size * fb));
    ZOO_CUDA(cudaMalloc(&r->b_next, (size_t)max_batch * stack_size * fb));
    ZOO_CUDA(cudaMalloc(&r->b_action, max_batch * sizeof(int)));
    ZOO_CUDA(cudaMalloc(&r->b_reward, max_batch * sizeof(float)));
    ZOO_CUDA(cudaMalloc(&r->b_terminal, max_batch));
    ZOO_CUDA(cudaMalloc(&r->b_priority, max_batch * sizeof(float)));
    ZOO_CUDA(cudaMalloc(&r->b_idx, max_batch * sizeof(long long)));
    ZOO_CUDA(cudaMalloc(&r->b_u, max_batch * sizeof(double)));
    ZOO_CUDA(cudaMalloc(&r->n_invalid, sizeof(unsigned)));
    ZOO_CUDA(cudaMemset(r->n_invalid, 0, sizeof(unsigned)));
    ZOO_CUDA(cudaHostAlloc(&r->h_frame, 8 * r->rec_bytes, cudaHostAllocDefault));
    if (prioritized) ZOO_TRY(zoo_sumtree_create(capacity, &r->tree));
    *out = r;
    return ZOO_OK;
}

zoo_status zoo_replay_destroy(zoo_replay *r) {
    if (!r) return ZOO_OK;
    cudaDeviceSynchronize();
    if (r->ev_push) cudaEventDestroy(r->ev_push);
    if (r->ev_use) cudaEventDestroy(r->ev_use);
    void *ptrs[] = {r->ring, r->b_state, r->b_next, r->b_action, r->b_reward, r->b_terminal, r->b_priority,
                    r->b_idx, r->b_u, r->n_invalid, r->b_spare};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (r->h_frame) cudaFreeHost(r->h_frame);
    if (r->tree) zoo_sumtree_destroy(r->tree);
    cudaStreamDestroy(r->st);
    delete r;
    return ZOO_OK;
}

// push!(trajectory; state, action, reward, terminal) (+ push!(t[:priority], p), common.jl:28-37): one transition
zoo_status zoo_replay_push(zoo_replay *r, const uint8_t *frame_host, int32_t action, float reward, uint8_t terminal, float priority) {
    ZOO_REQUIRE(r && frame_host, "zoo_replay_push: null argument");
    if (r->use_pending) { ZOO_CUDA(cudaStreamWaitEvent(r->st, r->ev_use, 0)); r->use_pending = false; }
    if (r->length == r->capacity) r->first = (r->first == r->capacity) ? 1 : r->first + 1;
    else r->length += 1;
    const long long slot = (r->first - 1 + r->length - 1) % r->capacity;
    const size_t fb = (size_t)r->h * r->w;
    // the record is assembled in a pinned staging ring (8 slots, drained once per lap) and crosses PCIe in one asynchronous copy
    const int k = (int)(r->n_pushed++ & 7);
    if (k == 0) ZOO_CUDA(cudaStreamSynchronize(r->st));
    uint8_t *hs = r->h_frame + k * r->rec_bytes;
    memcpy(hs, frame_host, fb);
    memcpy(hs + fb, &action, 4);
    memcpy(hs + fb + 4, &reward, 4);
    hs[fb + 8] = terminal;
    ZOO_CUDA(cudaMemcpyAsync(r->ring + slot * r->rec_bytes, hs, r->rec_bytes, cudaMemcpyHostToDevice, r->st));
    ZOO_CUDA(cudaEventRecord(r->ev_push, r->st));
    if (r->tree) ZOO_TRY(zoo_sumtree_push(r->tree, priority));
    return ZOO_OK;
}

// bulk form of zoo_replay_push (filling a shard, tests): n transitions, frames [n][h][w]; priorities may be NULL (-> 100)
zoo_status zoo_replay_push_many(zoo_replay *r, const uint8_t *frames_host, const int32_t *action, const float *reward, const uint8_t *terminal,
                                const float *priority, int64_t n) {
    ZOO_REQUIRE(r && frames_host && action && reward && terminal && n >= 0, "zoo_replay_push_many: null argument");
    if (r->use_pending) { ZOO_CUDA(cudaStreamWaitEvent(r->st, r->ev_use, 0)); r->use_pending = false; }
    const size_t fb = (size_t)r->h * r->w;
    const int64_t CH = 1024;
    uint8_t *stage = nullptr;
    ZOO_CUDA(cudaHostAlloc(&stage, CH * r->rec_bytes, cudaHostAllocDefault));
    int64_t done = 0;
    while (done < n) {
        const int64_t m = std::min<int64_t>(CH, n - done);
        for (int64_t j = 0; j < m; ++j) {
            if (r->length == r->capacity) r->first = (r->first == r->capacity) ? 1 : r->first + 1;
            else r->length += 1;
            const long long slot = (r->first - 1 + r->length - 1) % r->capacity;
            uint8_t *hs = stage + j * r->rec_bytes;
            memcpy(hs, frames_host + (done + j) * fb, fb);
            memcpy(hs + fb, action + done + j, 4);
            memcpy(hs + fb + 4, reward + done + j, 4);
            hs[fb + 8] = terminal[done + j];
            ZOO_CUDA(cudaMemcpyAsync(r->ring + slot * r->rec_bytes, hs, r->rec_bytes, cudaMemcpyHostToDevice, r->st));
        }
        ZOO_CUDA(cudaStreamSynchronize(r->st));  // the staging buffer is reused by the next chunk
        done += m;
    }
    cudaFreeHost(stage);
    ZOO_CUDA(cudaEventRecord(r->ev_push, r->st));
    if (r->tree) {
        std::vector<float> p((size_t)n, 100.0f);
        if (priority) p.assign(priority, priority + n);
        ZOO_TRY(zoo_sumtree_push_many(r->tree, p.data(), n));
    }
    return ZOO_OK;
}

zoo_status zoo_replay_length(const zoo_replay *r, int64_t *length, int64_t *first) {
    ZOO_REQUIRE(r, "zoo_replay_length: null replay");
    if (length) *length = r->length;
    if (first) *first = r->first;
    return ZOO_OK;
}

// sample(rng, t, sampler) -- common.jl:18 -- with the B uniform draws supplied by the caller (host pointer, f32 | f64).
// mode: ZOO_REPLAY_UNIFORM (i = lo + floor(u * n_valid)), or the SumTree modes ZOO_SAMPLE_INDEPENDENT / ZOO_SAMPLE_STRATIFIED
// (prioritized shard).  Fills `batch` with DEVICE pointers (on_device = 1) valid until the next sample; everything is enqueued
// on `stream` (pass the learner's stream).  inds_out (host, optional) receives the 1-based logical indices and synchronises.
zoo_status zoo_replay_sample(zoo_replay *r, void *stream, const void *u_host, int32_t u_dtype, int32_t B, int32_t mode, zoo_batch *batch,
                             int64_t *inds_out) {
    ZOO_REQUIRE(r && u_host && batch && B >= 1 && B <= r->max_batch, "zoo_replay_sample: bad argument (B %d, max_batch %d)", B, r ? r->max_batch : 0);
    ZOO_REQUIRE(u_dtype == ZOO_DRAW_F32 || u_dtype == ZOO_DRAW_F64, "zoo_replay_sample: bad draw dtype");
    ZOO_REQUIRE(r->length - r->n >= r->stack, "zoo_replay_sample: %lld transitions are not enough (stack %d, n %d)", r->length, r->stack, r->n);
    const bool per = mode != ZOO_REPLAY_UNIFORM;
    ZOO_REQUIRE(!per || r->tree, "zoo_replay_sample: prioritized sampling from a shard created without priorities");
    cudaStream_t st = (cudaStream_t)stream;
    ZOO_CUDA(cudaStreamWaitEvent(st, r->ev_push, 0));  // pushes have landed (device-side dependency, no host block)
    const size_t esz = u_dtype == ZOO_DRAW_F64 ? 8 : 4;
    ZOO_CUDA(cudaMemcpyAsync(r->b_u, u_host, B * esz, cudaMemcpyHostToDevice, st));
    ReplayView v = view_of(r);
    if (per) {
        long long tl, tf;
        sumtree_state(r->tree, &tl, &tf);
        ZOO_REQUIRE(tl == r->length && tf == r->first, "zoo_replay_sample: the priority tree is out of step with the trajectory");
        ZOO_TRY(sumtree_sample_on(r->tree, st, r->b_u, u_dtype, B, mode, r->b_idx, r->b_priority));
    }
    SumTreeView tv{};
    if (per) sumtree_view(r->tree, &tv);
    const int R = (per && r->b_spare && r->spare_dtype == u_dtype) ? r->spare_R : 0;
    if (u_dtype == ZOO_DRAW_F64)
        replay_indices_kernel<double><<<cdiv(B, 128), 128, 0, st>>>(v, (const double *)r->b_u, r->b_idx, B, per, r->n_invalid, tv, (const double *)r->b_spare, R, r->b_priority);
    else
        replay_indices_kernel<float><<<cdiv(B, 128), 128, 0, st>>>(v, (const float *)r->b_u, r->b_idx, B, per, r->n_invalid, tv, (const float *)r->b_spare, R, r->b_priority);
    ZOO_LAUNCH_CHECK();
    replay_fetch_frames_kernel<<<dim3(B, r->stack, 2), 256, 0, st>>>(v, r->b_idx, (uint4 *)r->b_state, (uint4 *)r->b_next);
    ZOO_LAUNCH_CHECK();
    replay_fetch_scalars_kernel<<<cdiv(B, 128), 128, 0, st>>>(v, r->b_idx, B, r->b_action, r->b_reward, r->b_terminal);
    ZOO_LAUNCH_CHECK();
    r->last_B = B;
    ZOO_CUDA(cudaEventRecord(r->ev_use, st));
    r->use_pending = true;
    memset(batch, 0, sizeof(*batch));
    batch->B = B; batch->obs_dtype = ZOO_OBS_U8; batch->state = r->b_state; batch->next_state = r->b_next; batch->action = r->b_action;
    batch->reward = r->b_reward; batch->terminal = r->b_terminal; batch->priority = per ? r->b_priority : nullptr; batch->on_device = 1;
    if (inds_out) {
        ZOO_CUDA(cudaMemcpyAsync(inds_out, r->b_idx, B * sizeof(long long), cudaMemcpyDeviceToHost, st));
        ZOO_CUDA(cudaStreamSynchronize(st));
    }
    return ZOO_OK;
}

// t[:priority][inds] .= priorities -- common.jl:22 -- for the indices of the last sample; prio_dev: device pointer (e.g. the
// learner's new-priority buffer, zoo_learner_priorities_dev), consumed on `stream`
zoo_status zoo_replay_update_priorities(zoo_replay *r, void *stream, const float *prio_dev) {
    ZOO_REQUIRE(r && r->tree && prio_dev && r->last_B > 0, "zoo_replay_update_priorities: nothing sampled / not prioritized");
    return sumtree_update_on(r->tree, (cudaStream_t)stream, r->b_idx, prio_dev, r->last_B);
}

// Spare uniforms for the redraws of out-of-range prioritized indices: spare_host [B][R] (same dtype as the draws of the next
// zoo_replay_sample calls; lane b uses row b).  Supplied by the host so that sampling stays a deterministic function of its
// inputs; consumed by every following sample until replaced (n_per_sample = 0 switches redraws off: clamp + count).
zoo_status zoo_replay_set_spare_draws(zoo_replay *r, void *stream, const void *spare_host, int32_t u_dtype, int32_t B, int32_t n_per_sample) {
    ZOO_REQUIRE(r && B >= 1 && B <= r->max_batch && n_per_sample >= 0 && n_per_sample <= 64, "zoo_replay_set_spare_draws: bad argument");
    ZOO_REQUIRE(u_dtype == ZOO_DRAW_F32 || u_dtype == ZOO_DRAW_F64, "zoo_replay_set_spare_draws: bad draw dtype");
    r->spare_R = 0;
    if (n_per_sample == 0) return ZOO_OK;
    ZOO_REQUIRE(spare_host, "zoo_replay_set_spare_draws: null draws");
    if (!r->b_spare) ZOO_CUDA(cudaMalloc(&r->b_spare, (size_t)r->max_batch * 64 * sizeof(double)));
    const size_t esz = u_dtype == ZOO_DRAW_F64 ? 8 : 4;
    ZOO_CUDA(cudaMemcpyAsync(r->b_spare, spare_host, (size_t)B * n_per_sample * esz, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    r->spare_R = n_per_sample; r->spare_dtype = u_dtype;
    return ZOO_OK;
}

zoo_status zoo_replay_invalid_count(zoo_replay *r, uint32_t *n) {
    ZOO_REQUIRE(r && n, "zoo_replay_invalid_count: null argument");
    ZOO_CUDA(cudaMemcpy(n, r->n_invalid, sizeof(unsigned), cudaMemcpyDeviceToHost));
    return ZOO_OK;
}

zoo_status zoo_replay_sumtree(zoo_replay *r, zoo_sumtree **tree) {
    ZOO_REQUIRE(r && tree, "zoo_replay_sumtree: null argument");
    *tree = r->tree;
    return ZOO_OK;
}

}  // extern "C"
