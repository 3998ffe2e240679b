// Loss heads of the value-based learners: TD target / Double-DQN argmax / Huber, C51 projection + cross-entropy,
// IQN target selection + pairwise quantile-Huber, PER importance weights and new priorities.
// All arithmetic that the reference evaluates as separate fp32 operations uses explicit round-to-nearest
// intrinsics so ptxas cannot contract it into FMAs.
#pragma once
#include "common.cuh"

namespace zoo {

struct HeadArgs {
    int B;            // samples on this rank
    int A;            // actions
    const int *action;
    const float *reward;
    const uint8_t *terminal;
    const uint8_t *mask;   // next_legal_actions_mask [B][A] or null
    const float *w_raw;    // PER: 1 / (p + 1e-10)^beta, or null
    const float *w_max;    // PER: max over the (global) batch
    float inv_B;           // 1 / (global batch size)
    float *per_sample;     // [B] out
};


#ifdef __CUDACC__
__device__ __forceinline__ float sample_scale(const HeadArgs &h, int b) {
    // dot(w, l) * 1//B with w ./= maximum(w)  (prioritized_dqn.jl:126,141) or mean (dqn.jl:137)
    if (h.w_raw) return __fmul_rn(__fdiv_rn(h.w_raw[b], h.w_max[0]), h.inv_B);
    return h.inv_B;
}
// Julia argmax/findmax order: NaN counts as the maximum, ties go to the lowest index (q8)
__device__ __forceinline__ bool better(float v, float best) { return v > best || (isnan(v) && !isnan(best)); }

// One sample of the DQN-family head (dqn.jl:115-142, prioritized_dqn.jl:129-146, basic_dqn.jl:78-87): TD target, Huber loss,
// d loss / d Q, on explicit rows: q_s = online Q(s_b), q_n = online Q(s'_b) (basic / double, else unused), q_tr = target Q(s'_b)
// (null for basic); dout_s / dout_n = the gradient rows of Q(s_b) / Q(s'_b) (dout_n: basic only).
__device__ __forceinline__ void dqn_head_rows(int kind, int double_dqn, const HeadArgs &h, int b, const float *q_s, const float *q_n, const float *q_tr,
                                              float gamma_n, float *dout_s, float *dout_n) {
    const int A = h.A;
    const uint8_t *mk = h.mask ? h.mask + (size_t)b * A : nullptr;
    float q2;
    int amax = 0;
    if (kind == ZOO_LEARNER_BASIC_DQN) {
        float best = q_n[0];  // Q(s') of the same network
        for (int a = 1; a < A; ++a) if (better(q_n[a], best)) { best = q_n[a]; amax = a; }
        q2 = best;
    } else {
        const float *qn = double_dqn ? q_n : q_tr;  // double: online Q(s') selects, target evaluates
        float best = qn[0] + ((mk && !mk[0]) ? -INFINITY : 0.f);
        for (int a = 1; a < A; ++a) {
            float v = qn[a] + ((mk && !mk[a]) ? -INFINITY : 0.f);
            if (better(v, best)) { best = v; amax = a; }
        }
        q2 = double_dqn ? q_tr[amax] : best;
    }
    // G = r .+ gamma^n .* (1 .- t) .* q'
    const float disc = __fmul_rn(gamma_n, h.terminal[b] ? 0.f : 1.f);
    const float G = __fadd_rn(h.reward[b], __fmul_rn(disc, q2));
    const int a_b = h.action[b];
    const float q = q_s[a_b];
    // huber_loss(G, q; delta = 1): strict < mask (q4)
    const float e = __fsub_rn(G, q), ae = fabsf(e);
    const bool quad = ae < 1.0f;
    h.per_sample[b] = quad ? __fmul_rn(__fmul_rn(ae, ae), 0.5f) : __fsub_rn(ae, 0.5f);
    const float dG = quad ? e : (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f));
    const float sc = sample_scale(h, b);
    for (int a = 0; a < A; ++a) dout_s[a] = a == a_b ? -dG * sc : 0.f;
    if (kind == ZOO_LEARNER_BASIC_DQN)  // gradient also flows through the target (q3)
        for (int a = 0; a < A; ++a) dout_n[a] = a == amax ? dG * disc * sc : 0.f;
}
// The same on whole matrices: q_on [rows][A] online output ([s; s'] for basic / double), q_t [B][A] target output on s' (null for
// basic), dout shaped like q_on.  Shared by dqn_head_kernel and the fused small-batch tails (narrow.cu).
__device__ __forceinline__ void dqn_head_sample(int kind, int double_dqn, const HeadArgs &h, int b, const float *q_on, const float *q_t, float gamma_n,
                                                float *dout) {
    const size_t A = h.A;
    dqn_head_rows(kind, double_dqn, h, b, q_on + (size_t)b * A, q_on + (size_t)(h.B + b) * A, q_t ? q_t + (size_t)b * A : nullptr, gamma_n,
                  dout + (size_t)b * A, dout + (size_t)(h.B + b) * A);
}
#endif

zoo_status launch_per_weights(const float *priority, float beta, int B, float *w_raw, float *w_max, cudaStream_t st);
// kind: ZOO_LEARNER_BASIC_DQN | DQN | PRIORITIZED_DQN.  q_on: [rows][A] online net output ([s; s'] for basic/double),
// q_t: [B][A] target net output on s' (null for basic).  dout: same shape as q_on.
zoo_status launch_dqn_head(int kind, int double_dqn, const HeadArgs &h, const float *q_on, const float *q_t, float gamma_n,
                           float *dout, cudaStream_t st);
zoo_status launch_c51_head(const HeadArgs &h, int n_atoms, const float *tlogits, const float *logits, const float *support,
                           float delta_z, float vmin, float vmax, float gamma_n, float *dout, float *target_dist, cudaStream_t st);
zoo_status launch_iqn_target(const HeadArgs &h, int Np, const float *zt, float gamma, float *target, cudaStream_t st);
zoo_status launch_iqn_loss(const HeadArgs &h, int N, int Np, const float *z, const float *target, const float *tau, float kappa,
                           float *dout, cudaStream_t st);
// QR-DQN (qr_dqn.jl:106-138): zt / z are [B][A][N]; y [B][N]
zoo_status launch_qr_target(const HeadArgs &h, int N, const float *zt, float gamma, float *y, cudaStream_t st);
zoo_status launch_qr_loss(const HeadArgs &h, int N, const float *z, const float *y, float kappa, float *dout, cudaStream_t st);
// REM-DQN (rem_dqn.jl:100-145): q_on / q_t are [B][E][A]; w [E] the convex combination
zoo_status launch_rem_head(const HeadArgs &h, int E, const float *q_on, const float *q_t, const float *w, float gamma_n, float *dout,
                           cudaStream_t st);
// loss = sum_b scale_b * per_b  (scale = w/B for PER, 1/B otherwise); priorities = (per + 1e-10)^beta
zoo_status launch_loss_reduce(const HeadArgs &h, float beta, float *loss, float *prio_out, cudaStream_t st);
zoo_status launch_tau(float *tau, long long n, unsigned long long seed, const unsigned long long *counter, int stream_id, cudaStream_t st);

}  // namespace zoo
