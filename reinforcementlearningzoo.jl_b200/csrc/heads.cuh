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
