// Loss heads (see heads.cuh).  Reference lines restated per kernel.
#include <curand_kernel.h>

#include "heads.cuh"

namespace zoo {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide reductions (blockDim.x multiple of 32, <= 1024); result identical in every thread
__device__ float block_sum(float v, float *red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    float s = 0.f;
    for (int i = 0; i < nw; ++i) s += red[i];
    return s;
}
__device__ float block_max(float v, float *red) {
    v = warp_max(v);
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    float s = red[0];
    for (int i = 1; i < nw; ++i) s = fmaxf(s, red[i]);
    return s;
}

// w = 1f0 ./ ((priority .+ 1f-10) .^ beta)  -- prioritized_dqn.jl:125, rainbow.jl:171, iqn.jl:197
__global__ void __launch_bounds__(1024) per_weights_kernel(const float *__restrict__ priority, float beta, int B,
                                                           float *__restrict__ w_raw, float *__restrict__ w_max) {
    __shared__ float red[32];
    float mx = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float w = __fdiv_rn(1.0f, powf(__fadd_rn(priority[b], 1e-10f), beta));
        w_raw[b] = w;
        mx = fmaxf(mx, w);
    }
    mx = block_max(mx, red);
    if (threadIdx.x == 0) w_max[0] = mx;
}

// dqn.jl:115-142, prioritized_dqn.jl:129-146, basic_dqn.jl:78-87
__global__ void dqn_head_kernel(int kind, int double_dqn, HeadArgs h, const float *__restrict__ q_on, const float *__restrict__ q_t,
                                float gamma_n, float *__restrict__ dout) {
    pdl_wait();
    pdl_trigger();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= h.B) return;
    dqn_head_sample(kind, double_dqn, h, b, q_on, q_t, gamma_n, dout);
}

// rainbow.jl:146-191 + project_distribution rainbow.jl:204-221 + logitcrossentropy (Dopamine_Rainbow_Atari.jl:54)
// one CTA of 64 threads per sample; thread j owns atom j
__global__ void __launch_bounds__(64) c51_head_kernel(HeadArgs h, int n, const float *__restrict__ tlogits, const float *__restrict__ logits,
                                                      const float *__restrict__ support, float delta_z, float vmin, float vmax,
                                                      float gamma_n, float *__restrict__ dout, float *__restrict__ target_dist) {
    __shared__ float red[2];
    __shared__ float p_sel[64], cl[64], dsl[64];
    const int b = blockIdx.x, j = threadIdx.x, A = h.A;
    const bool valid = j < n;
    const float zj = valid ? support[j] : 0.f;
    // next_probs = softmax over atoms; next_q = sum(support .* next_probs); mask; first-max action
    float best = 0.f, p_keep = 0.f;
    for (int a = 0; a < A; ++a) {
        const float x = valid ? tlogits[((size_t)b * A + a) * n + j] : -INFINITY;
        const float mx = block_max(x, red);
        const float ex = valid ? expf(x - mx) : 0.f;
        const float sm = block_sum(ex, red);
        const float p = ex / sm;
        float q = block_sum(valid ? __fmul_rn(zj, p) : 0.f, red);
        if (h.mask && !h.mask[(size_t)b * A + a]) q += -INFINITY;
        if (a == 0 || better(q, best)) { best = q; p_keep = p; }
    }
    // target_support = r .+ support * (gamma^n .* (1 .- t)); clamp to [Vmin, Vmax]
    const float disc = __fmul_rn(gamma_n, h.terminal[b] ? 0.f : 1.f);
    const float Tz = __fadd_rn(h.reward[b], __fmul_rn(zj, disc));
    p_sel[j] = valid ? p_keep : 0.f;
    cl[j] = fminf(fmaxf(Tz, vmin), vmax);
    __syncthreads();
    // out[j] = sum_i p[i] * clamp(1 - |clamp(Tz[i]) - z[j]| / dz, 0, 1)   (dense triangular kernel)
    float m = 0.f;
    if (valid) {
        for (int i = 0; i < n; ++i) {
            float k = __fsub_rn(1.0f, __fdiv_rn(fabsf(__fsub_rn(cl[i], zj)), delta_z));
            k = fminf(fmaxf(k, 0.f), 1.f);
            m = __fadd_rn(m, __fmul_rn(k, p_sel[i]));
        }
        if (target_dist) target_dist[(size_t)b * n + j] = m;
    }
    // batch_losses = -sum(target .* logsoftmax(select_logits)); d/dlogits = softmax * sum(target) - target
    const int a_b = h.action[b];
    const float x = valid ? logits[((size_t)b * A + a_b) * n + j] : -INFINITY;
    const float mx = block_max(x, red);
    const float zc = x - mx;
    const float lse = logf(block_sum(valid ? expf(zc) : 0.f, red));
    const float lsm = zc - lse;
    const float per = -block_sum(valid ? m * lsm : 0.f, red);
    const float msum = block_sum(m, red);
    const float sc = sample_scale(h, b);
    dsl[j] = valid ? (expf(lsm) * msum - m) * sc : 0.f;
    if (j == 0) h.per_sample[b] = per;
    __syncthreads();
    for (int idx = j; idx < A * n; idx += 64) {
        const int a = idx / n, jj = idx - a * n;
        dout[(size_t)b * A * n + idx] = a == a_b ? dsl[jj] : 0.f;
    }
}

// iqn.jl:175-189: mean over N' quantiles, (mask), first-max action, target = r + gamma * (1 - t) * z_t[a*]  (gamma^1: q1)
__global__ void iqn_target_kernel(HeadArgs h, int Np, const float *__restrict__ zt, float gamma, float *__restrict__ target) {
    __shared__ float red[32];
    const int b = blockIdx.x, i = threadIdx.x, A = h.A;
    const bool valid = i < Np;
    float best = 0.f;
    int sel = 0;
    for (int a = 0; a < A; ++a) {
        float s = block_sum(valid ? zt[((size_t)b * Np + i) * A + a] : 0.f, red);
        float avg = __fdiv_rn(s, (float)Np);
        if (h.mask && !h.mask[(size_t)b * A + a]) avg += -INFINITY;
        if (a == 0 || better(avg, best)) { best = avg; sel = a; }
    }
    if (valid) {
        const float disc = __fmul_rn(gamma, h.terminal[b] ? 0.f : 1.f);
        target[(size_t)b * Np + i] = __fadd_rn(h.reward[b], __fmul_rn(disc, zt[((size_t)b * Np + i) * A + sel]));
    }
}

// iqn.jl:203-222: TD[i,j] = target[i] - q[j]; raw = |tau[j] - 1{TD<0}| * huber_kappa(TD) / kappa;
// per-sample = mean_j sum_i raw (sum over N', mean over N: q2)
__global__ void iqn_loss_kernel(HeadArgs h, int N, int Np, const float *__restrict__ z, const float *__restrict__ target,
                                const float *__restrict__ tau, float kappa, float *__restrict__ dout) {
    extern __shared__ float tgt[];  // Np
    __shared__ float red[32];
    const int b = blockIdx.x, j = threadIdx.x, A = h.A;
    for (int i = j; i < Np; i += blockDim.x) tgt[i] = target[(size_t)b * Np + i];
    __syncthreads();
    const bool valid = j < N;
    const int a_b = h.action[b];
    float acc = 0.f, dacc = 0.f;
    if (valid) {
        const float q = z[((size_t)b * N + j) * A + a_b];
        const float tj = tau[(size_t)b * N + j];
        for (int i = 0; i < Np; ++i) {
            const float TD = __fsub_rn(tgt[i], q);
            const float ae = fabsf(TD);
            const float quad = fminf(ae, kappa);
            const float lin = __fsub_rn(ae, quad);
            const float hub = __fadd_rn(__fmul_rn(__fmul_rn(0.5f, quad), quad), __fmul_rn(kappa, lin));
            const float wgt = fabsf(__fsub_rn(tj, TD < 0.f ? 1.f : 0.f));
            acc = __fadd_rn(acc, __fdiv_rn(__fmul_rn(wgt, hub), kappa));
            dacc += wgt * fminf(fmaxf(TD, -kappa), kappa) / kappa;
        }
    }
    const float tot = block_sum(acc, red);
    if (j == 0) h.per_sample[b] = __fdiv_rn(tot, (float)N);
    if (valid) {
        const float dq = -dacc / (float)N * sample_scale(h, b);
        for (int a = 0; a < A; ++a) dout[((size_t)b * N + j) * A + a] = a == a_b ? dq : 0.f;
    }
}

// qr_dqn.jl:118-122: net output (N, A, B) column-major == [B][A][N]; mean over the N quantiles, first-max action,
// y = r + gamma * (1 - t) * z_t[:, a*]   (gamma, not gamma^n -- as written in the reference)
__global__ void qr_target_kernel(HeadArgs h, int N, const float *__restrict__ zt, float gamma, float *__restrict__ y) {
    __shared__ float red[32];
    const int b = blockIdx.x, i = threadIdx.x, A = h.A;
    const bool valid = i < N;
    float best = 0.f;
    int sel = 0;
    for (int a = 0; a < A; ++a) {
        const float s = block_sum(valid ? zt[((size_t)b * A + a) * N + i] : 0.f, red);
        const float avg = __fdiv_rn(s, (float)N);
        if (a == 0 || better(avg, best)) { best = avg; sel = a; }
    }
    if (valid) {
        const float disc = __fmul_rn(gamma, h.terminal[b] ? 0.f : 1.f);
        y[(size_t)b * N + i] = __fadd_rn(h.reward[b], __fmul_rn(disc, zt[((size_t)b * A + sel) * N + i]));
    }
}

// quantile_huber_loss(y_hat, y) -- qr_dqn.jl:3-14.  D[i,j] = y[i] - y_hat[j]; the cumulative probabilities
// cum_prob[i] = (i + 0.5) / N broadcast along the FIRST axis (the target index i, as written); loss = mean_{j,b} sum_i
__global__ void qr_loss_kernel(HeadArgs h, int N, const float *__restrict__ z, const float *__restrict__ y, float kappa,
                               float *__restrict__ dout) {
    extern __shared__ float tgt[];  // N
    __shared__ float red[32];
    const int b = blockIdx.x, j = threadIdx.x, A = h.A;
    for (int i = j; i < N; i += blockDim.x) tgt[i] = y[(size_t)b * N + i];
    __syncthreads();
    const bool valid = j < N;
    const int a_b = h.action[b];
    float acc = 0.f, dacc = 0.f;
    if (valid) {
        const float q = z[((size_t)b * A + a_b) * N + j];
        const float step = __fdiv_rn(1.0f, (float)N), first = __fdiv_rn(0.5f, (float)N);
        for (int i = 0; i < N; ++i) {
            const float D = __fsub_rn(tgt[i], q);
            const float ae = fabsf(D);
            const float quad = fminf(ae, kappa);
            const float lin = __fsub_rn(ae, quad);
            const float hub = __fadd_rn(__fmul_rn(__fmul_rn(0.5f, quad), quad), __fmul_rn(kappa, lin));
            const float cp = __fadd_rn(first, __fmul_rn((float)i, step));  // range(0.5f0 / N; length = N, step = 1f0 / N)
            const float wgt = fabsf(__fsub_rn(cp, D < 0.f ? 1.f : 0.f));
            acc = __fadd_rn(acc, __fmul_rn(wgt, hub));
            dacc += wgt * fminf(fmaxf(D, -kappa), kappa);
        }
    }
    const float tot = block_sum(acc, red);
    if (j == 0) h.per_sample[b] = __fdiv_rn(tot, (float)N);
    for (int idx = j; idx < A * N; idx += blockDim.x) dout[(size_t)b * A * N + idx] = 0.f;
    __syncthreads();
    if (valid) dout[((size_t)b * A + a_b) * N + j] = -dacc / (float)N * sample_scale(h, b);
}

// rem_dqn.jl:100-145: net output (A, E, B) column-major == [B][E][A]; q = sum_e w_e Q[:, e]; target from the target net
// (no double-DQN), legal-action mask, G = r + gamma^n (1 - t) max_a q_t; loss = huber(G, q[a]) (mean)
__global__ void rem_head_kernel(HeadArgs h, int E, const float *__restrict__ q_on, const float *__restrict__ q_t, const float *__restrict__ w,
                                float gamma_n, float *__restrict__ dout) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= h.B) return;
    const int A = h.A;
    float best = 0.f;
    for (int a = 0; a < A; ++a) {
        float s = 0.f;
        for (int e = 0; e < E; ++e) s = __fadd_rn(s, __fmul_rn(w[e], q_t[((size_t)b * E + e) * A + a]));
        if (h.mask && !h.mask[(size_t)b * A + a]) s += -INFINITY;
        if (a == 0 || better(s, best)) best = s;
    }
    const float G = __fadd_rn(h.reward[b], __fmul_rn(__fmul_rn(gamma_n, h.terminal[b] ? 0.f : 1.f), best));
    const int a_b = h.action[b];
    float q = 0.f;
    for (int e = 0; e < E; ++e) q = __fadd_rn(q, __fmul_rn(w[e], q_on[((size_t)b * E + e) * A + a_b]));
    const float err = __fsub_rn(G, q), ae = fabsf(err);
    h.per_sample[b] = ae < 1.f ? __fmul_rn(__fmul_rn(0.5f, ae), ae) : __fsub_rn(ae, 0.5f);  // Flux huber_loss, delta = 1, strict <
    const float dq = -fminf(fmaxf(err, -1.f), 1.f) * sample_scale(h, b);
    for (int e = 0; e < E; ++e)
        for (int a = 0; a < A; ++a) dout[((size_t)b * E + e) * A + a] = a == a_b ? w[e] * dq : 0.f;
}

__global__ void __launch_bounds__(1024) loss_reduce_kernel(HeadArgs h, float beta, float *__restrict__ loss, float *__restrict__ prio_out) {
    pdl_wait();
    pdl_trigger();
    __shared__ float red[32];
    float acc = 0.f;
    for (int b = threadIdx.x; b < h.B; b += blockDim.x) {
        const float l = h.per_sample[b];
        acc += l * sample_scale(h, b);
        // (batch_losses .+ 1f-10) .^ beta  -- prioritized_dqn.jl:143
        if (prio_out) prio_out[b] = powf(__fadd_rn(l, 1e-10f), beta);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) loss[0] = acc;
}

// tau ~ U[0,1) on the device (iqn.jl:173,191) when the batch carries none: Philox4x32-10, one counter tick per update
__global__ void tau_kernel(float *__restrict__ tau, long long n, unsigned long long seed, const unsigned long long *__restrict__ counter,
                           int stream_id) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    curandStatePhilox4_32_10_t s;
    curand_init(seed, (unsigned long long)(i / 4), (counter[0] * 2 + stream_id) * 4ull, &s);
    float4 r = curand_uniform4(&s);  // (0,1]
    float v[4] = {1.0f - r.x, 1.0f - r.y, 1.0f - r.z, 1.0f - r.w};  // -> [0,1)
    for (int k = 0; k < 4 && i + k < n; ++k) tau[i + k] = v[k];
}

// ---------------------------------------------------------------- launchers ---------------
zoo_status launch_per_weights(const float *priority, float beta, int B, float *w_raw, float *w_max, cudaStream_t st) {
    per_weights_kernel<<<1, 1024, 0, st>>>(priority, beta, B, w_raw, w_max);
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}
zoo_status launch_dqn_head(int kind, int double_dqn, const HeadArgs &h, const float *q_on, const float *q_t, float gamma_n,
                           float *dout, cudaStream_t st) {
    ZOO_CUDA(launch_pdl(dqn_head_kernel, dim3(cdiv(h.B, 128)), dim3(128), (size_t)0, st, kind, double_dqn, h, q_on, q_t, gamma_n, dout));
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}
zoo_status launch_c51_head(const HeadArgs &h, int n_atoms, const float *tlogits, const float *logits, const float *support,
                           float delta_z, float vmin, float vmax, float gamma_n, float *dout, float *target_dist, cudaStream_t st) {
    ZOO_REQUIRE(n_atoms >= 2 && n_atoms <= 64, "C51 head supports 2..64 atoms, got %d", n_atoms);
    c51_head_kernel<<<h.B, 64, 0, st>>>(h, n_atoms, tlogits, logits, support, delta_z, vmin, vmax, gamma_n, dout, target_dist);
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}
static int pow2_threads(int n) {
    int t = 32;
    while (t < n) t <<= 1;
    return t;
}
zoo_status launch_iqn_target(const HeadArgs &h, int Np, const float *zt, float gamma, float *target, cudaStream_t st) {
    ZOO_REQUIRE(Np >= 1 && Np <= 1024, "IQN: N' must be in 1..1024");
    iqn_target_kernel<<<h.B, pow2_threads(Np), 0, st>>>(h, Np, zt, gamma, target);
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}
zoo_status launch_iqn_loss(const HeadArgs &h, int N, int Np, const float *z, const float *target, const float *tau, float kappa,
                           float *dout, cudaStream_t st) {
    ZOO_REQUIRE(N >= 1 && N <= 1024 && Np >= 1 && Np <= 1024, "IQN: N and N' must be in 1..1024");
    iqn_loss_kernel<<<h.B, pow2_threads(N), Np * sizeof(float), st>>>(h, N, Np, z, target, tau, kappa, dout);
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}
zoo_status launch_qr_target(const HeadArgs &h, int N, const float *zt, float gamma, float *y, cudaStream_t st) {
    ZOO_REQUIRE(N <= 1024, "QR-DQN: n_quantile %d > 1024", N);
    qr_target_kernel<<<h.B, std::max(32, (N + 31) / 32 * 32), 0, st>>>(h, N, zt, gamma, y);
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}
zoo_status launch_qr_loss(const HeadArgs &h, int N, const float *z, const float *y, float kappa, float *dout, cudaStream_t st) {
    ZOO_REQUIRE(N <= 1024, "QR-DQN: n_quantile %d > 1024", N);
    qr_loss_kernel<<<h.B, std::max(32, (N + 31) / 32 * 32), N * sizeof(float), st>>>(h, N, z, y, kappa, dout);
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}
zoo_status launch_rem_head(const HeadArgs &h, int E, const float *q_on, const float *q_t, const float *w, float gamma_n, float *dout,
                           cudaStream_t st) {
    rem_head_kernel<<<cdiv(h.B, 128), 128, 0, st>>>(h, E, q_on, q_t, w, gamma_n, dout);
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

zoo_status launch_loss_reduce(const HeadArgs &h, float beta, float *loss, float *prio_out, cudaStream_t st) {
    ZOO_CUDA(launch_pdl(loss_reduce_kernel, dim3(1), dim3(1024), (size_t)0, st, h, beta, loss, prio_out));
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}
zoo_status launch_tau(float *tau, long long n, unsigned long long seed, const unsigned long long *counter, int stream_id, cudaStream_t st) {
    tau_kernel<<<cdiv(cdiv(n, 4), 256), 256, 0, st>>>(tau, n, seed, counter, stream_id);
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

}  // namespace zoo
