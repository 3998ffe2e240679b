// Narrow output layers (the Q head: Dense(512 => n_actions), n_actions <= 8) as HBM-streaming CUDA-core kernels.
// A 6-column GEMM has nothing for the tensor cores; what matters is reading the [rows][K] activation matrix once, with
// 16-byte loads, and folding everything that hangs off it into the same pass:
//   narrow_fwd_kernel : out[m][n] = b[n] + sum_k X[m][k] W[n][k]                               (one warp per row)
//   narrow_bwd_kernel : dX = (dY W) .* relu'(X),  dW += dY^T X,  db += colsum(dY),  db_prev += colsum(dX)
//                       -- the layer's data gradient, weight gradient, bias gradient AND the previous layer's bias
//                       gradient in ONE pass over X (they were four launches: colsum, wgrad GEMM, dgrad GEMM, colsum)
// Reference layers: Dense(512, n_actions) at src/experiments/atari/Dopamine_DQN_Atari.jl:34 and the IQN header
// Dopamine_IQN_Atari.jl:40-43 (rows = batch x quantiles, 32 768 at the bench size); differentiated at dqn.jl:135, iqn.jl:203.
#include "heads.cuh"

namespace zoo {

static constexpr int NARROW_NMAX = 8;

template <typename T> struct NarrowVec;
template <> struct NarrowVec<float> {
    static constexpr int V = 4;
    typedef float4 raw;
    static __device__ __forceinline__ void unpack(const raw &r, float (&f)[4]) { f[0] = r.x; f[1] = r.y; f[2] = r.z; f[3] = r.w; }
    static __device__ __forceinline__ raw pack(const float (&f)[4]) { return make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct NarrowVec<bf16> {
    static constexpr int V = 8;
    typedef uint4 raw;
    static __device__ __forceinline__ void unpack(const raw &r, float (&f)[8]) {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ raw pack(const float (&f)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t *>(&h);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// 16-byte red (four consecutive floats): the per-SM atomic issue rate bounds these flushes, so a quarter of the instructions
__device__ __forceinline__ void red_add_v4(float *p, const float *v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
__device__ __forceinline__ void flush_reds(float *dst, const float *src_s, int n, bool vec) {
    if (vec) {
        for (int i = threadIdx.x * 4; i < n; i += 1024) red_add_v4(dst + i, src_s + i);
    } else {
        for (int i = threadIdx.x; i < n; i += 256) atomicAdd(dst + i, src_s[i]);
    }
}

// one warp per row (grid-stride); lane owns the 16-byte chunks lane, lane + 32, ... of the row; W lives in registers
template <typename T, int CPL>
__global__ void __launch_bounds__(256) narrow_fwd_kernel(const T *__restrict__ X, const float *__restrict__ W, const float *__restrict__ bias,
                                                         float *__restrict__ out, int M, int N, int K) {
    using NV = NarrowVec<T>;
    constexpr int V = NV::V;
    extern __shared__ float w_s[];  // W [N][K], staged once per CTA with coalesced 16-byte loads
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int nchunks = K / V;
    for (int i = threadIdx.x; i < N * K / 4; i += 256) reinterpret_cast<float4 *>(w_s)[i] = __ldg(reinterpret_cast<const float4 *>(W) + i);
    __syncthreads();
    // weights as packed pairs: the dot products run on FFMA2 (two k per instruction) -- this kernel was issue-bound, not
    // memory-bound (ncu: 42 % of the issue slots at 12 % occupancy, 1.3 TB/s)
    float2 w[NARROW_NMAX][CPL * V / 2];
#pragma unroll
    for (int n = 0; n < NARROW_NMAX; ++n)
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int ch = lane + 32 * c;
#pragma unroll
            for (int e4 = 0; e4 < V; e4 += 4) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < N && ch < nchunks) t = *reinterpret_cast<const float4 *>(w_s + n * K + ch * V + e4);
                w[n][(c * V + e4) / 2] = make_float2(t.x, t.y); w[n][(c * V + e4) / 2 + 1] = make_float2(t.z, t.w);
            }
        }
    // after the butterfly below, lane l holds the row's output n = 4 b4 + 2 b3 + b2 (bits of l); lanes with l % 4 == 0 store it
    const int n_mine = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const float bn = n_mine < N ? __ldg(bias + n_mine) : 0.f;
    pdl_trigger();
    // ROWS rows in flight per warp: a streaming kernel with 8 resident warps per SM needs several independent 16-byte loads per
    // lane to cover the DRAM latency (one row at a time ran at 0.87 TB/s, profiles/r02_ncu_full_iqn_b512_bf16.txt)
    constexpr int ROWS = CPL >= 4 ? 1 : 4;  // (CPL = 4: the weights alone fill the register file)
    for (int mb = (blockIdx.x * 8 + (threadIdx.x >> 5)) * ROWS; mb < M; mb += gridDim.x * 8 * ROWS) {
        typename NV::raw r[ROWS][CPL];
#pragma unroll
        for (int u = 0; u < ROWS; ++u)
#pragma unroll
            for (int c = 0; c < CPL; ++c)
                if (mb + u < M && lane + 32 * c < nchunks) r[u][c] = reinterpret_cast<const typename NV::raw *>(X + (long long)(mb + u) * K)[lane + 32 * c];
#pragma unroll
        for (int u = 0; u < ROWS; ++u) {
            const int m = mb + u;
            if (m >= M) break;
            float2 acc[NARROW_NMAX];
#pragma unroll
            for (int n = 0; n < NARROW_NMAX; ++n) acc[n] = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                if (lane + 32 * c < nchunks) {
                    float x[V];
                    NV::unpack(r[u][c], x);
#pragma unroll
                    for (int n = 0; n < NARROW_NMAX; ++n)
#pragma unroll
                        for (int e = 0; e < V; e += 2) acc[n] = __ffma2_rn(make_float2(x[e], x[e + 1]), w[n][(c * V + e) / 2], acc[n]);
                }
            }
            // eight sums over 32 lanes in 9 shuffles: each exchange halves the values a lane still carries (16: n < 4 stay in the lanes
            // with bit 4 clear, n >= 4 in the others; 8 and 4 likewise), the last two steps are a plain butterfly on the one left
            float t4[4], t2[2];
            {
                const bool hi = lane & 16;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float a0 = acc[i].x + acc[i].y, a1 = acc[i + 4].x + acc[i + 4].y;
                    t4[i] = (hi ? a1 : a0) + __shfl_xor_sync(0xffffffffu, hi ? a0 : a1, 16);
                }
            }
            {
                const bool hi = lane & 8;
#pragma unroll
                for (int i = 0; i < 2; ++i) t2[i] = (hi ? t4[i + 2] : t4[i]) + __shfl_xor_sync(0xffffffffu, hi ? t4[i] : t4[i + 2], 8);
            }
            const bool hi4 = lane & 4;
            float v = (hi4 ? t2[1] : t2[0]) + __shfl_xor_sync(0xffffffffu, hi4 ? t2[0] : t2[1], 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if ((lane & 3) == 0 && n_mine < N) out[(long long)m * N + n_mine] = v + bn;
        }
    }
}

// TPR = K / 8 threads share a row (8 consecutive k each), 256 / TPR rows per pass; grid-stride over row groups.
// Shared memory: dW accumulator [N][K], db_prev accumulator [K], db accumulator [N] -- flushed with one atomic per element per CTA.
template <typename T>
__global__ void __launch_bounds__(256) narrow_bwd_kernel(const T *__restrict__ X, const float *__restrict__ dY, const float *__restrict__ W,
                                                         T *__restrict__ dX, float *__restrict__ dW, float *__restrict__ db, float *__restrict__ db_prev,
                                                         int M, int N, int K, int relu_prev) {
    extern __shared__ float acc_s[];
    float *dWs = acc_s, *dbp_s = acc_s + NARROW_NMAX * K, *dbn_s = dbp_s + K;
    const int TPR = K >> 3, RPP = 256 / TPR;
    const int kq = threadIdx.x % TPR, rl = threadIdx.x / TPR;
    pdl_wait();
    // W is staged through the (not yet used) dW accumulator with coalesced 16-byte loads, then each thread keeps its 8 columns
    for (int i = threadIdx.x; i < N * K / 4; i += 256) reinterpret_cast<float4 *>(dWs)[i] = __ldg(reinterpret_cast<const float4 *>(W) + i);
    __syncthreads();
    float2 w[NARROW_NMAX][4], dw[NARROW_NMAX][4];  // packed pairs of this thread's 8 columns: the row loop runs on FFMA2
    float dbp[8], dbn[NARROW_NMAX];
#pragma unroll
    for (int n = 0; n < NARROW_NMAX; ++n) {
        dbn[n] = 0.f;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (n < N) {
            a = *reinterpret_cast<const float4 *>(dWs + n * K + kq * 8);
            b = *reinterpret_cast<const float4 *>(dWs + n * K + kq * 8 + 4);
        }
        w[n][0] = make_float2(a.x, a.y); w[n][1] = make_float2(a.z, a.w); w[n][2] = make_float2(b.x, b.y); w[n][3] = make_float2(b.z, b.w);
#pragma unroll
        for (int e = 0; e < 4; ++e) dw[n][e] = make_float2(0.f, 0.f);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NARROW_NMAX * K + K + NARROW_NMAX; i += 256) acc_s[i] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) dbp[e] = 0.f;
    pdl_trigger();
    // the next row's input and output gradient are fetched before this row is computed (one row per thread group at a time left
    // the kernel latency-bound: 0.52 TB/s at 32 768 rows)
    uint4 nx0 = make_uint4(0u, 0u, 0u, 0u), nx1 = nx0;
    float ndy[NARROW_NMAX];
    auto fetch = [&](long long m) {
        if (sizeof(T) == 2) nx0 = *reinterpret_cast<const uint4 *>(reinterpret_cast<const bf16 *>(X) + m * K + kq * 8);
        else {
            nx0 = *reinterpret_cast<const uint4 *>(reinterpret_cast<const float *>(X) + m * K + kq * 8);
            nx1 = *reinterpret_cast<const uint4 *>(reinterpret_cast<const float *>(X) + m * K + kq * 8 + 4);
        }
#pragma unroll
        for (int n = 0; n < NARROW_NMAX; ++n) ndy[n] = n < N ? __ldg(dY + m * N + n) : 0.f;
    };
    const long long mstep = (long long)gridDim.x * RPP;
    if ((long long)blockIdx.x * RPP + rl < M) fetch((long long)blockIdx.x * RPP + rl);
    for (long long m = (long long)blockIdx.x * RPP + rl; m < M; m += mstep) {
        float x[8], dy[NARROW_NMAX];
        if (sizeof(T) == 2) {
            NarrowVec<bf16>::unpack(nx0, x);
        } else {
            x[0] = __uint_as_float(nx0.x); x[1] = __uint_as_float(nx0.y); x[2] = __uint_as_float(nx0.z); x[3] = __uint_as_float(nx0.w);
            x[4] = __uint_as_float(nx1.x); x[5] = __uint_as_float(nx1.y); x[6] = __uint_as_float(nx1.z); x[7] = __uint_as_float(nx1.w);
        }
#pragma unroll
        for (int n = 0; n < NARROW_NMAX; ++n) dy[n] = ndy[n];
        if (m + mstep < M) fetch(m + mstep);
        float dx[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {  // same per-element order of operations as the scalar loop: bit-identical results
            float2 s = make_float2(0.f, 0.f);
#pragma unroll
            for (int n = 0; n < NARROW_NMAX; ++n) s = __ffma2_rn(make_float2(dy[n], dy[n]), w[n][e], s);
            if (relu_prev && !(x[2 * e] > 0.f)) s.x = 0.f;
            if (relu_prev && !(x[2 * e + 1] > 0.f)) s.y = 0.f;
            dx[2 * e] = s.x; dx[2 * e + 1] = s.y;
            dbp[2 * e] += s.x; dbp[2 * e + 1] += s.y;
#pragma unroll
            for (int n = 0; n < NARROW_NMAX; ++n) dw[n][e] = __ffma2_rn(make_float2(dy[n], dy[n]), make_float2(x[2 * e], x[2 * e + 1]), dw[n][e]);
        }
        if (sizeof(T) == 2) {
            *reinterpret_cast<uint4 *>(reinterpret_cast<bf16 *>(dX) + m * K + kq * 8) = NarrowVec<bf16>::pack(dx);
        } else {
            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(dX) + m * K + kq * 8) = make_float4(dx[0], dx[1], dx[2], dx[3]);
            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(dX) + m * K + kq * 8 + 4) = make_float4(dx[4], dx[5], dx[6], dx[7]);
        }
        if (kq == 0) {
#pragma unroll
            for (int n = 0; n < NARROW_NMAX; ++n) dbn[n] += dy[n];
        }
    }
    __syncthreads();  // accumulators zeroed
#pragma unroll
    for (int n = 0; n < NARROW_NMAX; ++n) {
        if (n < N) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(dWs + n * K + kq * 8 + e, (e & 1) ? dw[n][e >> 1].y : dw[n][e >> 1].x);
            if (kq == 0) atomicAdd(dbn_s + n, dbn[n]);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(dbp_s + kq * 8 + e, dbp[e]);
    __syncthreads();
    flush_reds(dW, dWs, N * K, ((uintptr_t)dW & 15) == 0);
    if (db_prev) flush_reds(db_prev, dbp_s, K, ((uintptr_t)db_prev & 15) == 0);
    if (threadIdx.x < N) atomicAdd(db + threadIdx.x, dbn_s[threadIdx.x]);
}

bool narrow_supported(int N, long long K, int x_bf16) {
    if (N < 1 || N > NARROW_NMAX || K < 8 || K % 8 || K > 2048) return false;
    const int tpr = (int)(K / 8);
    if (256 % tpr) return false;                       // backward: K/8 threads per row must divide the CTA
    const int v = x_bf16 ? 8 : 4;
    return K % v == 0 && K / v <= 32 * 4;               // forward: at most 4 chunks of 16 bytes per lane
}

zoo_status narrow_forward(const void *X, int x_bf16, const float *W, const float *bias, float *out, long long M, int N, long long K, cudaStream_t st) {
    ZOO_REQUIRE(narrow_supported(N, K, x_bf16) && M > 0 && M < (1ll << 31), "narrow_forward: unsupported shape (N=%d K=%lld)", N, K);
    const int v = x_bf16 ? 8 : 4;
    const int cpl = cdiv(K / v, 32);
    const dim3 grid((unsigned)std::min<long long>(cdiv(M, 8 * (cpl <= 2 ? 4 : 1)), 148 * 2)), block(256);
    const size_t smem = (size_t)N * K * sizeof(float);
    static bool attr = false;
    if (!attr) {
        const int mx = NARROW_NMAX * 2048 * (int)sizeof(float);
        ZOO_CUDA(cudaFuncSetAttribute(narrow_fwd_kernel<bf16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        ZOO_CUDA(cudaFuncSetAttribute(narrow_fwd_kernel<bf16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        ZOO_CUDA(cudaFuncSetAttribute(narrow_fwd_kernel<bf16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        ZOO_CUDA(cudaFuncSetAttribute(narrow_fwd_kernel<float, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        ZOO_CUDA(cudaFuncSetAttribute(narrow_fwd_kernel<float, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        ZOO_CUDA(cudaFuncSetAttribute(narrow_fwd_kernel<float, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        attr = true;
    }
#define NFWD(T, C) ZOO_CUDA(launch_pdl(narrow_fwd_kernel<T, C>, grid, block, smem, st, (const T *)X, W, bias, out, (int)M, N, (int)K))
    if (x_bf16) { if (cpl <= 1) NFWD(bf16, 1); else if (cpl == 2) NFWD(bf16, 2); else NFWD(bf16, 4); }
    else { if (cpl <= 1) NFWD(float, 1); else if (cpl == 2) NFWD(float, 2); else NFWD(float, 4); }
#undef NFWD
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

zoo_status narrow_backward(const void *X, int x_bf16, const float *dY, const float *W, void *dX, float *dW, float *db, float *db_prev, long long M,
                           int N, long long K, int relu_prev, cudaStream_t st) {
    ZOO_REQUIRE(narrow_supported(N, K, x_bf16) && M > 0, "narrow_backward: unsupported shape (N=%d K=%lld)", N, K);
    const int rpp = 256 / (int)(K / 8);
    const dim3 grid((unsigned)std::min<long long>(cdiv(M, rpp), 148 * 2)), block(256);
    const size_t smem = (size_t)(NARROW_NMAX * K + K + NARROW_NMAX) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        ZOO_CUDA(cudaFuncSetAttribute(narrow_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((NARROW_NMAX * 2048 + 2048 + NARROW_NMAX) * sizeof(float))));
        ZOO_CUDA(cudaFuncSetAttribute(narrow_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((NARROW_NMAX * 2048 + 2048 + NARROW_NMAX) * sizeof(float))));
        attr = true;
    }
    if (x_bf16)
        ZOO_CUDA(launch_pdl(narrow_bwd_kernel<bf16>, grid, block, smem, st, (const bf16 *)X, dY, W, (bf16 *)dX, dW, db, db_prev, (int)M, N, (int)K, relu_prev));
    else
        ZOO_CUDA(launch_pdl(narrow_bwd_kernel<float>, grid, block, smem, st, (const float *)X, dY, W, (float *)dX, dW, db, db_prev, (int)M, N, (int)K, relu_prev));
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

// ---------------------------------------------------------------- fused small-batch tail of the DQN-family update
// One CTA does everything between "fc1's activations are ready" and "fc1's output gradient is ready" for BasicDQN / DQN /
// PrioritizedDQN at Dopamine batch sizes (rows <= TAIL_ROWS):  Q = h W2^T + b2 for the online rows ([s; s'] when the learner
// needs Q(s') of the online net), the TD target / Huber head per sample (dqn.jl:115-142), loss = sum_b scale_b * l_b and the new
// priorities (prioritized_dqn.jl:141-143), then the head layer's whole backward as in narrow_bwd_kernel.  It replaces four
// dependent launches on the critical path (narrow_fwd, dqn_head, loss_reduce, narrow_bwd).
static constexpr int TAIL_ROWS = 128;
template <typename T>
__global__ void __launch_bounds__(256) dqn_tail_kernel(int kind, int double_dqn, HeadArgs h, const T *__restrict__ X, const float *__restrict__ W,
                                                       const float *__restrict__ bias, const float *__restrict__ q_t, float gamma_n, float beta,
                                                       int rows_fwd, int rows_bwd, int K, int relu_prev, float *__restrict__ q_out, float *__restrict__ dq_out,
                                                       T *__restrict__ dX, float *__restrict__ dW, float *__restrict__ db, float *__restrict__ db_prev,
                                                       float *__restrict__ loss, float *__restrict__ prio_out) {
    extern __shared__ float sm_[];
    const int N = h.A;
    float *Ws = sm_;                               // [N][K]  (later the dW accumulator)
    float *dbp_s = Ws + NARROW_NMAX * K;           // [K]
    float *q_s = dbp_s + K;                        // [TAIL_ROWS][N]
    float *dq_s = q_s + TAIL_ROWS * NARROW_NMAX;   // [TAIL_ROWS][N]
    float *red = dq_s + TAIL_ROWS * NARROW_NMAX;   // [32] + dbn [N]
    float *dbn_s = red + 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_wait();
    for (int i = threadIdx.x; i < N * K / 4; i += 256) reinterpret_cast<float4 *>(Ws)[i] = __ldg(reinterpret_cast<const float4 *>(W) + i);
    for (int i = threadIdx.x; i < K + 32 + NARROW_NMAX; i += 256) (i < K ? dbp_s[i] : red[i - K]) = 0.f;
    for (int i = threadIdx.x; i < TAIL_ROWS * NARROW_NMAX; i += 256) dq_s[i] = 0.f;
    __syncthreads();
    // ---- forward: one warp per row, two rows in flight per warp; lane owns the 16-byte chunks lane, lane + 32, ... (<= 4) of a row,
    // all loaded before the first FMA (a scalar k loop here serialised 16 dependent global loads per row: 29 us for 64 rows)
    {
        using NV = NarrowVec<T>;
        constexpr int V = NV::V;
        const int nchunks = K / V;
        for (int mb = warp; mb < rows_fwd; mb += 16) {
            typename NV::raw r[2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int m = mb + 8 * u;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (m < rows_fwd && lane + 32 * c < nchunks) r[u][c] = reinterpret_cast<const typename NV::raw *>(X + (long long)m * K)[lane + 32 * c];
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int m = mb + 8 * u;
                if (m >= rows_fwd) break;
                float acc[NARROW_NMAX];
#pragma unroll
                for (int n = 0; n < NARROW_NMAX; ++n) acc[n] = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (lane + 32 * c < nchunks) {
                        float x[V];
                        NV::unpack(r[u][c], x);
                        const int k0 = (lane + 32 * c) * V;
#pragma unroll
                        for (int n = 0; n < NARROW_NMAX; ++n) {
                            if (n < N) {
#pragma unroll
                                for (int e4 = 0; e4 < V; e4 += 4) {
                                    const float4 wv = *reinterpret_cast<const float4 *>(Ws + n * K + k0 + e4);
                                    acc[n] = fmaf(x[e4], wv.x, acc[n]); acc[n] = fmaf(x[e4 + 1], wv.y, acc[n]);
                                    acc[n] = fmaf(x[e4 + 2], wv.z, acc[n]); acc[n] = fmaf(x[e4 + 3], wv.w, acc[n]);
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int n = 0; n < NARROW_NMAX; ++n) {
                    float v = acc[n];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == n && n < N) { v += __ldg(bias + n); q_s[m * N + n] = v; q_out[(long long)m * N + n] = v; }
                }
            }
        }
    }
    __syncthreads();
    // ---- head + loss (dqn_head_sample writes per_sample to global; the scale is re-read below)
    float part = 0.f;
    for (int b = threadIdx.x; b < h.B; b += 256) {
        dqn_head_sample(kind, double_dqn, h, b, q_s, q_t, gamma_n, dq_s);
        const float l = h.per_sample[b];
        part += l * sample_scale(h, b);
        if (prio_out) prio_out[b] = powf(__fadd_rn(l, 1e-10f), beta);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float sum = 0.f;
        for (int i = 0; i < 8; ++i) sum += red[i];
        loss[0] = sum;
    }
    for (int i = threadIdx.x; i < rows_bwd * N; i += 256) dq_out[i] = dq_s[i];
    pdl_trigger();
    // ---- backward of the head layer (as narrow_bwd_kernel; dY = dq_s)
    const int TPR = K >> 3, RPP = 256 / TPR;
    const int kq = threadIdx.x % TPR, rl = threadIdx.x / TPR;
    float w[NARROW_NMAX][8], dw[NARROW_NMAX][8], dbp[8], dbn[NARROW_NMAX];
#pragma unroll
    for (int n = 0; n < NARROW_NMAX; ++n) {
        dbn[n] = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) { w[n][e] = n < N ? Ws[n * K + kq * 8 + e] : 0.f; dw[n][e] = 0.f; }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) dbp[e] = 0.f;
    __syncthreads();  // every thread holds its W columns: Ws becomes the dW accumulator
    for (int i = threadIdx.x; i < NARROW_NMAX * K; i += 256) Ws[i] = 0.f;
    for (int m = rl; m < rows_bwd; m += RPP) {
        float x[8], dy[NARROW_NMAX];
        if (sizeof(T) == 2) {
            NarrowVec<bf16>::unpack(*reinterpret_cast<const uint4 *>(reinterpret_cast<const bf16 *>(X) + (long long)m * K + kq * 8), x);
        } else {
            const float4 a = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(X) + (long long)m * K + kq * 8);
            const float4 b = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(X) + (long long)m * K + kq * 8 + 4);
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        }
#pragma unroll
        for (int n = 0; n < NARROW_NMAX; ++n) dy[n] = n < N ? dq_s[m * N + n] : 0.f;
        float dx[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float s2 = 0.f;
#pragma unroll
            for (int n = 0; n < NARROW_NMAX; ++n) s2 = fmaf(dy[n], w[n][e], s2);
            if (relu_prev && !(x[e] > 0.f)) s2 = 0.f;
            dx[e] = s2;
            dbp[e] += s2;
#pragma unroll
            for (int n = 0; n < NARROW_NMAX; ++n) dw[n][e] = fmaf(dy[n], x[e], dw[n][e]);
        }
        if (sizeof(T) == 2) {
            *reinterpret_cast<uint4 *>(reinterpret_cast<bf16 *>(dX) + (long long)m * K + kq * 8) = NarrowVec<bf16>::pack(dx);
        } else {
            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(dX) + (long long)m * K + kq * 8) = make_float4(dx[0], dx[1], dx[2], dx[3]);
            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(dX) + (long long)m * K + kq * 8 + 4) = make_float4(dx[4], dx[5], dx[6], dx[7]);
        }
        if (kq == 0) {
#pragma unroll
            for (int n = 0; n < NARROW_NMAX; ++n) dbn[n] += dy[n];
        }
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < NARROW_NMAX; ++n) {
        if (n < N) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(Ws + n * K + kq * 8 + e, dw[n][e]);
            if (kq == 0) atomicAdd(dbn_s + n, dbn[n]);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(dbp_s + kq * 8 + e, dbp[e]);
    __syncthreads();
    // fire-and-forget reds (a read-modify-write here waits a DRAM round trip per element: 12 dependent trips per thread)
    flush_reds(dW, Ws, N * K, ((uintptr_t)dW & 15) == 0);
    if (db_prev) flush_reds(db_prev, dbp_s, K, ((uintptr_t)db_prev & 15) == 0);
    if (threadIdx.x < N) atomicAdd(db + threadIdx.x, dbn_s[threadIdx.x]);
}

// The same tail spread over the batch: CTA c owns the RPP = 2048 / K samples [c RPP, (c + 1) RPP) -- 4 at K = 512, so a batch of 32
// runs on 8 SMs -- and does for them what the single-CTA kernel does for the whole batch.  Every dependent global round trip
// happens once per CTA instead of once per group of rows: the head layer's input rows are read once (forward) and kept in shared
// memory for the backward.  The loss is summed in a fixed order by the last CTA to arrive (per-CTA partials + a ticket), so it is
// bit-reproducible.  It replaces narrow_fwd + dqn_head + loss_reduce + narrow_bwd (27 - 34 us of the B = 32 critical path).
template <typename T>
__global__ void __launch_bounds__(256) dqn_tail_multi_kernel(int kind, int double_dqn, HeadArgs h, const T *__restrict__ X, const float *__restrict__ W,
                                                             const float *__restrict__ bias, const float *__restrict__ q_t, float gamma_n, float beta,
                                                             int two_rows, int bwd_next, int K, int relu_prev, float *__restrict__ q_out,
                                                             float *__restrict__ dq_out, T *__restrict__ dX, float *__restrict__ dW, float *__restrict__ db,
                                                             float *__restrict__ db_prev, float *__restrict__ loss, float *__restrict__ prio_out,
                                                             float *__restrict__ part, unsigned *__restrict__ ticket) {
    extern __shared__ float sm_[];
    const int N = h.A, B = h.B;
    const int TPR = K >> 3, RPP = 256 / TPR;           // threads per row, samples per CTA
    float *Ws = sm_;                                   // [N][K]  (later the dW accumulator)
    float *dbp_s = Ws + NARROW_NMAX * K;               // [K]
    T *Xs = reinterpret_cast<T *>(dbp_s + K);          // [2 RPP][K]: rows of s, then rows of s'  (RPP K = 2048 elements)
    float *q_s = dbp_s + K + 2 * 2048;                 // [2 RPP][N]
    float *dq_s = q_s + 64 * NARROW_NMAX;              // [2 RPP][N]
    float *dbn_s = dq_s + 64 * NARROW_NMAX;            // [N]
    float *ls_s = dbn_s + NARROW_NMAX;                 // [RPP] scaled per-sample losses
    float *qt_s = ls_s + 32;                           // [RPP][N] target Q(s') of this CTA's samples
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b0 = blockIdx.x * RPP;
    const int nloc = min(RPP, B - b0);                 // samples of this CTA
    const int nrows = two_rows ? 2 * RPP : RPP;
    pdl_wait();
    for (int i = threadIdx.x; i < N * K / 4; i += 256) reinterpret_cast<float4 *>(Ws)[i] = __ldg(reinterpret_cast<const float4 *>(W) + i);
    for (int i = threadIdx.x; i < K + NARROW_NMAX; i += 256) (i < K ? dbp_s[i] : dbn_s[i - K]) = 0.f;
    for (int i = threadIdx.x; i < 64 * NARROW_NMAX; i += 256) dq_s[i] = 0.f;
    if (q_t && threadIdx.x < nloc * N) qt_s[threadIdx.x] = __ldg(q_t + (size_t)b0 * N + threadIdx.x);
    // ---- this CTA's input rows: global -> registers -> shared (all loads of a thread in flight together)
    {
        constexpr int V = NarrowVec<T>::V;
        const int cpr = K / V;  // 16-byte chunks per row
        typename NarrowVec<T>::raw r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = threadIdx.x + 256 * u;
            const int ri = j / cpr, ch = j - ri * cpr;
            if (ri < nrows) {
                const int loc = ri < RPP ? ri : ri - RPP;
                const long long m = ri < RPP ? b0 + loc : (long long)B + b0 + loc;
                r[u] = loc < nloc ? reinterpret_cast<const typename NarrowVec<T>::raw *>(X + m * K)[ch] : typename NarrowVec<T>::raw{};
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = threadIdx.x + 256 * u;
            if (j / cpr < nrows) reinterpret_cast<typename NarrowVec<T>::raw *>(Xs)[j] = r[u];
        }
    }
    __syncthreads();
    // ---- forward: one warp per row
    {
        using NV = NarrowVec<T>;
        constexpr int V = NV::V;
        const int nchunks = K / V;
        for (int ri = warp; ri < nrows; ri += 8) {
            const int loc = ri < RPP ? ri : ri - RPP;
            if (loc >= nloc) continue;
            float acc[NARROW_NMAX];
#pragma unroll
            for (int n = 0; n < NARROW_NMAX; ++n) acc[n] = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (lane + 32 * c < nchunks) {
                    float x[V];
                    NV::unpack(reinterpret_cast<const typename NV::raw *>(Xs + (size_t)ri * K)[lane + 32 * c], x);
                    const int k0 = (lane + 32 * c) * V;
#pragma unroll
                    for (int n = 0; n < NARROW_NMAX; ++n) {
                        if (n < N) {
#pragma unroll
                            for (int e4 = 0; e4 < V; e4 += 4) {
                                const float4 wv = *reinterpret_cast<const float4 *>(Ws + n * K + k0 + e4);
                                acc[n] = fmaf(x[e4], wv.x, acc[n]); acc[n] = fmaf(x[e4 + 1], wv.y, acc[n]);
                                acc[n] = fmaf(x[e4 + 2], wv.z, acc[n]); acc[n] = fmaf(x[e4 + 3], wv.w, acc[n]);
                            }
                        }
                    }
                }
            }
            const long long m = ri < RPP ? b0 + loc : (long long)B + b0 + loc;
#pragma unroll
            for (int n = 0; n < NARROW_NMAX; ++n) {
                float v = acc[n];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == n && n < N) { v += __ldg(bias + n); q_s[ri * N + n] = v; q_out[m * N + n] = v; }
            }
        }
    }
    __syncthreads();
    // ---- head of this CTA's samples; the partial loss goes to part[c], the last CTA to arrive sums the partials in order
    if (threadIdx.x < nloc) {
        const int loc = threadIdx.x, b = b0 + loc;
        dqn_head_rows(kind, double_dqn, h, b, q_s + loc * N, q_s + (RPP + loc) * N, q_t ? qt_s + loc * N : nullptr, gamma_n, dq_s + loc * N,
                      dq_s + (RPP + loc) * N);
        const float l = h.per_sample[b];
        ls_s[loc] = l * sample_scale(h, b);
        if (prio_out) prio_out[b] = powf(__fadd_rn(l, 1e-10f), beta);
        for (int n = 0; n < N; ++n) {
            dq_out[(size_t)b * N + n] = dq_s[loc * N + n];
            if (bwd_next) dq_out[(size_t)(B + b) * N + n] = dq_s[(RPP + loc) * N + n];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float sum = 0.f;
        for (int i = 0; i < nloc; ++i) sum += ls_s[i];
        part[blockIdx.x] = sum;
        __threadfence();
        const unsigned old = atomicAdd(ticket, 1u);
        if (old == gridDim.x - 1) {
            __threadfence();
            float tot = 0.f;
            for (unsigned c = 0; c < gridDim.x; ++c) tot += __ldcg(part + c);
            loss[0] = tot;
            *ticket = 0u;
        }
    }
    pdl_trigger();
    // ---- backward of the head layer (as narrow_bwd_kernel; rows from shared memory, dY = dq_s)
    const int kq = threadIdx.x % TPR, rl = threadIdx.x / TPR;
    float w[NARROW_NMAX][8], dw[NARROW_NMAX][8], dbp[8], dbn[NARROW_NMAX];
#pragma unroll
    for (int n = 0; n < NARROW_NMAX; ++n) {
        dbn[n] = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) { w[n][e] = n < N ? Ws[n * K + kq * 8 + e] : 0.f; dw[n][e] = 0.f; }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) dbp[e] = 0.f;
    __syncthreads();  // every thread holds its W columns: Ws becomes the dW accumulator
    for (int i = threadIdx.x; i < NARROW_NMAX * K; i += 256) Ws[i] = 0.f;
    for (int pass = 0; pass < (bwd_next ? 2 : 1); ++pass) {
        if (rl >= nloc) break;
        const int ri = pass * RPP + rl;
        const long long m = pass ? (long long)B + b0 + rl : b0 + rl;
        float x[8], dy[NARROW_NMAX];
        if (sizeof(T) == 2) {
            NarrowVec<bf16>::unpack(*reinterpret_cast<const uint4 *>(reinterpret_cast<const bf16 *>(Xs) + (size_t)ri * K + kq * 8), x);
        } else {
            const float4 a = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(Xs) + (size_t)ri * K + kq * 8);
            const float4 b = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(Xs) + (size_t)ri * K + kq * 8 + 4);
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        }
#pragma unroll
        for (int n = 0; n < NARROW_NMAX; ++n) dy[n] = n < N ? dq_s[ri * N + n] : 0.f;
        float dx[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float s2 = 0.f;
#pragma unroll
            for (int n = 0; n < NARROW_NMAX; ++n) s2 = fmaf(dy[n], w[n][e], s2);
            if (relu_prev && !(x[e] > 0.f)) s2 = 0.f;
            dx[e] = s2;
            dbp[e] += s2;
#pragma unroll
            for (int n = 0; n < NARROW_NMAX; ++n) dw[n][e] = fmaf(dy[n], x[e], dw[n][e]);
        }
        if (sizeof(T) == 2) {
            *reinterpret_cast<uint4 *>(reinterpret_cast<bf16 *>(dX) + m * K + kq * 8) = NarrowVec<bf16>::pack(dx);
        } else {
            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(dX) + m * K + kq * 8) = make_float4(dx[0], dx[1], dx[2], dx[3]);
            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(dX) + m * K + kq * 8 + 4) = make_float4(dx[4], dx[5], dx[6], dx[7]);
        }
        if (kq == 0) {
#pragma unroll
            for (int n = 0; n < NARROW_NMAX; ++n) dbn[n] += dy[n];
        }
    }
    __syncthreads();
    if (rl < nloc) {
#pragma unroll
        for (int n = 0; n < NARROW_NMAX; ++n) {
            if (n < N) {
#pragma unroll
                for (int e = 0; e < 8; ++e) atomicAdd(Ws + n * K + kq * 8 + e, dw[n][e]);
                if (kq == 0) atomicAdd(dbn_s + n, dbn[n]);
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(dbp_s + kq * 8 + e, dbp[e]);
    }
    __syncthreads();
    flush_reds(dW, Ws, N * K, ((uintptr_t)dW & 15) == 0);
    if (db_prev) flush_reds(db_prev, dbp_s, K, ((uintptr_t)db_prev & 15) == 0);
    if (threadIdx.x < N) atomicAdd(db + threadIdx.x, dbn_s[threadIdx.x]);
}

bool dqn_tail_multi_supported(int N, long long K, int x_bf16, int B) {
    // RPP <= 32 samples per CTA (K >= 64), at most 64 CTAs (the partial-loss scratch)
    return narrow_supported(N, K, x_bf16) && K >= 64 && cdiv(B, 2048 / (int)K) <= 64;
}

// part: >= 64 floats of scratch, ticket: one zero-initialised counter (left zero)
zoo_status dqn_tail_multi(int kind, int double_dqn, const HeadArgs &h, const void *X, int x_bf16, const float *W, const float *bias, const float *q_t,
                          float gamma_n, float beta, int two_rows, int bwd_next, long long K, int relu_prev, float *q_out, float *dq_out, void *dX,
                          float *dW, float *db, float *db_prev, float *loss, float *prio_out, float *part, unsigned *ticket, cudaStream_t st) {
    ZOO_REQUIRE(dqn_tail_multi_supported(h.A, K, x_bf16, h.B) && (!bwd_next || two_rows), "dqn_tail_multi: unsupported shape");
    const size_t smem = (size_t)(NARROW_NMAX * K + K + 2 * 2048 + 2 * 64 * NARROW_NMAX + NARROW_NMAX + 32 + 32 * NARROW_NMAX) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        const int mx = (int)((NARROW_NMAX * 2048 + 2048 + 2 * 2048 + 2 * 64 * NARROW_NMAX + NARROW_NMAX + 32 + 32 * NARROW_NMAX) * sizeof(float));
        ZOO_CUDA(cudaFuncSetAttribute(dqn_tail_multi_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        ZOO_CUDA(cudaFuncSetAttribute(dqn_tail_multi_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        attr = true;
    }
    const dim3 grid(cdiv(h.B, 2048 / (int)K));
    if (x_bf16)
        ZOO_CUDA(launch_pdl(dqn_tail_multi_kernel<bf16>, grid, dim3(256), smem, st, kind, double_dqn, h, (const bf16 *)X, W, bias, q_t, gamma_n, beta, two_rows,
                            bwd_next, (int)K, relu_prev, q_out, dq_out, (bf16 *)dX, dW, db, db_prev, loss, prio_out, part, ticket));
    else
        ZOO_CUDA(launch_pdl(dqn_tail_multi_kernel<float>, grid, dim3(256), smem, st, kind, double_dqn, h, (const float *)X, W, bias, q_t, gamma_n, beta, two_rows,
                            bwd_next, (int)K, relu_prev, q_out, dq_out, (float *)dX, dW, db, db_prev, loss, prio_out, part, ticket));
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

bool dqn_tail_supported(int N, long long K, int x_bf16, int rows_fwd) {
    return narrow_supported(N, K, x_bf16) && rows_fwd <= TAIL_ROWS;
}

zoo_status dqn_tail(int kind, int double_dqn, const HeadArgs &h, const void *X, int x_bf16, const float *W, const float *bias, const float *q_t, float gamma_n,
                    float beta, int rows_fwd, int rows_bwd, long long K, int relu_prev, float *q_out, float *dq_out, void *dX, float *dW, float *db,
                    float *db_prev, float *loss, float *prio_out, cudaStream_t st) {
    ZOO_REQUIRE(dqn_tail_supported(h.A, K, x_bf16, rows_fwd) && rows_bwd <= rows_fwd, "dqn_tail: unsupported shape");
    const size_t smem = (size_t)(NARROW_NMAX * K + K + 2 * TAIL_ROWS * NARROW_NMAX + 32 + NARROW_NMAX) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        const int mx = (int)((NARROW_NMAX * 2048 + 2048 + 2 * TAIL_ROWS * NARROW_NMAX + 32 + NARROW_NMAX) * sizeof(float));
        ZOO_CUDA(cudaFuncSetAttribute(dqn_tail_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        ZOO_CUDA(cudaFuncSetAttribute(dqn_tail_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        attr = true;
    }
    if (x_bf16)
        ZOO_CUDA(launch_pdl(dqn_tail_kernel<bf16>, dim3(1), dim3(256), smem, st, kind, double_dqn, h, (const bf16 *)X, W, bias, q_t, gamma_n, beta, rows_fwd, rows_bwd,
                            (int)K, relu_prev, q_out, dq_out, (bf16 *)dX, dW, db, db_prev, loss, prio_out));
    else
        ZOO_CUDA(launch_pdl(dqn_tail_kernel<float>, dim3(1), dim3(256), smem, st, kind, double_dqn, h, (const float *)X, W, bias, q_t, gamma_n, beta, rows_fwd,
                            rows_bwd, (int)K, relu_prev, q_out, dq_out, (float *)dX, dW, db, db_prev, loss, prio_out));
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

}  // namespace zoo
