// fp32 CUDA-core GEMM engine (ZOO_PREC_FP32 path and all contractions too small / misaligned for
// tcgen05).  D(m,n) = sum_k A(m,k) B(n,k), fp32 accumulate, operands fp32 or bf16 in either major.
// Used for: Dense / CrossCor-as-GEMM forward, dgrad, wgrad of the Q-networks
// (reference: Flux Dense/CrossCor in src/experiments/atari/Dopamine_DQN_Atari.jl:26-35).
#include "common.cuh"

namespace zoo {

static constexpr int SBM = 64, SBN = 64, SBK = 16;

template <bool AK, bool BKM>
__global__ void __launch_bounds__(256)
simt_gemm_kernel(Operand A, Operand B, int M, int N, int K, Epilogue epi, int k_per_split, float *ws) {
    __shared__ __align__(16) float As[SBK][SBM + 4];
    __shared__ __align__(16) float Bs[SBK][SBN + 4];
    pdl_wait();
    pdl_trigger();
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = kbeg; k0 < kend; k0 += SBK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int i, k;
            if (AK) { k = tid & 15; i = (tid >> 4) + 16 * j; }
            else { i = tid & 63; k = (tid >> 6) + 4 * j; }
            float v = 0.f;
            if (m0 + i < M && k0 + k < kend)
                v = AK ? ld_elem(A.p, A.is_bf16, (long long)(m0 + i) * A.ld + k0 + k)
                       : ld_elem(A.p, A.is_bf16, (long long)(k0 + k) * A.ld + m0 + i);
            As[k][i] = v;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int i, k;
            if (BKM) { k = tid & 15; i = (tid >> 4) + 16 * j; }
            else { i = tid & 63; k = (tid >> 6) + 4 * j; }
            float v = 0.f;
            if (n0 + i < N && k0 + k < kend)
                v = BKM ? ld_elem(B.p, B.is_bf16, (long long)(n0 + i) * B.ld + k0 + k)
                        : ld_elem(B.p, B.is_bf16, (long long)(k0 + k) * B.ld + n0 + i);
            Bs[k][i] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SBK; ++kk) {
            float4 a4 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            float4 b4 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            if (ws) atomicAdd(ws + (long long)m * N + n, acc[i][j]);
            else epi_store(epi, acc[i][j], m, n);
        }
    }
}

// Narrow output layers (fc2: N = n_actions): one CTA per row, every thread strides over K with N accumulators, warp-shuffle +
// shared-memory reduction, bias / relu in place.  One launch, no split-K workspace.
template <int NMAX>
__global__ void __launch_bounds__(128) dense_narrow_kernel(Operand A, Operand B, int N, int K, Epilogue epi) {
    pdl_wait();
    pdl_trigger();
    const int m = blockIdx.x;
    float acc[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
    for (int k = threadIdx.x; k < K; k += 128) {
        const float a = ld_elem(A.p, A.is_bf16, (long long)m * A.ld + k);
#pragma unroll
        for (int n = 0; n < NMAX; ++n)
            if (n < N) acc[n] = fmaf(a, ld_elem(B.p, B.is_bf16, (long long)n * B.ld + k), acc[n]);
    }
    __shared__ float red[4][NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
        float v = acc[n];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][n] = v;
    }
    __syncthreads();
    if (threadIdx.x < N) epi_store(epi, red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x], m, threadIdx.x);
}

// many rows (IQN: batch x quantiles): one warp per row, grid-stride, no shared memory
template <int NMAX>
__global__ void __launch_bounds__(256) dense_narrow_rows_kernel(Operand A, Operand B, int M, int N, int K, Epilogue epi) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    for (int m = blockIdx.x * 8 + (threadIdx.x >> 5); m < M; m += gridDim.x * 8) {
        float acc[NMAX];
#pragma unroll
        for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float a = ld_elem(A.p, A.is_bf16, (long long)m * A.ld + k);
#pragma unroll
            for (int n = 0; n < NMAX; ++n)
                if (n < N) acc[n] = fmaf(a, ldg_elem(B.p, B.is_bf16, (long long)n * B.ld + k), acc[n]);
        }
#pragma unroll
        for (int n = 0; n < NMAX; ++n) {
            float v = acc[n];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == n && n < N) epi_store(epi, v, m, n);  // (n >= N would land in the next row's first columns)
        }
    }
}

// finishes a split-K GEMM whose partial sums were atomically accumulated in ws[M][N]
// (and re-zeroes ws: the split-K workspace is all-zero between launches, shared with the tcgen05 engine)
__global__ void splitk_finish_kernel(float *ws, int M, int N, Epilogue epi) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * N) return;
    int m = (int)(i / N), n = (int)(i % N);
    const float v = ws[i];
    ws[i] = 0.f;
    epi_store(epi, v, m, n);
}

zoo_status splitk_finish(float *ws, int M, int N, const Epilogue &epi, cudaStream_t st) {
    long long tot = (long long)M * N;
    splitk_finish_kernel<<<cdiv(tot, 256), 256, 0, st>>>(ws, M, N, epi);
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

zoo_status simt_gemm(const Operand &A, const Operand &B, int M, int N, int K, const Epilogue &epi, int split_k,
                     float *ws, size_t ws_elems, cudaStream_t st) {
    if (M <= 0 || N <= 0) return ZOO_OK;
    if (N <= 16 && A.kmajor && B.kmajor && !epi.accumulate && split_k <= 0 && M > 1024) {
        const int blocks = (int)std::min<long long>(cdiv(M, 8), 148 * 8);
        if (N <= 8) ZOO_CUDA(launch_pdl(dense_narrow_rows_kernel<8>, dim3(blocks), dim3(256), (size_t)0, st, A, B, M, N, K, epi));
        else ZOO_CUDA(launch_pdl(dense_narrow_rows_kernel<16>, dim3(blocks), dim3(256), (size_t)0, st, A, B, M, N, K, epi));
        ZOO_LAUNCH_CHECK();
        return ZOO_OK;
    }
    if (N <= 16 && A.kmajor && B.kmajor && !epi.accumulate && split_k <= 0 && M <= 65535) {
        if (N <= 8) ZOO_CUDA(launch_pdl(dense_narrow_kernel<8>, dim3(M), dim3(128), (size_t)0, st, A, B, N, K, epi));
        else ZOO_CUDA(launch_pdl(dense_narrow_kernel<16>, dim3(M), dim3(128), (size_t)0, st, A, B, N, K, epi));
        ZOO_LAUNCH_CHECK();
        return ZOO_OK;
    }
    dim3 grid(cdiv(M, SBM), cdiv(N, SBN), 1);
    int kblocks = cdiv(K, SBK);
    if (split_k <= 0) {  // auto: fill ~2 waves of 148 SMs when the output grid alone cannot
        long long tiles = (long long)grid.x * grid.y;
        split_k = 1;
        if (tiles < 148 && kblocks >= 8) split_k = (int)std::min<long long>(kblocks / 4, (296 + tiles - 1) / tiles);
        if (split_k < 1) split_k = 1;
    }
    split_k = std::min(split_k, kblocks);
    int kps = cdiv(kblocks, split_k) * SBK;
    split_k = cdiv(K, kps);
    grid.z = split_k;
    float *wsp = nullptr;
    bool direct_acc = epi.accumulate && !epi.out_bf16;
    Epilogue e = epi;
    if (split_k == 1 && e.accumulate == 2) e.accumulate = 0;  // "overwrite or accumulate": nothing to sum
    if (split_k > 1) {
        if (direct_acc) {
            // accumulate straight into the (fp32) destination: every split adds its partial sum
            wsp = nullptr;
        } else {
            ZOO_REQUIRE(ws && ws_elems >= (size_t)M * N, "simt_gemm: split-K workspace too small (%zu < %lld)", ws_elems,
                        (long long)M * N);
            wsp = ws;  // all-zero on entry (invariant), re-zeroed by splitk_finish
        }
    }
#define LAUNCH(AKv, BKv) ZOO_CUDA(launch_pdl(simt_gemm_kernel<AKv, BKv>, grid, dim3(256), (size_t)0, st, A, B, M, N, K, e, kps, wsp))
    if (A.kmajor && B.kmajor) LAUNCH(true, true);
    else if (A.kmajor && !B.kmajor) LAUNCH(true, false);
    else if (!A.kmajor && B.kmajor) LAUNCH(false, true);
    else LAUNCH(false, false);
#undef LAUNCH
    ZOO_LAUNCH_CHECK();
    if (wsp) ZOO_TRY(splitk_finish(wsp, M, N, epi, st));
    return ZOO_OK;
}

zoo_status gemm(const Operand &A, const Operand &B, int M, int N, int K, const Epilogue &epi, int split_k, float *ws,
                size_t ws_elems, cudaStream_t st) {
    const bool want_tc = (A.is_bf16 && B.is_bf16) || (!A.is_bf16 && !B.is_bf16 && epi.tf32_passes > 0);
    if (want_tc && M <= 64 && N >= 128 && !epi.accumulate && !epi.rowmul && tc_gemm_supported(B, A, N, M, K)) {
        // swap-AB for a skinny batch against a wide layer: the weight matrix becomes the 128-row MMA operand (every byte of
        // its tile is payload, instead of a 128-row activation tile that is mostly padding) and the result is written transposed
        Epilogue t = epi;
        t.sm = epi.sn; t.sn = epi.sm;
        t.mask_sm = epi.mask_sn; t.mask_sn = epi.mask_sm;
        t.bias_mode = epi.bias_mode == 1 ? 2 : (epi.bias_mode == 2 ? 1 : 0);
        return tc_gemm(B, A, N, M, K, t, split_k, ws, ws_elems, st);
    }
    if (want_tc && split_k <= 0 && tc_gemm_supported(A, B, M, N, K) && tc_gemm_persist_ok(A, B, M, N, K, epi)) return tc_gemm_persist(A, B, M, N, K, epi, st);
    if (want_tc && tc_gemm_supported(A, B, M, N, K)) return tc_gemm(A, B, M, N, K, epi, split_k, ws, ws_elems, st);
    return simt_gemm(A, B, M, N, K, epi, split_k, ws, ws_elems, st);
}

}  // namespace zoo
