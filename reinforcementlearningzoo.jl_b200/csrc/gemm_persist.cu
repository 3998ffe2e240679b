// Persistent tcgen05 GEMM for the large-batch bf16 contractions (IQN header / phi layers at batch x quantiles = 32 768 rows,
// the fc1 gradients of Rainbow at B = 1024): D(m,n) = sum_k A(m,k) B(n,k), bf16 operands, fp32 accumulation in TMEM.
//
// Why a second kernel: the one-tile-per-CTA engine (gemm_tc.cuh) pays its fixed costs per 128 x 128 tile -- barrier init, TMEM
// allocation, first TMA round trip, and an epilogue that runs with the tensor pipe idle.  At 15 616 tiles (phi forward, fc1
// data gradient) that is ~4.5 us per tile against a 0.4-2 us main loop (profiles/r02g: TMA stores on or off, 237 vs 240 us).
// Here one CTA per SM walks its share of 128 x 256 tiles:
//   warp 0      TMA producer: one ring of 3 x 48 KB stages runs continuously across tile boundaries
//   warp 1      tcgen05.mma issuer (cta_group::1, M = 128, N = 256, K = 16): accumulator t & 1 of two 256-column TMEM buffers
//   warps 2..5  epilogue of tile t while the MMA warp is already on tile t + 1: tcgen05.ld (lane = row, 32 columns) -> bias /
//               relu / relu' mask / IQN row multiplier in registers -> 16-byte chunks into a SWIZZLE_128B staging block ->
//               one TMA store per 32-row x 128-byte block (double-buffered per warp)
// N = 256 per instruction also halves the shared-memory bytes per flop of the B operand relative to two N = 128 tiles; the main
// loop of this engine is bound by shared-memory bandwidth (TMA writes + UMMA reads), see DESIGN.md section 4.
// Reference layers served: ImplicitQuantileNet phi / header Dense layers (src/algorithms/dqns/iqn.jl:25-33,
// src/experiments/atari/Dopamine_IQN_Atari.jl:36-43), Dense(7744, 512) of the Nature-CNN (Dopamine_DQN_Atari.jl:33).
#include "gemm_tc.cuh"

namespace zoo {

static constexpr int PBN = 256;               // tile columns = UMMA N
static constexpr int PST = 3;                 // ring depth
static constexpr int P_A_BYTES = BM * 128;    // 16 KB
static constexpr int P_B_BYTES = PBN * 128;   // 32 KB
static constexpr int P_STAGE = P_A_BYTES + P_B_BYTES;
static constexpr int P_EPI_OFF = PST * P_STAGE;             // epilogue staging: 4 warps x 16 KB (a warp's whole 32 x 256 slab: 4 blocks of 32 rows x 128 B)
static constexpr int P_BIAS_OFF = P_EPI_OFF + 4 * 16384;     // per warp: bias [256] + rowmul [256] floats
static constexpr int P_BAR_OFF = P_BIAS_OFF + 4 * 2 * PBN * 4;
static constexpr int P_SMEM = P_BAR_OFF + 256;

struct PersistParams {
    int M, N, K;
    int ntn, ntiles;   // tiles along N, total tiles
    Epilogue epi;
    CUtensorMap tmC;   // output: box {128 bytes of columns, 32 rows}, SWIZZLE_128B
};

template <bool A_K, bool B_K>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ PersistParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint64_t *full_bar = (uint64_t *)(smem + P_BAR_OFF);
    uint64_t *empty_bar = full_bar + PST;
    uint64_t *acc_full = empty_bar + PST;   // [2]
    uint64_t *acc_empty = acc_full + 2;     // [2]
    uint32_t *tmem_slot = (uint32_t *)(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int BK = 64, MN_BOX = 64, BOX_BYTES = 64 * 128, KSTEP_MN = 16 * 128;
    const int total_kb = (P.K + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < PST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
            int it = 0;
            for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
                const int m0 = (tile / P.ntn) * BM, n0 = (tile % P.ntn) * PBN;
                for (int kb = 0; kb < total_kb; ++kb, ++it) {
                    const int s = it % PST;
                    mbar_wait(&empty_bar[s], ((it / PST) & 1) ^ 1);
                    uint8_t *sa = smem + s * P_STAGE, *sb = sa + P_A_BYTES;
                    mbar_expect_tx(&full_bar[s], P_STAGE);
                    const int k0 = kb * BK;
                    if (A_K) tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);
                    else {
#pragma unroll
                        for (int h = 0; h < BM / MN_BOX; ++h) tma_load_2d(sa + h * BOX_BYTES, &tmA, &full_bar[s], m0 + h * MN_BOX, k0);
                    }
                    if (B_K) tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);
                    else {
#pragma unroll
                        for (int h = 0; h < PBN / MN_BOX; ++h) tma_load_2d(sb + h * BOX_BYTES, &tmB, &full_bar[s], n0 + h * MN_BOX, k0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(PBN, !A_K, !B_K, 1u);
            int it = 0, t = 0;
            for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++t) {
                const int ab = t & 1;
                mbar_wait(&acc_empty[ab], ((t >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator (first use: passes)
                tc_fence_after();
                const uint32_t td = tmem_base + (uint32_t)(ab * PBN);
                for (int kb = 0; kb < total_kb; ++kb, ++it) {
                    const int s = it % PST;
                    mbar_wait(&full_bar[s], (it / PST) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * P_STAGE), sb = sa + P_A_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t ao = A_K ? k * 32 : k * KSTEP_MN, bo = B_K ? k * 32 : k * KSTEP_MN;
                        const uint64_t ad = A_K ? make_smem_desc(sa + ao, 16, 1024) : make_smem_desc(sa + ao, BOX_BYTES, 1024, 2);
                        const uint64_t bd = B_K ? make_smem_desc(sb + bo, 16, 1024) : make_smem_desc(sb + bo, BOX_BYTES, 1024, 2);
                        umma_bf16(td, ad, bd, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&acc_full[ab]);
            }
        }
    } else {
        // epilogue warps: TMEM lane quadrant = warp % 4, lane = tile row
        const int q = warp & 3;
        const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
        const Epilogue e = P.epi;
        const int Nn = P.N, out_bf16 = e.out_bf16, relu = e.relu, bias_mode = e.bias_mode;
        const void *const maskp = e.mask;
        const long long msm = e.mask_sm;
        const bool rmul = e.rowmul != nullptr;
        // everything below is private to the warp (its own bias / multiplier copy, its own staging slab): no CTA-level barrier in
        // the tile loop.  Per tile: the bias / multiplier loads are issued BEFORE the wait for the accumulator, the warp's 32 x 256
        // slab is converted into four SWIZZLE_128B blocks, ONE proxy fence, then one TMA store per block (fp32 output: two rounds of
        // four 32-column blocks); the stores of tile t are only waited for when tile t + 1 is about to overwrite the slab.
        float *const bias_s = reinterpret_cast<float *>(smem + P_BIAS_OFF) + q * 2 * PBN;
        float *const rm_s = bias_s + PBN;
        const uint32_t sbase = smem_u32(smem) + P_EPI_OFF + q * 16384;
        const int cpg = out_bf16 ? 64 : 32;  // columns per 128-byte staging row
        int t = 0;
        bool pending = false;
        for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++t) {
            const int ab = t & 1;
            const int m0 = (tile / P.ntn) * BM, n0 = (tile % P.ntn) * PBN;
            const int row0 = m0 + q * 32;
            float bv[PBN / 32], rv[PBN / 32];
#pragma unroll
            for (int j = 0; j < PBN / 32; ++j) {
                const int n = n0 + lane + 32 * j;
                bv[j] = (bias_mode == 1 && n < Nn) ? __ldg(e.bias + n) : 0.f;
                rv[j] = (rmul && n < Nn) ? ldg_elem(e.rowmul, out_bf16, (long long)(row0 / e.rowmul_div) * e.rowmul_ld + n) : 1.f;
            }
            mbar_wait(&acc_full[ab], (t >> 1) & 1);
            tc_fence_after();
            if (pending) {  // the previous tile's stores have read the slab
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                pending = false;
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < PBN / 32; ++j) { bias_s[lane + 32 * j] = bv[j]; rm_s[lane + 32 * j] = rv[j]; }
            __syncwarp();
            const bool rowok = row0 + lane < P.M;
            const int ncols = min(PBN, Nn - n0);
#pragma unroll 1
            for (int half = 0; half < (out_bf16 ? 1 : 2); ++half) {  // fp32: 128 columns (4 blocks of 32) per round
                if (half * 128 >= ncols) break;
                if (half == 1) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                }
#pragma unroll 1
                for (int blk = 0; blk < 4; ++blk) {
                    const int c = half * 128 + blk * cpg;
                    if (c >= ncols) break;
                    const uint32_t sb = sbase + blk * 4096 + lane * 128;
#pragma unroll 1
                    for (int h = 0; h < cpg / 32; ++h) {
                        uint32_t r[32];
                        tmem_ld32(tq + (uint32_t)(ab * PBN + c + 32 * h), r);
                        const int nb = n0 + c + 32 * h;
                        uint4 mk[4];
                        if (maskp) {  // relu' source: bf16 [M][msm], same rows / columns as the output
                            const bf16 *mp = (const bf16 *)maskp + (long long)(row0 + lane) * msm + nb;
                            if (rowok && nb + 32 <= Nn) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) mk[j] = __ldg(reinterpret_cast<const uint4 *>(mp) + j);
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    uint32_t w[4];
#pragma unroll
                                    for (int i = 0; i < 4; ++i) {
                                        const int col = j * 8 + 2 * i;
                                        const uint32_t lo = (rowok && nb + col < Nn) ? (uint32_t)__bfloat16_as_ushort(__ldg(mp + col)) : 0u;
                                        const uint32_t hi = (rowok && nb + col + 1 < Nn) ? (uint32_t)__bfloat16_as_ushort(__ldg(mp + col + 1)) : 0u;
                                        w[i] = lo | (hi << 16);
                                    }
                                    mk[j] = make_uint4(w[0], w[1], w[2], w[3]);
                                }
                            }
                        }
#pragma unroll
                        for (int j8 = 0; j8 < 4; ++j8) {
                            const uint32_t mw[4] = {mk[j8].x, mk[j8].y, mk[j8].z, mk[j8].w};
                            const int col0 = c + 32 * h + j8 * 8;
                            const float4 b0 = *reinterpret_cast<const float4 *>(bias_s + col0), b1 = *reinterpret_cast<const float4 *>(bias_s + col0 + 4);
                            const float4 m0v = *reinterpret_cast<const float4 *>(rm_s + col0), m1v = *reinterpret_cast<const float4 *>(rm_s + col0 + 4);
                            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                            const float mm[8] = {m0v.x, m0v.y, m0v.z, m0v.w, m1v.x, m1v.y, m1v.z, m1v.w};
                            float x[8];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float x0 = __uint_as_float(r[j8 * 8 + 2 * i]) + bb[2 * i], x1 = __uint_as_float(r[j8 * 8 + 2 * i + 1]) + bb[2 * i + 1];
                                const float r0f = fmaxf(x0, 0.f), r1f = fmaxf(x1, 0.f);
                                x0 = relu ? r0f : x0;
                                x1 = relu ? r1f : x1;
                                if (maskp) {
                                    x0 = __uint_as_float(mw[i] << 16) > 0.f ? x0 : 0.f;
                                    x1 = __uint_as_float(mw[i] & 0xffff0000u) > 0.f ? x1 : 0.f;
                                }
                                x[2 * i] = x0 * mm[2 * i];
                                x[2 * i + 1] = x1 * mm[2 * i + 1];
                            }
                            if (out_bf16) {  // 8 values = one 16-byte chunk; chunk index within the 128-byte row = h * 4 + j8
                                uint32_t pk[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
                                    pk[i] = *reinterpret_cast<const uint32_t *>(&hh);
                                }
                                const int jc = h * 4 + j8;
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + ((uint32_t)(jc ^ (lane & 7)) << 4)), "r"(pk[0]), "r"(pk[1]),
                                             "r"(pk[2]), "r"(pk[3]) : "memory");
                            } else {         // fp32: 8 values = two chunks; 32 columns fill the 128-byte row (h == 0)
#pragma unroll
                                for (int hc = 0; hc < 2; ++hc) {
                                    const int jc = j8 * 2 + hc;
                                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + ((uint32_t)(jc ^ (lane & 7)) << 4)),
                                                 "r"(__float_as_uint(x[4 * hc])), "r"(__float_as_uint(x[4 * hc + 1])), "r"(__float_as_uint(x[4 * hc + 2])),
                                                 "r"(__float_as_uint(x[4 * hc + 3])) : "memory");
                                }
                            }
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    for (int blk = 0; blk < 4; ++blk) {
                        const int c = half * 128 + blk * cpg;
                        if (c >= ncols) break;
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&P.tmC),
                                     "r"(sbase + blk * 4096), "r"(n0 + c), "r"(row0)
                                     : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                pending = true;
            }
            // every TMEM read of this tile is done (tcgen05.ld waits inside tmem_ld32): hand the accumulator back
            tc_fence_before();
            mbar_arrive(&acc_empty[ab]);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static bool persist_enabled() {
    static const bool on = [] { const char *e = getenv("ZOO_PERSIST_GEMM"); return !(e && e[0] == '0'); }();
    return on;
}

static int persist_min_k() {
    static const int k = [] { const char *e = getenv("ZOO_PERSIST_MIN_K"); return e ? atoi(e) : 256; }();
    return k;
}

// host: the shapes this engine takes (everything else stays with tc_gemm)
bool tc_gemm_persist_ok(const Operand &A, const Operand &B, int M, int N, int K, const Epilogue &epi) {
    if (!persist_enabled() || !A.is_bf16 || !B.is_bf16) return false;
    const long long tiles = (long long)cdiv(M, BM) * cdiv(N, PBN);
    // enough work to keep every SM busy without a K split -- and a main loop long enough to hide the epilogue of the previous tile
    // behind it (measured: the K = 64 phi GEMM, one K-block per tile, runs 428 us here against 237 us on the one-tile-per-CTA engine)
    if (N < 256 || K < persist_min_k() || tiles < 100 || (tiles < 2 * g_sm_budget && K < 4096)) return false;
    if (epi.accumulate == 1 || epi.bias_mode == 2 || epi.sn != 1 || !epi.out || ((uintptr_t)epi.out & 15)) return false;
    const int esz = epi.out_bf16 ? 2 : 4;
    if ((epi.sm * esz) % 16) return false;
    if (epi.rowmul && ((epi.rowmul_div & 31) || !epi.out_bf16)) return false;
    if (epi.mask && !(epi.mask_bf16 && epi.mask_sn == 1 && (epi.mask_sm & 7) == 0 && ((uintptr_t)epi.mask & 15) == 0)) return false;
    return true;
}

typedef CUresult (*EncodeTiledFnP)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static zoo_status pmap(CUtensorMap *tm, const void *p, CUtensorMapDataType dt, int esz, long long d0, long long d1, long long ld, int b0, int b1) {
    static EncodeTiledFnP enc = nullptr;
    if (!enc) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) {
            set_err("cuTensorMapEncodeTiled not available from the driver");
            return ZOO_CUDA_ERROR;
        }
        enc = (EncodeTiledFnP)fn;
    }
    cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
    cuuint32_t box[2] = {(cuuint32_t)b0, (cuuint32_t)b1}, estr[2] = {1, 1};
    CUresult r = enc(tm, dt, 2, const_cast<void *>(p), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_err("tc_gemm_persist: cuTensorMapEncodeTiled failed (%d)", (int)r); return ZOO_CUDA_ERROR; }
    return ZOO_OK;
}

template <bool A_K, bool B_K>
static zoo_status launch_persist(const CUtensorMap &tmA, const CUtensorMap &tmB, const PersistParams &P, int grid, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        ZOO_CUDA(cudaFuncSetAttribute(tc_gemm_persist_kernel<A_K, B_K>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
        attr = true;
    }
    ZOO_CUDA(launch_pdl(tc_gemm_persist_kernel<A_K, B_K>, dim3(grid), dim3(TC_THREADS), (size_t)P_SMEM, st, tmA, tmB, P));
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

zoo_status tc_gemm_persist(const Operand &A, const Operand &B, int M, int N, int K, const Epilogue &epi, cudaStream_t st) {
    ZOO_REQUIRE(tc_gemm_persist_ok(A, B, M, N, K, epi), "tc_gemm_persist: unsupported shape");
    PersistParams P;
    P.M = M; P.N = N; P.K = K; P.ntn = cdiv(N, PBN); P.ntiles = cdiv(M, BM) * P.ntn; P.epi = epi;
    if (P.epi.accumulate == 2) P.epi.accumulate = 0;  // "overwrite or accumulate": never K-split here
    const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUtensorMap tmA, tmB;
    if (A.kmajor) ZOO_TRY(pmap(&tmA, A.p, bf, 2, K, M, A.ld, 64, BM));
    else ZOO_TRY(pmap(&tmA, A.p, bf, 2, M, K, A.ld, 64, 64));
    if (B.kmajor) ZOO_TRY(pmap(&tmB, B.p, bf, 2, K, N, B.ld, 64, PBN));
    else ZOO_TRY(pmap(&tmB, B.p, bf, 2, N, K, B.ld, 64, 64));
    if (epi.out_bf16) ZOO_TRY(pmap(&P.tmC, epi.out, bf, 2, N, M, epi.sm, 64, 32));
    else ZOO_TRY(pmap(&P.tmC, epi.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, N, M, epi.sm, 32, 32));
    const int grid = (int)std::min<long long>(P.ntiles, g_sm_budget);
    if (A.kmajor && B.kmajor) return launch_persist<true, true>(tmA, tmB, P, grid, st);
    if (A.kmajor && !B.kmajor) return launch_persist<true, false>(tmA, tmB, P, grid, st);
    if (!A.kmajor && B.kmajor) return launch_persist<false, true>(tmA, tmB, P, grid, st);
    return launch_persist<false, false>(tmA, tmB, P, grid, st);
}

}  // namespace zoo
