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
// EW epilogue warps (8 or 16): TMEM lane quadrant = warp % 4, column slice = (warp - 2) / 4 of EW / 4 slices of 256 / (EW / 4) columns
template <int EW> struct PersistWarps {
    static constexpr int THREADS = 64 + EW * 32;  // + TMA producer, MMA issuer
    static constexpr int PCOLS = PBN / (EW / 4);  // columns per epilogue warp (128 / 64)
    static constexpr int SLAB = 65536 / EW;       // staging bytes per warp: two 4 KB blocks (8 warps) or one (16 warps)
    static constexpr int NBUF = SLAB / 4096;
};
static constexpr int P_A_BYTES = BM * 128;    // 16 KB
static constexpr int P_RING = 3 * (P_A_BYTES + PBN * 128);  // 144 KB of operand ring: 3 x 48 KB stages (one CTA per tile) or 4 x 32 KB (CTA pair)
static constexpr int P_EPI_OFF = P_RING;                    // epilogue staging: 4 warps x 16 KB (a warp's whole 32 x 256 slab: 4 blocks of 32 rows x 128 B)
static constexpr int P_BIAS_OFF = P_EPI_OFF + 4 * 16384;     // per warp: bias [256] + rowmul [256] floats
static constexpr int P_BAR_OFF = P_BIAS_OFF + 4 * 2 * PBN * 4;
static constexpr int P_SMEM = P_BAR_OFF + 256;

struct PersistParams {
    int M, N, K;
    int ntn, ntiles;   // tiles along N, total tiles (tile = 128 * CG rows x 256 columns)
    int ablate;        // timing ablations only (ZOO_PERSIST_ABLATE; results are wrong): 1 = no TMA stores, 2 = no conversion / staging
    Epilogue epi;
    CUtensorMap tmC;   // output: box {128 bytes of columns, 32 rows}, SWIZZLE_128B
};

// ---- CTA-pair (cta_group::2) forms: two CTAs of one cluster (one TPC) share a 256 x 256 tile.  Each CTA stages its own 128 rows
// of A and HALF of the B tile (128 of the 256 columns); the leader's tcgen05.mma (M = 256) reads both halves of B out of both
// CTAs' shared memory and writes rows 0..127 into its own TMEM, rows 128..255 into the peer's.  Per SM and K-block that is
// 32 KB through the L2 -> SM path instead of 48 KB: the one-CTA tiles of this engine run AT the L2 slice throughput cap
// (512 tiles x 121 K-blocks x 48 KB = 3.04 GB in 250 us = 12.2 TB/s, B300_MICROARCH "LTS throughput cap" ~6300 B/clk).
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {  // shared::cta address -> shared::cluster address of CTA `rank`
    uint32_t d;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d) : "r"(saddr), "r"(rank));
    return d;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// both CTAs load into their own shared memory; the bytes are counted on the LEADER's barrier (bar_cluster = mapa(bar, 0))
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *tm, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {  // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((unsigned short)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// the registers of the preceding tcgen05.ld are in/out operands: nothing that reads them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]),
                   "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]),
                   "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]),
                   "+r"(r[31])
                 :
                 : "memory");
}

// One 32-row x 32-column chunk of the accumulator (lane = row, r[] = its 32 columns) -> bias / relu / relu' mask / multiplier ->
// 16-byte pieces of the lane's 128-byte row of a SWIZZLE_128B staging block.  bf16 output: the chunk is half (h) of a 64-column
// block; fp32 output: the chunk is the whole 32-column block.
template <bool OUT_BF16, bool MASK>
__device__ __forceinline__ void epi_chunk(const uint32_t (&r)[32], const uint4 (&mk)[4], const float *bias_c, const float *rm_c, int relu, uint32_t sb,
                                          int h, int lane) {
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
        const uint32_t mw[4] = {mk[j8].x, mk[j8].y, mk[j8].z, mk[j8].w};
        const float4 b0 = *reinterpret_cast<const float4 *>(bias_c + j8 * 8), b1 = *reinterpret_cast<const float4 *>(bias_c + j8 * 8 + 4);
        const float4 m0v = *reinterpret_cast<const float4 *>(rm_c + j8 * 8), m1v = *reinterpret_cast<const float4 *>(rm_c + j8 * 8 + 4);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const float mm[8] = {m0v.x, m0v.y, m0v.z, m0v.w, m1v.x, m1v.y, m1v.z, m1v.w};
        float x[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float x0 = __uint_as_float(r[j8 * 8 + 2 * i]) + bb[2 * i], x1 = __uint_as_float(r[j8 * 8 + 2 * i + 1]) + bb[2 * i + 1];
            const float r0f = fmaxf(x0, 0.f), r1f = fmaxf(x1, 0.f);
            x0 = relu ? r0f : x0;
            x1 = relu ? r1f : x1;
            if (MASK) {
                x0 = __uint_as_float(mw[i] << 16) > 0.f ? x0 : 0.f;
                x1 = __uint_as_float(mw[i] & 0xffff0000u) > 0.f ? x1 : 0.f;
            }
            x[2 * i] = x0 * mm[2 * i];
            x[2 * i + 1] = x1 * mm[2 * i + 1];
        }
        if (OUT_BF16) {  // 8 values = one 16-byte piece; piece index within the 128-byte row = h * 4 + j8
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
                pk[i] = *reinterpret_cast<const uint32_t *>(&hh);
            }
            const int jc = h * 4 + j8;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + ((uint32_t)(jc ^ (lane & 7)) << 4)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                         "r"(pk[3]) : "memory");
        } else {         // fp32: 8 values = two pieces; 32 columns fill the 128-byte row
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {
                const int jc = j8 * 2 + hc;
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + ((uint32_t)(jc ^ (lane & 7)) << 4)), "r"(__float_as_uint(x[4 * hc])),
                             "r"(__float_as_uint(x[4 * hc + 1])), "r"(__float_as_uint(x[4 * hc + 2])), "r"(__float_as_uint(x[4 * hc + 3])) : "memory");
            }
        }
    }
}

// The same chunk for the bf16 outputs without a relu' mask (phi forward with the IQN merge, fc1 forward, fc1 data gradient), on
// packed fp32 pairs: y = fma(acc, rm, bias rm) as one FFMA2 per two columns, relu folded into the bf16x2 conversion -- one
// instruction per column instead of ten (ncu: these epilogues were issue-bound, 40 % of the issue slots with 8 warps).  With
// rm = 1 (no multiplier) this is bit-identical to (acc + bias); with the IQN multiplier it needs rm >= 0 (relu(x) rm = relu(x rm)),
// which the caller asserts through Epilogue::rowmul_nonneg.
template <bool RELU>
__device__ __forceinline__ void epi_chunk_fma(const uint32_t (&r)[32], const float *brm_c, const float *rm_c, uint32_t sb, int h, int lane) {
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
        const float4 b0 = *reinterpret_cast<const float4 *>(brm_c + j8 * 8), b1 = *reinterpret_cast<const float4 *>(brm_c + j8 * 8 + 4);
        const float4 m0v = *reinterpret_cast<const float4 *>(rm_c + j8 * 8), m1v = *reinterpret_cast<const float4 *>(rm_c + j8 * 8 + 4);
        const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
        const float2 mm[4] = {make_float2(m0v.x, m0v.y), make_float2(m0v.z, m0v.w), make_float2(m1v.x, m1v.y), make_float2(m1v.z, m1v.w)};
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 y = __ffma2_rn(make_float2(__uint_as_float(r[j8 * 8 + 2 * i]), __uint_as_float(r[j8 * 8 + 2 * i + 1])), mm[i], bb[i]);
            if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(y.y), "f"(y.x));
            else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(y.y), "f"(y.x));
        }
        const int jc = h * 4 + j8;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + ((uint32_t)(jc ^ (lane & 7)) << 4)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                     "r"(pk[3]) : "memory");
    }
}

template <bool A_K, bool B_K, int CG, int EW = 8>
__global__ void __launch_bounds__(PersistWarps<EW>::THREADS, 1)
tc_gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ PersistParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint64_t *full_bar = (uint64_t *)(smem + P_BAR_OFF);
    uint64_t *empty_bar = full_bar + 4;
    uint64_t *acc_full = empty_bar + 4;     // [2]
    uint64_t *acc_empty = acc_full + 2;     // [2]
    uint32_t *tmem_slot = (uint32_t *)(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int BK = 64, MN_BOX = 64, BOX_BYTES = 64 * 128, KSTEP_MN = 16 * 128;
    constexpr int PST = CG == 2 ? 4 : 3;            // ring depth
    constexpr int BNH = PBN / CG;                   // B-tile rows (tile columns) staged by this CTA
    constexpr int P_B_BYTES = BNH * 128;
    constexpr int P_STAGE = P_A_BYTES + P_B_BYTES;  // 48 KB / 32 KB
    static_assert(PST * P_STAGE <= P_RING, "operand ring");
    const int total_kb = (P.K + BK - 1) / BK;
    const int rank = CG == 2 ? (int)cluster_rank() : 0;
    const int first_tile = blockIdx.x / CG, tile_step = gridDim.x / CG;

    if (threadIdx.x == 0) {
        for (int s = 0; s < PST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], EW * 32 * CG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CG == 2) tmem_alloc_pair(tmem_slot, 512);
        else tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anything is signalled across
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
            int it = 0;
            const uint32_t full_leader = CG == 2 ? mapa_rank(smem_u32(full_bar), 0) : 0u;
            for (int tile = first_tile; tile < P.ntiles; tile += tile_step) {
                const int m0 = (tile / P.ntn) * (BM * CG) + rank * BM, n0 = (tile % P.ntn) * PBN + rank * BNH;
                for (int kb = 0; kb < total_kb; ++kb, ++it) {
                    const int s = it % PST;
                    mbar_wait(&empty_bar[s], ((it / PST) & 1) ^ 1);
                    uint8_t *sa = smem + s * P_STAGE, *sb = sa + P_A_BYTES;
                    if (rank == 0) mbar_expect_tx(&full_bar[s], P_STAGE * CG);  // the leader's barrier counts both CTAs' bytes
                    const int k0 = kb * BK;
                    auto ld = [&](void *dst, const CUtensorMap *tm, int c0, int c1) {
                        if (CG == 2) tma_load_2d_pair(dst, tm, full_leader + 8u * s, c0, c1);
                        else tma_load_2d(dst, tm, &full_bar[s], c0, c1);
                    };
                    if (A_K) ld(sa, &tmA, k0, m0);
                    else {
#pragma unroll
                        for (int h = 0; h < BM / MN_BOX; ++h) ld(sa + h * BOX_BYTES, &tmA, m0 + h * MN_BOX, k0);
                    }
                    if (B_K) ld(sb, &tmB, k0, n0);
                    else {
#pragma unroll
                        for (int h = 0; h < BNH / MN_BOX; ++h) ld(sb + h * BOX_BYTES, &tmB, n0 + h * MN_BOX, k0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // M field (bits 24..28) = 128 * CG >> 4
            constexpr uint32_t idesc = (make_idesc(PBN, !A_K, !B_K, 1u) & ~(0x1Fu << 24)) | ((uint32_t)((BM * CG) >> 4) << 24);
            int it = 0, t = 0;
            for (int tile = first_tile; tile < P.ntiles; tile += tile_step, ++t) {
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
                        if (CG == 2) umma_bf16_pair(td, ad, bd, idesc, (kb | k) != 0);
                        else umma_bf16(td, ad, bd, idesc, (kb | k) != 0);
                    }
                    if (CG == 2) umma_commit_pair(&empty_bar[s]);
                    else umma_commit(&empty_bar[s]);
                }
                if (CG == 2) umma_commit_pair(&acc_full[ab]);
                else umma_commit(&acc_full[ab]);
            }
        }
    } else {
        // epilogue warps 2..9: TMEM lane quadrant = warp % 4 (lane = tile row), column half = (warp - 2) / 4 -- two warps per
        // scheduler, because one warp alone issues at ~0.2 IPC here (tcgen05.ld -> wait -> dependent math -> STS; ncu source page
        // in profiles/r02_ncu_persist_epilogue.txt)
        const int q = warp & 3, hh = (warp - 2) >> 2;
        const int ew = hh * 4 + q;  // private slab / bias copy of this warp
        constexpr int PHALF = PersistWarps<EW>::PCOLS, NBUF = PersistWarps<EW>::NBUF;
        constexpr bool LEAN = EW == 16;  // sixteen warps have 96 registers each: only the packed bf16 path without a relu' mask is compiled in
        const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
        const Epilogue e = P.epi;
        const int Nn = P.N, out_bf16 = e.out_bf16, relu = e.relu, bias_mode = e.bias_mode;
        const void *const maskp = LEAN ? nullptr : e.mask;
        const long long msm = e.mask_sm;
        const bool rmul = e.rowmul != nullptr;
        const bool fma_path = LEAN || (out_bf16 && !maskp && (!rmul || e.rowmul_nonneg) && !(P.ablate & 4));
        // everything below is private to the warp (its own bias / multiplier copy, its own staging slab): no CTA-level barrier in
        // the tile loop.  Per tile the warp converts its 32 rows x 128 columns into two SWIZZLE_128B blocks (fp32 output: two rounds
        // of two 32-column blocks), ONE proxy fence per round, then one TMA store per block; the stores of tile t are only waited
        // for when tile t + 1 is about to overwrite the slab.
        float *const bias_s = reinterpret_cast<float *>(smem + P_BIAS_OFF) + ew * 2 * PHALF;
        float *const rm_s = bias_s + PHALF;
        const uint32_t sbase = smem_u32(smem) + P_EPI_OFF + ew * PersistWarps<EW>::SLAB;
        int t = 0;
        int nblk = 0;  // staging blocks this warp has handed to the TMA unit so far (buffer = nblk & 1)
        // bias / multiplier of a tile are fetched one tile ahead (raw bits; converted when they are copied to shared memory), and
        // the relu' source of a 32-column chunk one chunk ahead, so that no global-load latency sits in the per-tile critical path
        // (ncu source page of the K = 64 phi GEMM: 41 % of all samples were on those loads before)
        float bv[PHALF / 32];
        uint32_t rraw[PHALF / 32];
        const unsigned short *const rmp = reinterpret_cast<const unsigned short *>(e.rowmul);
        auto fetch_cols = [&](int tile_) {
            const int n0_ = (tile_ % P.ntn) * PBN + hh * PHALF;
            const int r0_ = min((tile_ / P.ntn) * (BM * CG) + rank * BM + q * 32, P.M - 1);
            const long long rb = rmul ? (long long)(r0_ / e.rowmul_div) * e.rowmul_ld : 0;
#pragma unroll
            for (int j = 0; j < PHALF / 32; ++j) {
                const int n = n0_ + lane + 32 * j;
                bv[j] = (bias_mode == 1 && n < Nn) ? __ldg(e.bias + n) : 0.f;
                rraw[j] = (rmul && n < Nn) ? (uint32_t)__ldg(rmp + rb + n) : 0x3f80u;
            }
        };
        if (first_tile < P.ntiles) fetch_cols(first_tile);
        const uint32_t acc_empty_leader = CG == 2 ? mapa_rank(smem_u32(acc_empty), 0) : 0u;
        for (int tile = first_tile; tile < P.ntiles; tile += tile_step, ++t) {
            const int ab = t & 1;
            const int m0 = (tile / P.ntn) * (BM * CG) + rank * BM, n0 = (tile % P.ntn) * PBN + hh * PHALF;
            const int row0 = m0 + q * 32;
            const bool rowok = row0 + lane < P.M;
            const int ncols = max(0, min(PHALF, Nn - n0));   // columns of this warp's half that exist
            const bf16 *const mrow = (const bf16 *)maskp + (long long)(row0 + lane) * msm + n0;
            auto fetch_mask = [&](int c, uint4 *mk) {  // relu' source: bf16 [M][msm], same rows / columns as the output
                const bf16 *mp = mrow + c;
                const int nb = n0 + c;
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
            };
            uint4 mk[4] = {};
            if (!LEAN && maskp && ncols > 0) fetch_mask(0, mk);
            mbar_wait(&acc_full[ab], (t >> 1) & 1);
            tc_fence_after();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < PHALF / 32; ++j) {
                const float rmf = __uint_as_float(rraw[j] << 16);
                bias_s[lane + 32 * j] = fma_path ? bv[j] * rmf : bv[j];  // fma path: y = acc rm + (bias rm)
                rm_s[lane + 32 * j] = rmf;
            }
            __syncwarp();
            if (tile + tile_step < P.ntiles) fetch_cols(tile + tile_step);
            const int nchunk = (ncols + 31) >> 5;            // 32-column chunks of this warp's half
            const uint32_t tcol = tq + (uint32_t)(ab * PBN + hh * PHALF);
            // A staging block = 32 rows x 128 bytes (bf16: two 32-column chunks, fp32: one).  Each finished block is fenced, stored and
            // committed as its own bulk group, alternating between the warp's two 4 KB buffers; a buffer is only waited for (its store of
            // two blocks ago has READ it) when it is about to be overwritten, so a store has a whole block's conversion time to drain.
            // (Ablation, phi forward K = 64: with the stores waited for at the top of every tile they cost 33 us of 170, the fc1 data
            // gradient 56 us of 237 -- profiles/r02_persist_epilogue_ablation.txt.)
            auto process = [&](const uint32_t (&rc)[32], int ch) {
                const int c = ch * 32;
                const int h = out_bf16 ? (ch & 1) : 0;       // chunk within its block
                if (h == 0) {
                    if (nblk >= NBUF && lane == 0) {
                        if (NBUF == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    __syncwarp();
                }
                const uint32_t bufa = sbase + (uint32_t)(nblk & (NBUF - 1)) * 4096;
                const uint32_t sb = bufa + lane * 128;
                uint4 mkc[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) mkc[j] = mk[j];
                if (!LEAN && maskp && ch + 1 < nchunk) fetch_mask(c + 32, mk);
                if (P.ablate & 2) {
                } else if (fma_path) {
                    if (relu) epi_chunk_fma<true>(rc, bias_s + c, rm_s + c, sb, h, lane);
                    else epi_chunk_fma<false>(rc, bias_s + c, rm_s + c, sb, h, lane);
                } else if (LEAN) {
                } else if (out_bf16) {
                    if (maskp) epi_chunk<true, true>(rc, mkc, bias_s + c, rm_s + c, relu, sb, h, lane);
                    else epi_chunk<true, false>(rc, mkc, bias_s + c, rm_s + c, relu, sb, h, lane);
                } else {
                    if (maskp) epi_chunk<false, true>(rc, mkc, bias_s + c, rm_s + c, relu, sb, h, lane);
                    else epi_chunk<false, false>(rc, mkc, bias_s + c, rm_s + c, relu, sb, h, lane);
                }
                if (!out_bf16 || h == 1 || ch == nchunk - 1) {  // the block is staged: one proxy fence, one TMA store
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        if (!(P.ablate & 1))
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&P.tmC), "r"(bufa),
                                         "r"(n0 + (ch - h) * 32), "r"(row0)
                                         : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    ++nblk;
                }
            };
            // two register sets: the tcgen05.ld of chunk ch + 1 is in flight while chunk ch is converted
            uint32_t ra[32], rb[32];
            if (nchunk > 0) tmem_ld32_nowait(tcol, ra);
#pragma unroll 1
            for (int ch = 0; ch < nchunk; ch += 2) {
                tmem_ld_wait(ra);
                if (ch + 1 < nchunk) tmem_ld32_nowait(tcol + (uint32_t)(32 * (ch + 1)), rb);
                process(ra, ch);
                if (ch + 1 < nchunk) {
                    tmem_ld_wait(rb);
                    if (ch + 2 < nchunk) tmem_ld32_nowait(tcol + (uint32_t)(32 * (ch + 2)), ra);
                    process(rb, ch + 1);
                }
            }
            // every TMEM read of this tile is done (tcgen05.wait::ld above): hand the accumulator back
            tc_fence_before();
            if (CG == 2) mbar_arrive_cluster(acc_empty_leader + 8u * ab);
            else mbar_arrive(&acc_empty[ab]);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();  // nothing of the pair's shared memory / TMEM is torn down while the other half still uses it
    if (warp == 1) {
        if (CG == 2) tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

static bool persist_enabled() {
    static const bool on = [] { const char *e = getenv("ZOO_PERSIST_GEMM"); return !(e && e[0] == '0'); }();
    return on;
}

static int persist_min_k() {
    static const int k = [] { const char *e = getenv("ZOO_PERSIST_MIN_K"); return e ? atoi(e) : 64; }();
    return k;
}

// host: the shapes this engine takes (everything else stays with tc_gemm)
bool tc_gemm_persist_ok(const Operand &A, const Operand &B, int M, int N, int K, const Epilogue &epi) {
    if (!persist_enabled() || !A.is_bf16 || !B.is_bf16) return false;
    const long long tiles = (long long)cdiv(M, BM) * cdiv(N, PBN);
    // enough work to keep every SM busy without a K split.  Even the K = 64 phi GEMM (one K-block per tile, all epilogue) is faster
    // here since the epilogue runs on 8 warps with its loads prefetched: 190 us against 237 us on the one-tile-per-CTA engine
    // (it was 428 us with 4 epilogue warps and the bias / multiplier loads in the tile's critical path)
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

// CTA pairs (cta_group::2, 256 x 256 tiles) unless ZOO_PAIR_GEMM=0
static bool pair_enabled() {
    static const bool on = [] { const char *e = getenv("ZOO_PAIR_GEMM"); return !(e && e[0] == '0'); }();
    return on;
}

static int pair_min_k() {
    static const int k = [] { const char *e = getenv("ZOO_PAIR_MIN_K"); return e ? atoi(e) : 512; }();
    return k;
}

template <bool A_K, bool B_K, int CG, int EW = 8>
static zoo_status launch_persist(const CUtensorMap &tmA, const CUtensorMap &tmB, const PersistParams &P, int grid, cudaStream_t st) {
    static bool attr = false;
    auto kern = tc_gemm_persist_kernel<A_K, B_K, CG, EW>;
    constexpr int P_THREADS = PersistWarps<EW>::THREADS;
    if (!attr) {
        ZOO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
        attr = true;
    }
    if (CG == 1) {
        ZOO_CUDA(launch_pdl(kern, dim3(grid), dim3(P_THREADS), (size_t)P_SMEM, st, tmA, tmB, P));
    } else {  // a cluster of two CTAs per tile (one TPC)
        if (g_timeline_on) timeline_mark(st, 0);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(P_THREADS); cfg.dynamicSmemBytes = P_SMEM; cfg.stream = st;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = (pdl_enabled() && !g_timeline_on) ? 2 : 1;
        ZOO_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, P));
        if (g_timeline_on) timeline_mark(st, 1);
    }
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

template <int CG>
static zoo_status launch_persist_cg(bool ak, bool bk, const CUtensorMap &tmA, const CUtensorMap &tmB, const PersistParams &P, int grid, cudaStream_t st, int ew = 8) {
    if (CG == 1 && ew == 16 && ak && bk) return launch_persist<true, true, 1, 16>(tmA, tmB, P, grid, st);
    if (CG == 2 && ew == 16 && ak && !bk) return launch_persist<true, false, 2, 16>(tmA, tmB, P, grid, st);
    if (ak && bk) return launch_persist<true, true, CG>(tmA, tmB, P, grid, st);
    if (ak && !bk) return launch_persist<true, false, CG>(tmA, tmB, P, grid, st);
    if (!ak && bk) return launch_persist<false, true, CG>(tmA, tmB, P, grid, st);
    return launch_persist<false, false, CG>(tmA, tmB, P, grid, st);
}

zoo_status tc_gemm_persist(const Operand &A, const Operand &B, int M, int N, int K, const Epilogue &epi, cudaStream_t st) {
    ZOO_REQUIRE(tc_gemm_persist_ok(A, B, M, N, K, epi), "tc_gemm_persist: unsupported shape");
    // pairs pay where the main loop is long (operand traffic through L2 is the bound: fc1 forward 215 vs 255 us, weight gradient
    // 209 vs 254 us, fc1 data gradient K = 512: 225 vs 231 us); short-K, store-bound shapes are faster one CTA per tile (phi forward
    // K = 64: 150 vs 180 us)
    const int cg = (pair_enabled() && M > BM && g_sm_budget >= 2 && K >= pair_min_k()) ? 2 : 1;
    PersistParams P;
    static const int ablate = [] { const char *e = getenv("ZOO_PERSIST_ABLATE"); return e ? atoi(e) : 0; }();
    P.ablate = ablate;
    P.M = M; P.N = N; P.K = K; P.ntn = cdiv(N, PBN); P.ntiles = cdiv(M, BM * cg) * P.ntn; P.epi = epi;
    if (P.epi.accumulate == 2) P.epi.accumulate = 0;  // "overwrite or accumulate": never K-split here
    const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUtensorMap tmA, tmB;
    if (A.kmajor) ZOO_TRY(pmap(&tmA, A.p, bf, 2, K, M, A.ld, 64, BM));
    else ZOO_TRY(pmap(&tmA, A.p, bf, 2, M, K, A.ld, 64, 64));
    if (B.kmajor) ZOO_TRY(pmap(&tmB, B.p, bf, 2, K, N, B.ld, 64, PBN / cg));
    else ZOO_TRY(pmap(&tmB, B.p, bf, 2, N, K, B.ld, 64, 64));
    if (epi.out_bf16) ZOO_TRY(pmap(&P.tmC, epi.out, bf, 2, N, M, epi.sm, 64, 32));
    else ZOO_TRY(pmap(&P.tmC, epi.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, N, M, epi.sm, 32, 32));
    static const int ew_pref = [] { const char *e = getenv("ZOO_PERSIST_EW"); return e ? atoi(e) : 16; }();
    const bool lean_ok = ew_pref >= 16 && epi.out_bf16 && !epi.mask && (!epi.rowmul || epi.rowmul_nonneg) && !(ablate & 4);
    if (cg == 2) {
        const int grid = 2 * (int)std::min<long long>(P.ntiles, g_sm_budget / 2);
        // (the fc1 data gradient, K = 512: eight K-blocks per 256 x 256 tile -- its epilogue is as long as its main loop)
        // sixteen warps measured SLOWER here (239 vs 225 us): the pair's main loop already competes with the epilogue for shared-memory
        // bandwidth; ZOO_PERSIST_EW=32 forces them for the A/B
        const int ew2 = (ew_pref == 32 && lean_ok && K <= 512 && A.kmajor && !B.kmajor) ? 16 : 8;
        return launch_persist_cg<2>(A.kmajor, B.kmajor, tmA, tmB, P, grid, st, ew2);
    }
    const int grid = (int)std::min<long long>(P.ntiles, g_sm_budget);
    // short-K, all-epilogue shapes (the IQN phi forward: K = 64, 0.5 GB of bf16 output): sixteen epilogue warps, four per scheduler, so
    // that one warp's tcgen05.ld / convert / fence / bulk-store chain overlaps three others' (ZOO_PERSIST_EW=8: the eight-warp kernel)
    const int ew = (lean_ok && K <= 256 && A.kmajor && B.kmajor) ? 16 : 8;
    return launch_persist_cg<1>(A.kmajor, B.kmajor, tmA, tmB, P, grid, st, ew);
}

}  // namespace zoo
