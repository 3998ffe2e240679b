// tcgen05 GEMM engine for sm_100a.
//   D(m,n) = sum_k A(m,k) B(n,k), fp32 accumulation in TMEM, three operand modes:
//     bf16    (EB = 2)        kind::f16  on bf16 operands                          -- ZOO_PREC_BF16
//     tf32    (EB = 4)        kind::tf32 on fp32 operands, one pass                -- ZOO_PREC_TF32
//     3xtf32  (EB = 4, X3)    error-compensated: the landed fp32 tile is the hi operand (kind::tf32 ignores the low 13
//                             mantissa bits), the (otherwise idle) epilogue warps write lo = x - trunc(x) beside it, and
//                             the MMA warp issues A hi x [B hi; B lo] and A lo x B hi: fp32-grade results -- ZOO_PREC_FP32
// One 128 x BN output tile (x one K-split) per CTA:
//   warp 0      TMA producer  (cp.async.bulk.tensor 2D / 4D, SWIZZLE_128B, mbarrier ring whose depth is a launch parameter)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (cta_group::1, M=128, N=BN or 2 BN, K = 32 bytes)
//   warps 2..5  3xTF32 converters during the main loop, then the epilogue: tcgen05.ld 32x32b -> shared-memory transpose ->
//               bias / relu / relu' mask -> global (fp32 | bf16 | red add), split-K finish by the last CTA of a tile
// Operands may be K-major ([rows][K]) or MN-major ([K][rows]); the latter is what dgrad (W as [N'][K']) and
// wgrad (dY and X as [batch rows][features]) need, so no operand is ever transposed in memory.
// Contractions served: CrossCor-as-GEMM and Dense forward / dgrad / wgrad of the Nature-CNN and the IQN net
// (reference layers: src/experiments/atari/Dopamine_DQN_Atari.jl:26-35, Dopamine_IQN_Atari.jl:30-44).
#include <algorithm>
#include <mutex>

#pragma once
#include "common.cuh"

namespace zoo {

static constexpr int BM = 128;      // UMMA M (cta_group::1)
// per stage: one 128-byte swizzle span of K per row -> BK = 128 / EB elements, 4 MMAs of K = 32 / EB each
static constexpr int TC_THREADS = 192;
// Epilogue / converter warps.  Four (one per TMEM lane quadrant) everywhere, EIGHT for the fp32-grade kernels with tiles of 64
// columns or more: two warps per quadrant, each owning half of the tile's columns.  One warp alone on its scheduler issues at
// ~0.2 IPC through the converter and the epilogues (load -> wait -> dependent math -> store); a second one doubles that, halves the
// register-resident chain sums per thread (no more spills) and -- at the Dopamine batch sizes, where every CTA owns its SM for its
// whole life and the update is bound by SM residency -- shortens the part of a CTA's life that is not its main loop.
template <int BN, bool X3, int GA, bool WIDE = false>
struct TcWarps {
    // (convolution kernels only: forward / data gradient GA = 2, weight gradient GA = 4.  They own their SM anyway -- 190 KB of ring --
    // so 320 threads x up to 204 registers cost nothing; the plain-GEMM instantiations keep four warps and 168 registers, because
    // the 64-wide ones run two CTAs per SM where that pays (fc1's weight gradient).)
    // WIDE is chosen per launch: grids of at most one CTA per SM (the Dopamine batch sizes).  Many-wave grids keep four warps and
    // 168 registers so that two CTAs still share an SM (Rainbow B = 1024 fp32-grade: 2.46 ms with four warps, 2.71 ms with eight).
    static constexpr int EW = (WIDE && X3 && BN >= 64 && (GA == 2 || GA == 4)) ? 8 : 4;
    static constexpr int MIN_CTAS = EW == 8 ? 1 : 0;
    static constexpr int NT = EW * 32;          // converter / epilogue threads
    static constexpr int THREADS = 64 + NT;     // + TMA producer warp + MMA warp
    static constexpr int HC = BN / (EW / 4);    // tile columns per epilogue warp
};
static constexpr int X3_CHAIN_KB = 4;         // 3xTF32: K-blocks (of 32 floats) per TMEM accumulation chain (see the kernel)
static constexpr int TC_WS_COUNTERS = 4096;  // tail of the split-K workspace: per-tile arrival counters

// ---------------------------------------------------------------- PTX wrappers ------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address >> 4 in [0,14),
// leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version = 1 in [46,48),
// layout type SWIZZLE_128B = 2 in [61,64).
// 32-bit MN-major operands must use SWIZZLE_128B_BASE32B = 1 (32-byte swizzle atoms, pattern period 4 rows).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 at [4,6), a/b format BF16 = 1 at [7,10)/[10,13),
// a_major at 15, b_major at 16 (0 = K-major, 1 = MN-major), N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn, uint32_t fmt /*1 = BF16, 2 = TF32*/) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct TcParams {
    int M, N, K;
    int kb_per_split;  // K blocks (of 128 bytes of K) per grid.z slice
    int stages;        // depth of the shared-memory ring of this launch (<= TcCfg::MAX_STAGES)
    int xmode;         // timing ablations only (results are wrong): 1 = 3xTF32 converters skip the lo split, 2 = hi x hi MMA only,
                       // 4 = no proxy fence after the split, 8 = split computed but not stored
    Epilogue epi;
    float *ws;      // split-K workspace [M][N] (atomic accumulate, kept all-zero between launches) or nullptr
    unsigned *ctr;  // split-K per-tile arrival counters (kept all-zero between launches)
    unsigned long long *dbg;  // test hook: CTA 0 writes %globaltimer at its phase boundaries (nullptr = off)
    ConvGeom conv;  // GA kernels only: how the A operand is gathered
    // TMA-store epilogue (plain GEMMs with a bf16 row-major output): the warps stage finished 32 x 64 blocks in shared memory in
    // the SWIZZLE_128B layout and one lane hands each block to the TMA unit (cp.async.bulk.tensor ... global.shared::cta), which
    // also clips rows / columns past M / N.  tmC: 2-D map over the output, box {64 columns, 32 rows}.
    int tma_store = 0;
    int span = 128;  // bytes of K per stage row (TcCfg SPAN); 64 = the SWIZZLE_64B convolution variants
    CUtensorMap tmC;
    // tma_store == 2 (TMA data gradient): the whole tile is staged [rows][ts_span-byte column blocks] and stored with one 4-D box
    // per column block into the class's strided view of dX; one map per residue class (grid.z), tpi_max = tiles per image of the
    // largest class (grid.x = images * tpi_max; smaller classes' surplus CTAs exit)
    int ts_span = 128, tpi_max = 1;
    CUtensorMap tmC4[4];
};

extern unsigned long long *g_tc_dbg;  // set by zoo_test_gemm_trace (defined in gemm_tc.cu)

__device__ __forceinline__ void dbg_stamp(const TcParams &P, int slot) {
    if (P.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.dbg[slot] = t;
    }
}

// EB: operand element bytes (2 = bf16, 4 = fp32 consumed as tf32).  X3: 3xTF32 split (EB = 4 only).
// smem per stage: A tile 128 rows x 128 B (16 KB) + B tile BN rows x 128 B; X3 doubles it for the lo tiles.
// SPAN: bytes of K per tile row and stage = the TMA / UMMA swizzle span.  128 everywhere except the bf16 TMA-im2col convolutions
// whose innermost contiguous run is only 64 bytes (conv over a 32-channel bf16 activation tensor; the first conv's kernel row of
// 8 taps x 4 channels): those run with SWIZZLE_64B tiles, 32 K elements per stage, two tcgen05.mma per stage.
// MNT: both operands MN-major (a tile is then always 128 bytes of K x rows, whatever the span)
template <int EB, int BN, bool X3, int SPAN = 128, bool MNT = false>
struct TcCfg {
    static constexpr int BK = SPAN / EB;         // K elements per stage
    static constexpr int KSTEPS = SPAN / 32;     // tcgen05.mma per stage (K = 32 bytes each)
    static constexpr int KM_SBO = 8 * SPAN;      // K-major operand: 8-row groups this far apart
    static constexpr int KM_LAYOUT = SPAN == 128 ? 2 : 4;  // UMMA layout type: SWIZZLE_128B = 2, SWIZZLE_64B = 4
    static constexpr int UMMA_K = 32 / EB;       // K elements per tcgen05.mma
    // MN-major operands: boxes of MN_ROWS k-rows x SPAN bytes of MN (MN_BOX elements); BM / MN_BOX (BN / MN_BOX) boxes per tile
    static constexpr int MN_ROWS = 128 / EB;     // k-rows per box = K elements per stage of an MN-major operand (64 bf16, 32 fp32)
    static constexpr int MN_BOX = SPAN / EB;     // MN elements per row of a box
    static constexpr int BOX_BYTES = MN_ROWS * SPAN;
    static constexpr int KSTEP_MN = UMMA_K * SPAN;  // MN-major: one SPAN-byte row per k
    // MN-major swizzle atoms: bf16 -> 8 k-rows (SWIZZLE_128B / SWIZZLE_64B); fp32 -> 4 k-rows (SWIZZLE_128B_BASE32B, the only
    // MN-major layout tcgen05 accepts for 32-bit operands)
    static constexpr int MN_SBO = EB == 2 ? 8 * SPAN : 512;
    static constexpr int MN_LAYOUT = EB == 2 ? (SPAN == 128 ? 2 : 4) : 1;
    static constexpr int A_BYTES = BM * (MNT ? 128 : SPAN);
    static constexpr int B_BYTES = BN * (MNT ? 128 : SPAN);
    static constexpr int RAW_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGE_BYTES = RAW_BYTES * (X3 ? 2 : 1);
    // 3xTF32 stage: [A hi][B hi][B lo][A lo] -- B lo directly behind B hi, so that [B hi; B lo] is ONE operand of 2 BN rows
    static constexpr int LO_B_OFF = RAW_BYTES;             // from the stage base
    static constexpr int LO_A_OFF = RAW_BYTES + B_BYTES;
    // bf16 / tf32: two CTAs per SM (<= ~97 KB each); 3xTF32 owns an SM (TMA -> split -> MMA needs the deeper ring)
    // ring depth is chosen per launch (TcParams::stages): MAX_STAGES when the grid does not fill the GPU anyway (a lone CTA per
    // SM is TMA-latency bound, so depth is what matters), TWO_CTA_STAGES when it does (two CTAs per SM overlap instead)
    static constexpr int MAX_STAGES = (232448 - 256) / STAGE_BYTES < 6 ? (232448 - 256) / STAGE_BYTES : 6;
    static constexpr int TWO_CTA_STAGES = (113 * 1024) / STAGE_BYTES < 2 ? 2 : ((113 * 1024) / STAGE_BYTES > MAX_STAGES ? MAX_STAGES : (113 * 1024) / STAGE_BYTES);
    static constexpr int STAGES = MAX_STAGES;  // barrier arrays are sized for the deepest ring
    static constexpr int SMEM = MAX_STAGES * STAGE_BYTES + 256 /*barriers*/;
    static constexpr int smem_for(int stages) { return stages * STAGE_BYTES + 256; }
    static constexpr int ACC_COLS = X3 ? 2 * BN : BN;  // 3xTF32: columns [0, BN) = hi hi + lo hi, [BN, 2 BN) = hi lo
    static constexpr int TMEM_COLS = X3 ? 2 * ACC_COLS : (BN < 32 ? 32 : BN);  // 3xTF32: two accumulators (chains alternate)
};

template <int EB>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    if (EB == 2) umma_bf16(tmem_d, ad, bd, idesc, acc);
    else umma_tf32(tmem_d, ad, bd, idesc, acc);
}

// Implicit-GEMM gather, split in two so the (row-independent) index arithmetic is done once per K-block by 8 lanes:
//   conv_decode: K index of a 16-byte chunk -> (ky, kx, channel offset), or c < 0 past the end of K
//   conv_fetch : that chunk of one tile row (CH = 16 / EB consecutive K elements), zero outside the image
struct ConvTap { int ky, kx, c; };

__device__ __forceinline__ ConvTap conv_decode(const ConvGeom &g, int kk, int Ktot) {
    ConvTap t;
    if (kk >= Ktot) { t.ky = 0; t.kx = 0; t.c = -1; return t; }
    if (g.mode == CONV_NHWC || g.mode == CONV_DGRAD) {
        const int cw = g.mode == CONV_NHWC ? g.C : g.Cout;
        const int tap = kk / cw;
        t.c = kk - tap * cw;
        t.ky = tap / g.ks;
        t.kx = tap - t.ky * g.ks;
    } else {  // raw observation: K order (c, ky, kx); ks % CH == 0 keeps a chunk inside one kernel row
        const int kk2 = g.ks * g.ks;
        t.c = kk / kk2;
        const int rem = kk - t.c * kk2;
        t.ky = rem / g.ks;
        t.kx = rem - t.ky * g.ks;
    }
    return t;
}

template <int EB>
__device__ __forceinline__ uint4 conv_fetch(const ConvGeom &g, bool valid, int b, int y0, int x0, const ConvTap t) {
    constexpr int CH = 16 / EB;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (!valid || t.c < 0) return val;
    if (g.mode == CONV_NHWC) {
        const int iy = y0 + t.ky, ix = x0 + t.kx;
        if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
            val = __ldg(reinterpret_cast<const uint4 *>((const uint8_t *)g.x + ((((long long)b * g.H + iy) * g.W + ix) * g.C + t.c) * EB));
    } else if (g.mode == CONV_DGRAD) {
        const int ty = y0 + g.p - t.ky, tx = x0 + g.p - t.kx;  // y0, x0 = input pixel
        if (ty >= 0 && tx >= 0) {
            const int oy = g.s == 1 ? ty : (g.s == 2 ? ty >> 1 : ty / g.s), ox = g.s == 1 ? tx : (g.s == 2 ? tx >> 1 : tx / g.s);
            if (oy * g.s == ty && ox * g.s == tx && oy < g.OH && ox < g.OW)
                val = __ldg(reinterpret_cast<const uint4 *>((const uint8_t *)g.x + ((((long long)b * g.OH + oy) * g.OW + ox) * g.Cout + t.c) * EB));
        }
    } else {
        const int iy = y0 + t.ky;
        if (iy >= 0 && iy < g.H) {
            const long long rowoff = (((long long)b * g.C + t.c) * g.H + iy) * g.W;
            float f[CH];
#pragma unroll
            for (int e = 0; e < CH; ++e) {
                const int ix = x0 + t.kx + e;
                float v = 0.f;
                if (ix >= 0 && ix < g.W) {
                    v = g.mode == CONV_RAW_U8 ? (float)__ldg((const uint8_t *)g.x + rowoff + ix) : __ldg((const float *)g.x + rowoff + ix);
                    if (g.scale255) v = __fdiv_rn(v, 255.f);
                }
                f[e] = v;
            }
            if (EB == 4) {
                val = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
            } else {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1 % CH]), h1 = __floats2bfloat162_rn(f[2 % CH], f[3 % CH]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4 % CH], f[5 % CH]), h3 = __floats2bfloat162_rn(f[6 % CH], f[7 % CH]);
                val = make_uint4(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1), *reinterpret_cast<uint32_t *>(&h2),
                                 *reinterpret_cast<uint32_t *>(&h3));
            }
        }
    }
    return val;
}

// global address of a chunk for the cp.async path (NHWC / DGRAD modes), nullptr = zero fill
template <int EB>
__device__ __forceinline__ const void *conv_src(const ConvGeom &g, bool valid, int b, int y0, int x0, const ConvTap t) {
    if (!valid || t.c < 0) return nullptr;
    if (g.mode == CONV_NHWC) {
        const int iy = y0 + t.ky, ix = x0 + t.kx;
        if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) return (const uint8_t *)g.x + ((((long long)b * g.H + iy) * g.W + ix) * g.C + t.c) * EB;
        return nullptr;
    }
    const int ty = y0 + g.p - t.ky, tx = x0 + g.p - t.kx;  // DGRAD: y0, x0 = input pixel
    if (ty < 0 || tx < 0) return nullptr;
    const int oy = g.s == 1 ? ty : (g.s == 2 ? ty >> 1 : ty / g.s), ox = g.s == 1 ? tx : (g.s == 2 ? tx >> 1 : tx / g.s);
    if (oy * g.s != ty || ox * g.s != tx || oy >= g.OH || ox >= g.OW) return nullptr;
    return (const uint8_t *)g.x + ((((long long)b * g.OH + oy) * g.OW + ox) * g.Cout + t.c) * EB;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, const void *fallback) {
    const int n = src ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src ? src : fallback), "r"(n) : "memory");
}

// all 8 chunks of this thread's tile row for the K-block starting at k0 (lanes j and j+8.. decode chunk j & 7 and share it)
template <int EB>
__device__ __forceinline__ void conv_gather_row(const ConvGeom &g, bool valid, int b, int y0, int x0, int k0, int Ktot, int lane, uint4 (&v)[8]) {
    const ConvTap mine = conv_decode(g, k0 + (lane & 7) * (16 / EB), Ktot);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        ConvTap t;
        t.ky = __shfl_sync(0xffffffffu, mine.ky, j);
        t.kx = __shfl_sync(0xffffffffu, mine.kx, j);
        t.c = __shfl_sync(0xffffffffu, mine.c, j);
        v[j] = conv_fetch<EB>(g, valid, b, y0, x0, t);
    }
}

__device__ __forceinline__ uint4 tf32_lo(const uint4 v) {
    uint4 l;
    l.x = __float_as_uint(__fsub_rn(__uint_as_float(v.x), __uint_as_float(v.x & 0xffffe000u)));
    l.y = __float_as_uint(__fsub_rn(__uint_as_float(v.y), __uint_as_float(v.y & 0xffffe000u)));
    l.z = __float_as_uint(__fsub_rn(__uint_as_float(v.z), __uint_as_float(v.z & 0xffffe000u)));
    l.w = __float_as_uint(__fsub_rn(__uint_as_float(v.w), __uint_as_float(v.w & 0xffffe000u)));
    return l;
}

// Epilogue row walkers.  One warp owns a 32 x 32 chunk staged in shared memory; lane = column, LD = row stride of the
// staging buffer.  A single warp per scheduler runs this, so the cost is instruction count x dependent-issue latency:
// 8 rows per batch (independent chains), 32-bit strides, output type resolved outside the loop.
template <typename T>
__device__ __forceinline__ void epi_put(T *p, float x);
template <>
__device__ __forceinline__ void epi_put<float>(float *p, float x) { *p = x; }
template <>
__device__ __forceinline__ void epi_put<bf16>(bf16 *p, float x) { *p = __float2bfloat16_rn(x); }

// (values that originate in the kernel parameter bank are laundered through empty asm statements: otherwise the compiler
// re-reads them from the constant bank next to every store instead of keeping them in registers.)
template <typename P>
__device__ __forceinline__ P *in_reg(P *p) { asm volatile("" : "+l"(p)); return p; }
__device__ __forceinline__ int in_reg(int v) { asm volatile("" : "+r"(v)); return v; }

// rm: per-column multiplier applied after bias / relu (IQN: the Hadamard merge psi[b] .* relu(phi), iqn.jl:28-30; 1 otherwise)
template <typename T, int LD, bool MASK>
__device__ __forceinline__ void epi_rows_store(const float *cp, int rows, T *op, int stride, float bv, int relu, unsigned mbits, float rm = 1.f) {
    op = in_reg(op); stride = in_reg(stride); relu = in_reg(relu);
#pragma unroll 1
    for (int r0 = 0; r0 < rows; r0 += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = cp[(r0 + u) * LD];
        T *o = op + (long long)r0 * stride;
        const unsigned mb = mbits >> r0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float x = v[u] + bv;
            const float xr = fmaxf(x, 0.f);
            x = relu ? xr : x;
            if (MASK) x = (mb & (1u << u)) ? x : 0.f;
            v[u] = x * rm;
        }
        if (r0 + 8 <= rows) {  // whole batch in range: straight-line stores, no per-row branch
#pragma unroll
            for (int u = 0; u < 8; ++u) epi_put<T>(o + u * stride, v[u]);
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (r0 + u < rows) epi_put<T>(o + u * stride, v[u]);
        }
    }
}
template <int LD>
__device__ __forceinline__ void epi_rows_red(const float *cp, int rows, float *op, int stride) {
    op = in_reg(op); stride = in_reg(stride);
#pragma unroll 1
    for (int r0 = 0; r0 < rows; r0 += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = cp[(r0 + u) * LD];
        float *o = op + (long long)r0 * stride;
        if (r0 + 8 <= rows) {
#pragma unroll
            for (int u = 0; u < 8; ++u) atomicAdd(o + u * stride, v[u]);
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (r0 + u < rows) atomicAdd(o + u * stride, v[u]);
        }
    }
}
// the same with 16-byte reds (red.global.add.v4.f32): eight lanes cover the 32 columns of a row, the warp four rows per instruction --
// a quarter of the red instructions.  The per-SM atomic issue rate, not L2, bounds these epilogues: the weight-gradient and split-K
// CTAs spent 5 - 6 us issuing 8 192 scalar reds each.  cp = the chunk's element (0, 0); op = its output element (0, 0), 16-byte aligned.
template <int LD>
__device__ __forceinline__ void epi_rows_red_v4(const float *cp, int rows, float *op, int stride, int lane) {
    op = in_reg(op); stride = in_reg(stride);
    const int rr = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll 1
    for (int r0 = 0; r0 < rows; r0 += 8) {
        float v[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int e = 0; e < 4; ++e) v[u][e] = cp[(r0 + 4 * u + rr) * LD + c4 + e];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int r = r0 + 4 * u + rr;
            if (r < rows)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op + (long long)r * stride + c4), "f"(v[u][0]), "f"(v[u][1]), "f"(v[u][2]),
                             "f"(v[u][3])
                             : "memory");
        }
    }
}
// packed bf16 rows: lane = column pair, staging row stride 66 floats; mask (relu') rows are bf16x2 as well
template <bool MASK>
__device__ __forceinline__ void epi_rows_store_bf16x2(const float *cp, int rows, __nv_bfloat162 *op, int stride2, float b0, float b1, int relu,
                                                      const __nv_bfloat162 *mp, int mstride2, float rm0 = 1.f, float rm1 = 1.f) {
    op = in_reg(op); stride2 = in_reg(stride2); relu = in_reg(relu);
    if (MASK) { mp = in_reg(mp); mstride2 = in_reg(mstride2); }
#pragma unroll 1
    for (int r0 = 0; r0 < rows; r0 += 8) {
        float2 v[8];
        __nv_bfloat162 mk[8];
        const bool full = r0 + 8 <= rows;
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float2 *>(cp + (r0 + u) * 66);
        if (MASK) {
            const __nv_bfloat162 *m = mp + (long long)r0 * mstride2;
#pragma unroll
            for (int u = 0; u < 8; ++u) mk[u] = __ldg(m + (full ? u : min(u, rows - 1 - r0)) * mstride2);
        }
        __nv_bfloat162 *o = op + (long long)r0 * stride2;
        __nv_bfloat162 w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float x0 = v[u].x + b0, x1 = v[u].y + b1;
            const float r0f = fmaxf(x0, 0.f), r1f = fmaxf(x1, 0.f);
            x0 = relu ? r0f : x0;
            x1 = relu ? r1f : x1;
            if (MASK) {
                x0 = __low2float(mk[u]) > 0.f ? x0 : 0.f;
                x1 = __high2float(mk[u]) > 0.f ? x1 : 0.f;
            }
            w[u] = __floats2bfloat162_rn(x0 * rm0, x1 * rm1);
        }
        if (full) {
#pragma unroll
            for (int u = 0; u < 8; ++u) o[u * stride2] = w[u];
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (r0 + u < rows) o[u * stride2] = w[u];
        }
    }
}

// transposed chunk (swap-AB): lane = tile row, walk the columns; the output index advances by `stride` per column
template <typename T, bool MASK>
__device__ __forceinline__ void epi_cols_store(const float *rp, int ncols, T *op, int stride, float bm, int relu, unsigned mbits) {
    op = in_reg(op); stride = in_reg(stride); relu = in_reg(relu);
#pragma unroll 1
    for (int c0 = 0; c0 < ncols; c0 += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = rp[c0 + u];
        T *o = op + (long long)c0 * stride;
        const unsigned mb = mbits >> c0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float x = v[u] + bm;
            const float xr = fmaxf(x, 0.f);
            x = relu ? xr : x;
            if (MASK) x = (mb & (1u << u)) ? x : 0.f;
            v[u] = x;
        }
        if (c0 + 8 <= ncols) {
#pragma unroll
            for (int u = 0; u < 8; ++u) epi_put<T>(o + u * stride, v[u]);
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (c0 + u < ncols) epi_put<T>(o + u * stride, v[u]);
        }
    }
}

// GA = 1: the A operand is gathered by warps 2..5 (implicit-GEMM convolution, see ConvGeom) instead of TMA-loaded.
// GA = 2: implicit-GEMM convolution with the im2col rows fetched by TMA itself: one M tile = R whole output rows of one
//         image, one K-block = one kernel tap (or kernel row), loaded as a single 4-D box of the NHWC activation tensor
//         (element strides = conv stride, out-of-bounds = zero padding).
// GA = 4: implicit-GEMM convolution WEIGHT gradient, dW[o][(ky,kx,c)] = sum over pixels of dY[pixel][o] * X[pixel @ tap][c]:
//         M = Cout (one tile, rows past Cout stay zero), N = the K x K x C features, contraction over the output pixels.  Both
//         operands are MN-major and arrive as TMA boxes of R whole output rows of one image: dY through a 3-D map
//         {Cout, pixels, images}, the activations through the same 4-D strided boxes as the forward conv (one per tap / kernel
//         row) -- so no im2col matrix exists in either direction.  A K-block's pixel rows past R x OW are never written
//         (the ring is zeroed once at kernel start), grid.z splits the images, partial sums are red-added into the gradient.
template <int EB, int BN, bool A_K, bool B_K, bool X3, int GA = 0, int SPAN = 128, bool WIDE = false>
#ifndef ZOO_BF16_CONV_MIN_CTAS
#define ZOO_BF16_CONV_MIN_CTAS 5  // measured (Rainbow B = 1024 bf16): 0 -> 0.811 ms, 3 -> 0.770, 4 -> 0.707, 5 -> 0.684 ms per update (conv1 forward 108 -> 61 us)
#endif
// bf16 TMA-im2col convolutions with narrow tiles (BN <= 64) are many-wave, fixed-cost-bound grids at the large batch sizes
// (Rainbow B = 1024 conv1: 4 096 tiles of ~4 us of prologue / epilogue around a 0.5 us main loop): what helps is more resident CTAs
// per SM, which the 168 registers of the unconstrained build cap at two.
__global__ void __launch_bounds__((TcWarps<BN, X3, GA, WIDE>::THREADS), ((EB == 2 && !X3 && BN <= 64 && GA == 2) ? ZOO_BF16_CONV_MIN_CTAS : TcWarps<BN, X3, GA, WIDE>::MIN_CTAS))
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ TcParams P) {
    static_assert(GA == 0 || GA == 4 || A_K, "the gathered A operand is K-major");
    static_assert(GA != 4 || (!A_K && !B_K), "the weight-gradient operands are MN-major");
    static_assert(SPAN == 128 || (SPAN == 64 && EB == 2 && !X3 && ((GA == 2 && A_K && B_K) || GA == 4)), "64-byte spans: bf16 TMA-im2col convolutions only");
    using C = TcCfg<EB, BN, X3, SPAN, GA == 4>;
    using TW = TcWarps<BN, X3, GA, WIDE>;
    constexpr int NT = TW::NT, HC = TW::HC;
    // SWIZZLE_128B tiles need 1024-byte alignment; declaring it on the array (instead of rounding the pointer up) keeps
    // the shared address space visible to the compiler, so the epilogue staging below compiles to LDS/STS
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    const int NST = GA == 1 ? C::STAGES : P.stages;  // ring depth of this launch
    uint64_t *full_bar = (uint64_t *)(smem + NST * C::STAGE_BYTES);
    uint64_t *empty_bar = full_bar + C::STAGES;
    uint64_t *conv_bar = empty_bar + C::STAGES;  // X3: lo tiles written (128 converter threads arrive)
    uint64_t *acc_full = conv_bar + C::STAGES;   // [2] accumulator buffer b holds a finished chain (tcgen05.commit)
    uint64_t *acc_empty = acc_full + 2;          // [2] buffer b has been drained to registers (128 threads arrive)
    uint32_t *tmem_slot = (uint32_t *)(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int m0 = blockIdx.x * BM, vrows = min(BM, P.M - (int)blockIdx.x * BM);  // first output row / valid rows of this M tile
    int cimg = 0, coy0 = 0;
    // TMA data gradient (GA = 2, CONV_TMA_DGRAD): geometry of this CTA's residue class (see common.cuh)
    const bool dgrad = GA == 2 && P.conv.mode == CONV_TMA_DGRAD;
    int cls_py = 0, cls_px = 0, cols_c = 0, cls_R = 0, ky0 = 0, kx0 = 0, kt_x = 1, dy0 = 0, dx0 = 0, cls_kb = 0;
    if (GA == 2 && !dgrad) {
        cimg = blockIdx.x / P.conv.tpi;
        coy0 = (blockIdx.x - cimg * P.conv.tpi) * P.conv.R;
        m0 = (cimg * P.conv.OH + coy0) * P.conv.OW;
        vrows = min(P.conv.R, P.conv.OH - coy0) * P.conv.OW;
    }
    if (dgrad) {
        const ConvGeom &g = P.conv;
        cls_py = blockIdx.z / g.s; cls_px = blockIdx.z - cls_py * g.s;
        // every class is tiled like the largest one (ceil(H/s) x ceil(W/s) pixels, the A boxes are the same shape for all):
        // pixels a smaller class does not have are computed and then clipped by its output map
        const int rows_c = (g.H + g.s - 1) / g.s;
        cols_c = (g.W + g.s - 1) / g.s;
        cls_R = min(rows_c, BM / cols_c);
        cimg = blockIdx.x / P.tpi_max;
        coy0 = (blockIdx.x - cimg * P.tpi_max) * cls_R;
        vrows = min(cls_R, rows_c - coy0) * cols_c;
        m0 = 0;
        ky0 = (cls_py + g.p) % g.s; kx0 = (cls_px + g.p) % g.s;
        const int kt_y = (g.ks - ky0 + g.s - 1) / g.s;
        kt_x = (g.ks - kx0 + g.s - 1) / g.s;
        dy0 = (cls_py + g.p - ky0) / g.s; dx0 = (cls_px + g.p - kx0) / g.s;
        cls_kb = kt_y * kt_x * g.cblocks;
    }
    const int n0 = blockIdx.y * BN;
    if (threadIdx.x == 0) dbg_stamp(P, 0);
    const int total_kb = GA == 4 ? P.K : (dgrad ? cls_kb : (P.K + C::BK - 1) / C::BK);  // GA = 4: P.K counts K-blocks (images x row groups)
    const int kb0 = dgrad ? 0 : blockIdx.z * P.kb_per_split;
    const int kb1 = dgrad ? total_kb : min(total_kb, kb0 + P.kb_per_split);
    const int nkb = kb1 - kb0;  // host guarantees >= 1
    // 3xTF32: the K loop is cut into chains of CHAIN K-blocks that alternate between two TMEM accumulators; the
    // converter warps sum finished chains in registers with round-to-nearest fp32 adds, because a TMEM accumulator
    // truncates on every add and that bias grows with the chain length.  Other modes: one chain, one accumulator.
    constexpr int CHAIN = X3 ? X3_CHAIN_KB : (1 << 30);
    const int nchains = (nkb + CHAIN - 1) / CHAIN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&conv_bar[s], NT);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], NT);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (GA == 4) {  // pad rows of every stage (pixel rows past R x OW, feature boxes past N, channel rows past Cout) must read as zero
        for (int i = threadIdx.x; i < NST * C::STAGE_BYTES / 16; i += TW::THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) dbg_stamp(P, 1);
    pdl_wait();  // everything above ran under the predecessor's tail; from here on global memory is touched

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
            for (int i = 0; i < nkb; ++i) {
                const int s = i % NST;
                const uint32_t ph = (i / NST) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t *sa = smem + s * C::STAGE_BYTES;
                uint8_t *sb = sa + C::A_BYTES;
                if (GA == 4) {
                    const ConvGeom &g = P.conv;
                    const int kb = kb0 + i;
                    const int img = kb / g.tpi, oy0 = (kb - img * g.tpi) * g.R;
                    const int abox = (P.M + C::MN_BOX - 1) / C::MN_BOX;                       // Cout blocks of dY
                    const int bbox = min(BN / C::MN_BOX, (P.N - n0 + C::MN_BOX - 1) / C::MN_BOX);  // feature blocks of this N tile
                    const int box_bytes = g.R * g.OW * SPAN;                                   // R x OW pixel rows of SPAN bytes
                    mbar_expect_tx(&full_bar[s], (abox + bbox) * box_bytes);
                    for (int h = 0; h < abox; ++h)
                        asm volatile(
                            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                            ::"r"(smem_u32(sa + h * C::BOX_BYTES)), "l"(&tmA), "r"(smem_u32(&full_bar[s])), "r"(h * C::MN_BOX), "r"(oy0 * g.OW), "r"(img)
                            : "memory");
                    for (int j = 0; j < bbox; ++j) {
                        const int f = n0 + j * C::MN_BOX;  // first feature of the box: (ky, kx, c) with c fastest
                        int c0, c1, c2;
                        if (g.first) { c0 = 0; c1 = 0; c2 = oy0 * g.s + f / (g.ks * g.C); }  // one kernel row of the padded frame stack
                        else { const int tap = f / g.C, ky = tap / g.ks, kx = tap - ky * g.ks; c0 = f - tap * g.C; c1 = kx - g.p; c2 = oy0 * g.s + ky - g.p; }
                        asm volatile(
                            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                            ::"r"(smem_u32(sb + j * C::BOX_BYTES)), "l"(&tmB), "r"(smem_u32(&full_bar[s])), "r"(c0), "r"(c1), "r"(c2), "r"(img)
                            : "memory");
                    }
                    if (i == 0) dbg_stamp(P, 2);
                    continue;
                }
                mbar_expect_tx(&full_bar[s], GA == 1 ? C::B_BYTES : (GA == 2 ? (dgrad ? cls_R * cols_c : P.conv.R * P.conv.OW) * SPAN + C::B_BYTES : C::RAW_BYTES));
                const int k0 = (kb0 + i) * C::BK;
                if (GA == 2) {
                    const ConvGeom &g = P.conv;
                    const int kb = kb0 + i;
                    int c0, c1, c2;
                    if (g.first) { c0 = 0; c1 = 0; c2 = coy0 * g.s + kb; }  // K-block = kernel row ky of the padded frame stack
                    else if (g.mode == CONV_TMA_DGRAD) {  // data gradient: a stride-1 conv of dY with the class's (mirrored) sub-kernel
                        const int t = kb / g.cblocks, a = t / kt_x, b = t - a * kt_x;
                        c0 = (kb - t * g.cblocks) * C::BK; c1 = dx0 - b; c2 = coy0 + dy0 - a;
                    } else {
                        const int tap = kb / g.cblocks, ky = tap / g.ks, kx = tap - ky * g.ks;
                        c0 = (kb - tap * g.cblocks) * C::BK; c1 = kx - g.p; c2 = coy0 * g.s + ky - g.p;
                    }
                    asm volatile(
                        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                        ::"r"(smem_u32(sa)), "l"(&tmA), "r"(smem_u32(&full_bar[s])), "r"(c0), "r"(c1), "r"(c2), "r"(cimg)
                        : "memory");
                }
                if (GA == 0) {
                    if (A_K) {
                        tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);  // box {BK k, 128 rows}
                    } else {
#pragma unroll
                        for (int h = 0; h < BM / C::MN_BOX; ++h)  // boxes {MN_BOX m, BK k}
                            tma_load_2d(sa + h * C::BOX_BYTES, &tmA, &full_bar[s], m0 + h * C::MN_BOX, k0);
                    }
                }
                if (B_K) {
                    tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);  // box {BK k, BN rows}
                } else {
                    // conv dgrad: K runs over (tap, o); tap t's [Cout][C] slice of W sits at column offset t*C of [Cout][ks*ks*C]
                    int nb = n0, kb = k0;
                    if (GA != 0 && !dgrad) { const int tap = k0 / P.conv.Cout; nb = n0 + tap * P.conv.C; kb = k0 - tap * P.conv.Cout; }
                    if (dgrad) {  // class tap (a, b) -> kernel tap (ky0 + a s, kx0 + b s)
                        const int kbi = kb0 + i, t = kbi / P.conv.cblocks, a = t / kt_x, b = t - a * kt_x;
                        const int tap = (ky0 + a * P.conv.s) * P.conv.ks + kx0 + b * P.conv.s;
                        nb = n0 + tap * P.conv.C; kb = (kbi - t * P.conv.cblocks) * C::BK;
                    }
#pragma unroll
                    for (int h = 0; h < BN / C::MN_BOX; ++h)
                        tma_load_2d(sb + h * C::BOX_BYTES, &tmB, &full_bar[s], nb + h * C::MN_BOX, kb);
                }
                if (i == 0) dbg_stamp(P, 2);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN, !A_K, !B_K, EB == 2 ? 1u : 2u);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % NST;
                const uint32_t ph = (i / NST) & 1;
                const int chain = i / CHAIN, cb = chain & 1;
                const bool chain_start = (i % CHAIN) == 0;
                if (X3 && chain_start && chain >= 2) {  // the buffer's previous chain must have been drained
                    mbar_wait(&acc_empty[cb], ((chain >> 1) - 1) & 1);
                    tc_fence_after();
                }
                if (X3 || GA == 1) mbar_wait(&conv_bar[s], ph);  // converter / gather warps are done with the stage
                if (!X3) mbar_wait(&full_bar[s], ph);       // TMA bytes have landed (the 3xTF32 converters waited for them)
                tc_fence_after();
                if (i == 0) dbg_stamp(P, 3);
                if (i == 1) dbg_stamp(P, 12);
                if (i == 2) dbg_stamp(P, 13);
                if (i == nkb - 1) dbg_stamp(P, 14);
                const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
                const uint32_t sb = sa + C::A_BYTES;
                const uint32_t td = tmem_base + (uint32_t)(cb * C::ACC_COLS);
                const int ksteps = GA == 4 ? P.conv.ksteps : C::KSTEPS;  // GA = 4: ceil(R x OW / UMMA_K) (pad rows are zero)
#pragma unroll
                for (int k = 0; k < (GA == 4 ? 4 : C::KSTEPS); ++k) {
                    if (GA == 4 && k >= ksteps) break;
                    // K-major SW128: 8-row groups 1024 B apart (SBO); one MMA's K slice = 32 B inside the swizzle span.
                    // MN-major SW128: MN blocks of 128 B BOX_BYTES apart (LBO), 8-k groups 1024 B apart (SBO).
                    const uint32_t ao = A_K ? k * 32 : k * C::KSTEP_MN;
                    const uint32_t bo = B_K ? k * 32 : k * C::KSTEP_MN;
                    const uint64_t ad = A_K ? make_smem_desc(sa + ao, 16, C::KM_SBO, C::KM_LAYOUT) : make_smem_desc(sa + ao, C::BOX_BYTES, C::MN_SBO, C::MN_LAYOUT);
                    const uint64_t bd = B_K ? make_smem_desc(sb + bo, 16, C::KM_SBO, C::KM_LAYOUT) : make_smem_desc(sb + bo, C::BOX_BYTES, C::MN_SBO, C::MN_LAYOUT);
                    if (X3) {
                        // every tcgen05.mma costs about the same ~130 clocks for N <= 128 (measured), so the three products are
                        // issued as TWO instructions: A hi x [B hi; B lo] (one operand of 2 BN rows -> accumulator columns
                        // [0, BN) and [BN, 2 BN)), then A lo x B hi into [0, BN); the halves are summed in registers.
                        constexpr uint32_t idesc2 = make_idesc(2 * BN, !A_K, !B_K, 2u);
                        const uint64_t adl = A_K ? make_smem_desc(sa + C::LO_A_OFF + ao, 16, 1024)
                                                 : make_smem_desc(sa + C::LO_A_OFF + ao, C::BOX_BYTES, C::MN_SBO, C::MN_LAYOUT);
                        if (P.xmode & 2) {
                            umma<EB>(td, ad, bd, idesc, !(chain_start && k == 0));
                        } else {
                            umma<EB>(td, ad, bd, idesc2, !(chain_start && k == 0));
                            umma<EB>(td, adl, bd, idesc, 1);
                        }
                    } else {
                        umma<EB>(td, ad, bd, idesc, (i | k) != 0);
                    }
                }
                umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
                if ((i % CHAIN) == CHAIN - 1 || i == nkb - 1) umma_commit(&acc_full[cb]);  // chain complete
            }
            dbg_stamp(P, 4);
        }
    } else {
        // warps 2..5 (2..9): TMEM lane quadrant = warp % 4, thread = accumulator row; with eight warps, column half = (warp - 2) / 4.
        const int q = warp & 3;
        const int c0 = TW::EW == 8 ? ((warp - 2) >> 2) * HC : 0;  // first tile column of this warp
        const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
        if (P.epi.mask && GA != 4) {
            // The relu' mask of the epilogue is an activation tensor the forward pass wrote ~100 us earlier; by now the optimiser's
            // and the weight gradients' streams have pushed it out of L2, and the epilogue's loads paid a DRAM round trip under load
            // (measured 2.5 - 3 us per 128-byte line of a row).  Each thread asks L2 for its share of the tile's mask now.
            const int t = threadIdx.x - 64;
            const int mesz = P.epi.mask_bf16 ? 2 : 4;
            const uint8_t *mbase = (const uint8_t *)P.epi.mask;
            if (GA == 2 && P.tma_store == 2) {
                if (t < vrows) {
                    const int ry = t / cols_c, rx = t - ry * cols_c;
                    const int y = (coy0 + ry) * P.conv.s + cls_py, xx = rx * P.conv.s + cls_px;
                    if (y < P.conv.H && xx < P.conv.W) {
                        const uint8_t *row = mbase + ((((long long)cimg * P.conv.H + y) * P.conv.W + xx) * P.epi.mask_sm + n0) * mesz;
                        const int bytes = min(BN, P.N - n0) * mesz;
                        for (int b = 0; b < bytes; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + b));
                    }
                }
            } else if (GA == 0) {
                if (P.epi.mask_sn == 1) {  // rows of the mask run along n
                    if (t < vrows) {
                        const uint8_t *row = mbase + ((long long)(m0 + t) * P.epi.mask_sm + n0) * mesz;
                        const int bytes = min(BN, P.N - n0) * mesz;
                        for (int b = 0; b < bytes; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + b));
                    }
                } else if (P.epi.mask_sm == 1) {  // swap-AB outputs: the mask runs along m
                    const int segs = (BM * mesz) / 128;  // 128-byte pieces per column of the tile
                    for (int j = t; j < BN * segs; j += NT) {
                        const int n = n0 + j / segs, mm = m0 + (j % segs) * (128 / mesz);
                        if (n < P.N && mm < P.M) asm volatile("prefetch.global.L2 [%0];" ::"l"(mbase + ((long long)n * P.epi.mask_sn + mm) * mesz));
                    }
                }
            }
        }
        float acc[X3 ? HC : 1];  // running fp32 sum of the drained chains (3xTF32 only): this warp's columns [c0, c0 + HC)
        if (X3) {
#pragma unroll
            for (int j = 0; j < HC; ++j) acc[j] = 0.f;
        }
        if (X3 || GA == 1) {
            // per K-block: (GA) gather this thread's A row into the stage -- hi and, for 3xTF32, lo -- and (3xTF32) split the
            // TMA-landed fp32 tiles into hi (the tile itself) and lo; one chain behind the MMA warp, fold finished chains
            // into acc
            const int t = threadIdx.x - 64;  // 0..NT-1
            int drained = 0;
            const int va = vrows, vb = min(BN, P.N - n0);  // valid rows of the A / B tiles
            const int a_units = GA == 1 ? 0 : (A_K ? va * 128 : ((va + C::MN_BOX - 1) / C::MN_BOX) * C::BOX_BYTES) / 16;
            const int b_units = (B_K ? vb * 128 : ((vb + C::MN_BOX - 1) / C::MN_BOX) * C::BOX_BYTES) / 16;
            const int conv_units = a_units + b_units;
            // GA: this thread owns tile row t = output pixel (forward) / input pixel (dgrad) m0 + t
            bool gvalid = false;
            int gb = 0, gy0 = 0, gx0 = 0;
            constexpr int GA_AHEAD = C::STAGES - 2;  // K-blocks of cp.async gathers in flight ahead of the one being finished
            if (GA == 1) {
                const ConvGeom &g = P.conv;
                const int m = m0 + t;
                gvalid = m < P.M;
                const int gw = g.mode == CONV_DGRAD ? g.W : g.OW, gh = g.mode == CONV_DGRAD ? g.H : g.OH;
                const int px = m % gw, tmp = m / gw;
                const int py = tmp % gh;
                gb = tmp / gh;
                gy0 = g.mode == CONV_DGRAD ? py : py * g.s - g.p;
                gx0 = g.mode == CONV_DGRAD ? px : px * g.s - g.p;
            }
            for (int i = 0; i < nkb; ++i) {
                const int s = i % NST;
                const uint32_t ph = (i / NST) & 1;
                uint8_t *stage = smem + s * C::STAGE_BYTES;
                if (GA == 1) {
                    if (P.conv.mode == CONV_NHWC || P.conv.mode == CONV_DGRAD) {
                        // cp.async pipeline: the 16-byte chunks of K-block i+GA_AHEAD go straight from global memory into their
                        // swizzled slots (zero-filled outside the image) while block i, issued GA_AHEAD iterations ago, lands
                        for (int ib = (i == 0 ? 0 : i + GA_AHEAD); ib <= i + GA_AHEAD; ++ib) {
                            if (ib < nkb) {
                                const int sI = ib % C::STAGES;
                                mbar_wait(&empty_bar[sI], ((ib / C::STAGES) & 1) ^ 1);  // MMAs on the stage's old contents retired
                                const uint32_t sbase = smem_u32(smem + sI * C::STAGE_BYTES) + t * 128;
                                const ConvTap mine = conv_decode(P.conv, (kb0 + ib) * C::BK + (lane & 7) * (16 / EB), P.K);
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    ConvTap tp;
                                    tp.ky = __shfl_sync(0xffffffffu, mine.ky, j);
                                    tp.kx = __shfl_sync(0xffffffffu, mine.kx, j);
                                    tp.c = __shfl_sync(0xffffffffu, mine.c, j);
                                    cp_async16(sbase + ((j ^ (t & 7)) << 4), conv_src<EB>(P.conv, gvalid, gb, gy0, gx0, tp), P.conv.x);
                                }
                            }
                            asm volatile("cp.async.commit_group;" ::: "memory");  // one group per K-block, empty past the end
                        }
                        asm volatile("cp.async.wait_group %0;" ::"n"(GA_AHEAD) : "memory");  // block i's own chunks have landed
                        if (X3) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int off = t * 128 + (j << 4);
                                *reinterpret_cast<uint4 *>(stage + C::LO_A_OFF + off) = tf32_lo(*reinterpret_cast<const uint4 *>(stage + off));
                            }
                        }
                    } else {
                        uint4 gcur[8];
                        conv_gather_row<EB>(P.conv, gvalid, gb, gy0, gx0, (kb0 + i) * C::BK, P.K, lane, gcur);
                        mbar_wait(&empty_bar[s], ph ^ 1);  // the MMAs that read this stage's previous contents have retired
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int off = t * 128 + ((j ^ (t & 7)) << 4);  // SWIZZLE_128B: chunk j of row t sits at j ^ (t % 8)
                            *reinterpret_cast<uint4 *>(stage + off) = gcur[j];
                            if (X3) *reinterpret_cast<uint4 *>(stage + C::LO_A_OFF + off) = tf32_lo(gcur[j]);
                        }
                    }
                }
                if (X3) {
                    mbar_wait(&full_bar[s], ph);
                    // tcgen05 kind::tf32 ignores the low 13 mantissa bits of an fp32 operand (measured: bit-identical results
                    // for raw and pre-truncated inputs), so the landed tile itself is the hi operand; only lo = x - trunc(x)
                    // is written.  Rows past M / N are never stored, so they are skipped.
                    const uint4 *raw = reinterpret_cast<const uint4 *>(stage);
                    uint4 *lo = reinterpret_cast<uint4 *>(stage + C::RAW_BYTES);  // [B lo][A lo]
                    // batches of 8 x 16 B per thread: all loads first, then all stores -- raw and lo may alias as far as the
                    // compiler knows, so an element-wise loop serialises on shared-memory latency (measured 0.4 us per stage)
                    for (int j0 = t; j0 < ((P.xmode & 1) ? 0 : conv_units); j0 += 8 * NT) {
                        uint4 v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int j = j0 + u * NT;
                            const int jj = j < a_units ? j : j - a_units + C::A_BYTES / 16;
                            if (j < conv_units) v[u] = raw[jj];
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int j = j0 + u * NT;
                            const int jl = j < a_units ? j + C::B_BYTES / 16 : j - a_units;
                            if (j < conv_units && !(P.xmode & 8)) lo[jl] = tf32_lo(v[u]);
                        }
                    }
                }
                if (!(P.xmode & 4)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to UMMA
                mbar_arrive(&conv_bar[s]);
                // chain `drained` ended at K-block (drained+1)*CHAIN - 1; by the time two more blocks have been converted
                // its MMAs have retired, so the wait below does not stall the conversion pipeline
                if (X3 && drained < nchains - 1 && i + 1 >= (drained + 1) * CHAIN + 2) {
                    const int cb = drained & 1;
                    mbar_wait(&acc_full[cb], (drained >> 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int cc = 0; cc < HC; cc += 32) {
                        const int c = c0 + cc;
                        uint32_t r[32], r2[32];
                        tmem_ld32(tq + (uint32_t)(cb * C::ACC_COLS + c), r);
                        tmem_ld32(tq + (uint32_t)(cb * C::ACC_COLS + BN + c), r2);
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            acc[X3 ? cc + j : 0] = __fadd_rn(acc[X3 ? cc + j : 0], __fadd_rn(__uint_as_float(r[j]), __uint_as_float(r2[j])));
                    }
                    tc_fence_before();
                    mbar_arrive(&acc_empty[cb]);
                    ++drained;
                }
            }
            for (; X3 && drained < nchains - 1; ++drained) {  // chains whose delayed drain slot never came (short tail)
                const int cb = drained & 1;
                mbar_wait(&acc_full[cb], (drained >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int cc = 0; cc < HC; cc += 32) {
                    const int c = c0 + cc;
                    uint32_t r[32], r2[32];
                    tmem_ld32(tq + (uint32_t)(cb * C::ACC_COLS + c), r);
                    tmem_ld32(tq + (uint32_t)(cb * C::ACC_COLS + BN + c), r2);
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        acc[X3 ? cc + j : 0] = __fadd_rn(acc[X3 ? cc + j : 0], __fadd_rn(__uint_as_float(r[j]), __uint_as_float(r2[j])));
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[cb]);
            }
        }
        // epilogue.  Each warp pulls 32 x 32 chunks of the last chain's accumulator, adds the register sums, transposes
        // the chunk through shared memory (the pipeline stages are idle once the last chain has been committed) and then
        // walks the rows with lane = column, so every global access below is one coalesced 128-byte row segment.
        const int lastc = nchains - 1, lb = lastc & 1;
        // Whole-tile TMA store (data gradient): relu' of this thread's pixel as one bit per output column, fetched and packed NOW,
        // while the last chain's MMAs are still retiring -- inside the epilogue the loads cost a 2.5 us round trip per 32 columns
        // on the update's critical path (the epilogue warps run alone on their schedulers: nothing hides a dependent load).
        unsigned dmb[HC / 32];  // bits of this warp's chunks: dmb[i] <-> columns [c0 + 32 i, c0 + 32 i + 32)
#pragma unroll
        for (int i = 0; i < HC / 32; ++i) dmb[i] = 0xffffffffu;
        if (GA == 2 && P.tma_store == 2 && P.epi.mask && !(P.xmode & 16)) {
            const int r = q * 32 + lane;
            const int vN = min(BN, P.N - n0);
            const uint8_t *mrow = nullptr;
            if (r < vrows) {
                const int ry = r / cols_c, rx = r - ry * cols_c;
                const int y = (coy0 + ry) * P.conv.s + cls_py, xx = rx * P.conv.s + cls_px;
                if (y < P.conv.H && xx < P.conv.W)
                    mrow = (const uint8_t *)P.epi.mask + ((((long long)cimg * P.conv.H + y) * P.conv.W + xx) * P.epi.mask_sm + n0) * (P.epi.mask_bf16 ? 2 : 4);
            }
#pragma unroll
            for (int i = 0; i < HC / 32; ++i) {
                const int c = c0 + i * 32;
                unsigned bits = 0u;
                if (mrow && c < vN) {
                    if (P.epi.mask_bf16) {
                        uint4 mk4[4];
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            mk4[j4] = make_uint4(0u, 0u, 0u, 0u);
                            if (c + j4 * 8 + 8 <= vN) mk4[j4] = __ldg(reinterpret_cast<const uint4 *>(mrow + (c + j4 * 8) * 2));
                        }
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const uint32_t mw[4] = {mk4[j4].x, mk4[j4].y, mk4[j4].z, mk4[j4].w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                bits |= (__uint_as_float(mw[k] << 16) > 0.f ? 1u : 0u) << (j4 * 8 + 2 * k);
                                bits |= (__uint_as_float(mw[k] & 0xffff0000u) > 0.f ? 1u : 0u) << (j4 * 8 + 2 * k + 1);
                            }
                        }
                    } else {
                        float4 mk8[8];
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            mk8[j4] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (c + j4 * 4 + 4 <= vN) mk8[j4] = __ldg(reinterpret_cast<const float4 *>(mrow + (c + j4 * 4) * 4));
                        }
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4)
                            bits |= ((mk8[j4].x > 0.f ? 1u : 0u) | (mk8[j4].y > 0.f ? 2u : 0u) | (mk8[j4].z > 0.f ? 4u : 0u) | (mk8[j4].w > 0.f ? 8u : 0u)) << (4 * j4);
                    }
                }
                dmb[i] = bits;
            }
        }
        // the row-walker epilogues (plain GEMMs: fc1's data gradient) take their relu' bits the same way
        const bool pre_mask = GA == 0 && P.epi.mask && !P.ws && !P.epi.accumulate && P.tma_store == 0 &&
                              (!P.epi.rowmul || ((P.epi.rowmul_div & 31) == 0 && P.epi.sn == 1));
        if (pre_mask) {
            const bool ptrans = !P.epi.rowmul && P.epi.sm == 1 && P.epi.sn != 1;
            const int prow0 = m0 + q * 32, prows = max(0, min(32, vrows - q * 32));
            const long long pmsm = P.epi.mask_sm, pmsn = P.epi.mask_sn;
#pragma unroll
            for (int i = 0; i < HC / 32; ++i) {
                const int c = c0 + i * 32, n = n0 + c + lane;
                unsigned bits = 0u;
                if (n0 + c < P.N) {
                    const int cnt = ptrans ? min(32, P.N - (n0 + c)) : prows;
                    const bool mine = ptrans ? lane < prows : n < P.N;
                    const long long mbase = ptrans ? (long long)(prow0 + lane) * pmsm + (long long)(n0 + c) * pmsn : (long long)prow0 * pmsm + (long long)n * pmsn;
                    const long long mstep = ptrans ? pmsn : pmsm;
                    if (P.epi.mask_bf16) {
                        unsigned short mv[32];
                        const unsigned short *mp16 = (const unsigned short *)P.epi.mask + mbase;
#pragma unroll
                        for (int k = 0; k < 32; ++k) mv[k] = (mine && k < cnt) ? __ldg(mp16 + k * mstep) : (unsigned short)0x8000u;
#pragma unroll
                        for (int k = 0; k < 32; ++k) bits |= (__uint_as_float((unsigned)mv[k] << 16) > 0.f ? 1u : 0u) << k;
                    } else {
                        float mv[32];
                        const float *mp32 = (const float *)P.epi.mask + mbase;
#pragma unroll
                        for (int k = 0; k < 32; ++k) mv[k] = (mine && k < cnt) ? __ldg(mp32 + k * mstep) : 0.f;
#pragma unroll
                        for (int k = 0; k < 32; ++k) bits |= (mv[k] > 0.f ? 1u : 0u) << k;
                    }
                }
                dmb[i] = bits;
            }
        }
        mbar_wait(&acc_full[lb], (lastc >> 1) & 1);
        tc_fence_after();
        if (!(P.xmode & 32)) pdl_trigger();  // main loop done: the next kernel may become resident and run its prologue under this epilogue
        if (threadIdx.x == 64) dbg_stamp(P, 5);
        float(*stg)[33] = reinterpret_cast<float(*)[33]>(smem) + q * 32;  // 32 x 33 floats per warp, in pipeline stage 0
        const int row0 = m0 + q * 32;
        const int rows = max(0, min(32, vrows - q * 32));
        // the epilogue description is copied out of the parameter bank once: re-reading it per row costs a chain of
        // uniform constant loads per store (measured 100 ns per row)
        const Epilogue e = P.epi;
        const bool split = P.ws != nullptr;
        float *const ws = P.ws;
        const int Nn = P.N;
        const int bias_mode = e.bias_mode, relu = e.relu, out_bf16 = e.out_bf16;
        const long long sm = e.sm, sn = e.sn;
        // IQN merge in the epilogue: every row of a warp's 32-row chunk belongs to the same sample when rowmul_div % 32 == 0
        const bool rmul = e.rowmul != nullptr && (e.rowmul_div & 31) == 0 && GA == 0 && sn == 1;
        const void *const rmp = e.rowmul;
        const long long rm_ld = e.rowmul_ld;
        const int rm_div = e.rowmul_div;
        const bool simple = !e.rowmul || rmul;
        const bool trans = !e.rowmul && sm == 1 && sn != 1;
        const void *const maskp = e.mask;
        const int mask_bf16 = e.mask_bf16;
        const long long msm = e.mask_sm, msn = e.mask_sn;
        // CODE SIZE matters here: this code runs once per CTA, so it is fetched cold -- the column-chunk loops below are
        // deliberately NOT unrolled and the row loops only by 4 (measured: the fully unrolled epilogue took 1.4 us per
        // 32 x 32 chunk, instruction-fetch bound).
        // 3xTF32: the register sums are indexed statically, so one small unrolled pass first folds them into the last
        // chain and parks the whole 32 x BN slab of this warp in shared memory (row stride BN + 1: conflict-free both ways)
        constexpr int XLD = BN + 1;
        float *const xst = reinterpret_cast<float *>(smem) + q * 32 * XLD;
        if (X3 && GA == 0 && BN >= 64 && P.tma_store == 3) {
            // fp32-grade TMA-store epilogue: lane = tile row already holds 32 consecutive columns after tcgen05.ld, so the chain sums,
            // bias / relu / relu' mask / row multiplier are applied in registers and the row's 128 bytes go to a SWIZZLE_128B staging
            // block as eight 16-byte pieces (piece j of row r at j ^ (r % 8)); one bulk store per 32 x 32 chunk, two blocks per warp
            // (staging: 8 KB per warp = two 32 x 128-byte blocks; bias / multiplier copies behind the last warp's)
            float *const bias_s = reinterpret_cast<float *>(smem + TW::EW * 8192);
            float *const rm_s = bias_s + 128 + q * 128;  // per quadrant: the two warps of a quadrant write the same values
            {
                const int t = threadIdx.x - 64;
                if (t < BN) bias_s[t] = (bias_mode == 1 && n0 + t < Nn) ? __ldg(e.bias + n0 + t) : 0.f;
                for (int j = lane; j < BN; j += 32) rm_s[j] = (rmul && n0 + j < Nn) ? ldg_elem(rmp, 0, (long long)(min(row0, P.M - 1) / rm_div) * rm_ld + n0 + j) : 1.f;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
            const uint32_t sbase = smem_u32(smem) + (uint32_t)((c0 / HC) * 4 + q) * 8192;
            const bool rowok = lane < rows;
            int issued = 0;
#pragma unroll
            for (int cc = 0; cc < HC; cc += 32) {
                const int c = c0 + cc;
                if (n0 + c < Nn) {
                    if (issued >= 2) {  // the block about to be overwritten has been read by its bulk store
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        __syncwarp();
                    }
                    unsigned mbits = 0xffffffffu;
                    if (maskp) {  // relu' source: fp32 [M][msm]; one bit per column of this lane's row
                        mbits = 0u;
                        const float *mp = (const float *)maskp + (long long)(row0 + lane) * msm + n0 + c;
                        if (rowok && n0 + c + 32 <= Nn) {
                            float4 mv[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) mv[j] = __ldg(reinterpret_cast<const float4 *>(mp) + j);
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                mbits |= ((mv[j].x > 0.f ? 1u : 0u) | (mv[j].y > 0.f ? 2u : 0u) | (mv[j].z > 0.f ? 4u : 0u) | (mv[j].w > 0.f ? 8u : 0u)) << (4 * j);
                        } else {
#pragma unroll 1
                            for (int j = 0; j < 32; ++j)
                                if (rowok && n0 + c + j < Nn && __ldg(mp + j) > 0.f) mbits |= 1u << j;
                        }
                    }
                    uint32_t r[32], r2[32];
                    tmem_ld32(tq + (uint32_t)(lb * C::ACC_COLS + c), r);
                    tmem_ld32(tq + (uint32_t)(lb * C::ACC_COLS + BN + c), r2);
                    const uint32_t sb = sbase + (issued & 1) * 4096 + lane * 128;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        float x[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int j = j4 * 4 + i;
                            float v = __fadd_rn(__fadd_rn(__uint_as_float(r[j]), __uint_as_float(r2[j])), acc[X3 ? cc + j : 0]) + bias_s[c + j];
                            const float vr = fmaxf(v, 0.f);
                            v = relu ? vr : v;
                            v = ((mbits >> j) & 1u) ? v : 0.f;
                            x[i] = v * rm_s[c + j];
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + ((uint32_t)(j4 ^ (lane & 7)) << 4)), "r"(__float_as_uint(x[0])),
                                     "r"(__float_as_uint(x[1])), "r"(__float_as_uint(x[2])), "r"(__float_as_uint(x[3]))
                                     : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&P.tmC),
                                     "r"(sbase + (issued & 1) * 4096), "r"(n0 + c), "r"(row0)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    ++issued;
                }
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncwarp();
        } else {
        if (X3 && !(GA == 2 && P.tma_store == 2)) {  // (the whole-tile TMA store path reads TMEM itself)
#pragma unroll
            for (int cc = 0; cc < HC; cc += 32) {
                const int c = c0 + cc;
                uint32_t r[32], r2[32];
                tmem_ld32(tq + (uint32_t)(lb * C::ACC_COLS + c), r);
                tmem_ld32(tq + (uint32_t)(lb * C::ACC_COLS + BN + c), r2);
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    xst[lane * XLD + c + j] = __fadd_rn(__fadd_rn(__uint_as_float(r[j]), __uint_as_float(r2[j])), acc[X3 ? cc + j : 0]);
            }
            __syncwarp();
            if (threadIdx.x == 64) dbg_stamp(P, 9);
        }
        // bf16 output, plain row-major store: 64 columns at a time so that every lane writes a packed bf16x2 (128-byte row
        // segments instead of 64-byte ones) -- the wide, K-short GEMMs (IQN phi: K = 64, 0.5 GB of output) are store-bound
        if (GA == 2 && P.tma_store == 2) {
            // Whole-tile TMA store (data gradient).  Lane = tile row: every thread converts its row (32 columns per tcgen05.ld),
            // applies relu' of the pixel it belongs to and writes 16-byte chunks into the staging tile [column block][row][ts_span
            // bytes] (swizzled like the tensor map: conflict-free); one thread then issues one 4-D box per column block into the
            // residue class's strided view of dX -- rows past the class's last row are clipped by the map.
            const int vN = min(BN, Nn - n0);
            const int esz = out_bf16 ? 2 : 4;
            const int span = P.ts_span;
            const int r = q * 32 + lane;
            const uint32_t swz = span == 128 ? (uint32_t)(r & 7) : (span == 64 ? (uint32_t)((r >> 1) & 3) : 0u);
            const uint32_t stg0 = smem_u32(smem);
#pragma unroll
            for (int cc = 0; cc < HC; cc += 32) {
                const int c = c0 + cc;
                if (c < vN) {
                    uint32_t rr[32];
                    float x[32];
                    tmem_ld32(tq + (uint32_t)(lb * C::ACC_COLS + c), rr);
                    if (X3) {
                        uint32_t r2[32];
                        tmem_ld32(tq + (uint32_t)(lb * C::ACC_COLS + BN + c), r2);
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = __fadd_rn(__fadd_rn(__uint_as_float(rr[j]), __uint_as_float(r2[j])), acc[X3 ? cc + j : 0]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(rr[j]);
                    }
                    if (threadIdx.x == 64) dbg_stamp(P, cc == 0 ? 12 : 14);
                    if (bias_mode == 1) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] += (n0 + c + j < Nn) ? __ldg(e.bias + n0 + c + j) : 0.f;
                    }
                    if (relu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
                    }
                    if (maskp) {
                        const unsigned mb = dmb[cc / 32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = ((mb >> j) & 1u) ? x[j] : 0.f;
                    }
                    if (threadIdx.x == 64) dbg_stamp(P, cc == 0 ? 13 : 15);
                    if (out_bf16) {
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            uint32_t pk[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const __nv_bfloat162 hh = __floats2bfloat162_rn(x[j4 * 8 + 2 * i], x[j4 * 8 + 2 * i + 1]);
                                pk[i] = *reinterpret_cast<const uint32_t *>(&hh);
                            }
                            const int off = c * 2 + j4 * 16, box = off / span, chunk = (off - box * span) >> 4;
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg0 + (uint32_t)(box * BM * span + r * span) + (((uint32_t)chunk ^ swz) << 4)),
                                         "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                        }
                    } else {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const int off = c * 4 + j4 * 16, box = off / span, chunk = (off - box * span) >> 4;
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg0 + (uint32_t)(box * BM * span + r * span) + (((uint32_t)chunk ^ swz) << 4)),
                                         "r"(__float_as_uint(x[j4 * 4])), "r"(__float_as_uint(x[j4 * 4 + 1])), "r"(__float_as_uint(x[j4 * 4 + 2])),
                                         "r"(__float_as_uint(x[j4 * 4 + 3])) : "memory");
                        }
                    }
                }
            }
            if (threadIdx.x == 64) dbg_stamp(P, 9);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
            if (threadIdx.x == 64) {
                dbg_stamp(P, 10);
                const int nbox = (vN * esz + span - 1) / span;
                for (int bx = 0; bx < nbox; ++bx)
                    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&P.tmC4[blockIdx.z]),
                                 "r"(stg0 + (uint32_t)(bx * BM * span)), "r"(n0 + bx * (span / esz)), "r"(0), "r"(coy0), "r"(cimg)
                                 : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                dbg_stamp(P, 11);
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            }
        } else
        if (BN >= 64 && !X3 && GA == 0 && P.tma_store) {
            // TMA-store epilogue.  tcgen05.ld leaves lane = tile row with 32 consecutive columns in registers -- exactly one
            // 64-byte half of a 128-byte row of the staging block, so the values go to shared memory as four 16-byte chunks per
            // half (chunk j of row r at j ^ (r % 8): conflict-free, and the layout the SWIZZLE_128B tensor map expects) with no
            // transpose; bias / rowmul are staged once per CTA / warp and read as broadcasts.  Two staging blocks per warp, so
            // the bulk store of one 64-column group overlaps the conversion of the next.
            float *const bias_s = reinterpret_cast<float *>(smem + 32768);
            float *const rm_s = bias_s + 128 + q * 128;
            {
                const int t = threadIdx.x - 64;
                if (t < BN) bias_s[t] = (bias_mode == 1 && n0 + t < Nn) ? __ldg(e.bias + n0 + t) : 0.f;
                for (int j = lane; j < BN; j += 32) rm_s[j] = (rmul && n0 + j < Nn) ? ldg_elem(rmp, 1, (long long)(row0 / rm_div) * rm_ld + n0 + j) : 1.f;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const uint32_t sbase = smem_u32(smem) + q * 8192;
            const bool rowok = lane < rows;
            int buf = 0, issued = 0;
#pragma unroll 1
            for (int c = 0; c < BN; c += 64) {
                if (n0 + c >= Nn) break;
                if (issued >= 2) {  // the block about to be overwritten has been read by its bulk store
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                }
                const uint32_t sb = sbase + buf * 4096 + lane * 128;
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    uint32_t r[32];
                    tmem_ld32(tq + (uint32_t)(lb * BN + c + 32 * h), r);
                    const int nb = n0 + c + 32 * h;
                    uint4 mk[4];
                    if (maskp) {
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
                        uint32_t pk[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int col = c + 32 * h + j8 * 8 + 2 * i;
                            float x0 = __uint_as_float(r[j8 * 8 + 2 * i]) + bias_s[col], x1 = __uint_as_float(r[j8 * 8 + 2 * i + 1]) + bias_s[col + 1];
                            const float r0f = fmaxf(x0, 0.f), r1f = fmaxf(x1, 0.f);
                            x0 = relu ? r0f : x0;
                            x1 = relu ? r1f : x1;
                            if (maskp) {
                                x0 = __uint_as_float(mw[i] << 16) > 0.f ? x0 : 0.f;
                                x1 = __uint_as_float(mw[i] & 0xffff0000u) > 0.f ? x1 : 0.f;
                            }
                            const __nv_bfloat162 hh = __floats2bfloat162_rn(x0 * rm_s[col], x1 * rm_s[col + 1]);
                            pk[i] = *reinterpret_cast<const uint32_t *>(&hh);
                        }
                        const int jc = h * 4 + j8;
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + ((uint32_t)(jc ^ (lane & 7)) << 4)), "r"(pk[0]), "r"(pk[1]),
                                     "r"(pk[2]), "r"(pk[3])
                                     : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&P.tmC),
                                 "r"(sbase + buf * 4096), "r"(n0 + c), "r"(row0)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                ++issued;
                buf ^= 1;
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncwarp();
        } else
        if (BN >= 64 && !X3 && simple && !trans && !split && !e.accumulate && out_bf16 && sn == 1 && (Nn & 1) == 0 && (sm & 1) == 0 && bias_mode != 2 &&
            sm < (1 << 24) && msm < (1 << 24) && (!rmul || (rm_ld & 1) == 0) &&
            (!maskp || (mask_bf16 && msn == 1 && (msm & 1) == 0))) {
            float(*stw)[66] = reinterpret_cast<float(*)[66]>(smem) + q * 32;  // 32 x 66 floats per warp (spans idle stages)
#pragma unroll 1
            for (int c = 0; c < BN; c += 64) {
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    uint32_t r[32];
                    tmem_ld32(tq + (uint32_t)(lb * BN + c + 32 * h), r);
#pragma unroll
                    for (int j = 0; j < 32; ++j) stw[lane][32 * h + j] = __uint_as_float(r[j]);
                }
                __syncwarp();
                const int n = n0 + c + 2 * lane;
                if (n < Nn) {
                    float b0 = 0.f, b1 = 0.f;
                    if (bias_mode == 1) { b0 = __ldg(e.bias + n); b1 = __ldg(e.bias + n + 1); }
                    __nv_bfloat162 *op = reinterpret_cast<__nv_bfloat162 *>((bf16 *)e.out + (long long)row0 * sm + n);
                    const __nv_bfloat162 *mp = maskp ? reinterpret_cast<const __nv_bfloat162 *>((const bf16 *)maskp + (long long)row0 * msm + n) : nullptr;
                    float rm0 = 1.f, rm1 = 1.f;
                    if (rmul) {
                        const __nv_bfloat162 r2 = __ldg(reinterpret_cast<const __nv_bfloat162 *>((const bf16 *)rmp + (long long)(row0 / rm_div) * rm_ld + n));
                        rm0 = __low2float(r2); rm1 = __high2float(r2);
                    }
                    if (mp) epi_rows_store_bf16x2<true>(&stw[0][2 * lane], rows, op, (int)(sm >> 1), b0, b1, relu, mp, (int)(msm >> 1), rm0, rm1);
                    else epi_rows_store_bf16x2<false>(&stw[0][2 * lane], rows, op, (int)(sm >> 1), b0, b1, relu, nullptr, 0, rm0, rm1);
                }
                __syncwarp();
            }
        } else {
        // Split-K without a finishing launch: pass 0 adds this CTA's partial tile into the workspace; the last CTA to arrive
        // on the tile (per-tile counter behind the workspace) runs pass 1: it re-reads the summed chunks (coalesced, all 32
        // rows of a column in flight at once), zeroes them for the next GEMM, restages them in shared memory and then
        // stores them exactly like an unsplit tile -- so a transposed (swap-AB) output is written coalesced too.
        bool red_ws = split, refill = false;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
#pragma unroll 1
        for (int c = c0; c < c0 + HC; c += 32) {
            if (n0 + c >= Nn) break;  // the whole chunk lies past N (uniform)
            constexpr int LDC = X3 ? XLD : 33;
            float *const cbuf = X3 ? xst + c : &stg[0][0];  // element (r, j) of the chunk at cbuf[r * LDC + j]
            const int n = n0 + c + lane;
            if (refill) {
#pragma unroll 1
                for (int h = 0; h < 32; h += 16) {  // 16 rows of the column in flight at a time
                    float v[16];
                    float *wp = in_reg(ws + (long long)(row0 + h) * Nn + n);  // (pins the address arithmetic inside the loop: hoisted, it spilled)
                    if (n < Nn && h + 16 <= rows) {
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) v[rr] = __ldcg(wp + (long long)rr * Nn);
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) __stcg(wp + (long long)rr * Nn, 0.f);
                    } else {
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) v[rr] = (n < Nn && h + rr < rows) ? __ldcg(wp + (long long)rr * Nn) : 0.f;
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr)
                            if (n < Nn && h + rr < rows) __stcg(wp + (long long)rr * Nn, 0.f);
                    }
#pragma unroll
                    for (int rr = 0; rr < 16; ++rr) cbuf[(h + rr) * LDC + lane] = v[rr];
                }
                __syncwarp();
            } else if (!X3) {
                uint32_t r[32];
                tmem_ld32(tq + (uint32_t)(lb * BN + c), r);
                if (c == 0 && threadIdx.x == 64) dbg_stamp(P, 9);
#pragma unroll
                for (int j = 0; j < 32; ++j) stg[lane][j] = __uint_as_float(r[j]);
                __syncwarp();
            }
            const float *const chunk = cbuf;
            constexpr int ld = LDC;
            if (c == 0 && threadIdx.x == 64 && !refill) dbg_stamp(P, 10);
            const bool direct = refill || (!red_ws && !e.accumulate);
            // relu' mask of the 32 x 32 chunk as one bit per element: 32 independent loads in flight instead of one
            // dependent load per stored row.  Bit i = row i of column `lane` (or column i of row `lane` when transposed).
            unsigned mbits = 0xffffffffu;
            const void *mq = maskp;
            asm volatile("" : "+l"(mq));  // keeps the mask loads below inside this loop body
            if (pre_mask) {  // (selected by compares: a dynamic index would send the bits through local memory)
                mbits = dmb[0];
                if (HC > 32 && c - c0 == 32) mbits = dmb[HC > 32 ? 1 : 0];
                if (HC > 64 && c - c0 == 64) mbits = dmb[HC > 64 ? 2 : 0];
                if (HC > 96 && c - c0 == 96) mbits = dmb[HC > 96 ? 3 : 0];
            } else if (mq && direct && simple) {
                mbits = 0u;
                const int cnt = trans ? min(32, Nn - (n0 + c)) : rows;
                const bool mine = trans ? lane < rows : n < Nn;
                const long long mbase = trans ? (long long)(row0 + lane) * msm + (long long)(n0 + c) * msn : (long long)row0 * msm + (long long)n * msn;
                const long long mstep = trans ? msn : msm;
                // (one load per register: written as "load, then test" per element type, the compiler funnels all 32 loads through
                // one destination register and they serialise on L2 latency -- measured 6.8 us for this block on fc1's data gradient)
                if (mask_bf16) {
                    unsigned short mv[32];
                    const unsigned short *mp16 = (const unsigned short *)mq + mbase;
#pragma unroll
                    for (int i = 0; i < 32; ++i) mv[i] = (mine && i < cnt) ? __ldg(mp16 + i * mstep) : (unsigned short)0x8000u;
#pragma unroll
                    for (int i = 0; i < 32; ++i) mbits |= (__uint_as_float((unsigned)mv[i] << 16) > 0.f ? 1u : 0u) << i;
                } else {
                    float mv[32];
                    const float *mp32 = (const float *)mq + mbase;
#pragma unroll
                    for (int i = 0; i < 32; ++i) mv[i] = (mine && i < cnt) ? __ldg(mp32 + i * mstep) : 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) mbits |= (mv[i] > 0.f ? 1u : 0u) << i;
                }
            }
            if (c == 0 && threadIdx.x == 64 && !refill) dbg_stamp(P, 15);
            const bool narrow = sm < (1 << 24) && sn < (1 << 24) && Nn < (1 << 24);  // 32-bit strides inside a chunk
            const bool hasmask = maskp != nullptr;
            if (trans && direct && narrow && bias_mode != 1) {
                // transposed output (swap-AB dispatch: consecutive m are contiguous in memory): lane = tile row
                const int m = row0 + lane;
                const int ncols = min(32, Nn - (n0 + c));
                if (lane < rows) {
                    const float bm = bias_mode == 2 ? __ldg(e.bias + m) : 0.f;
                    const long long o0 = (long long)m * sm + (long long)(n0 + c) * sn;
                    const float *rp = chunk + lane * ld;
                    if (out_bf16) {
                        if (hasmask) epi_cols_store<bf16, true>(rp, ncols, (bf16 *)e.out + o0, (int)sn, bm, relu, mbits);
                        else epi_cols_store<bf16, false>(rp, ncols, (bf16 *)e.out + o0, (int)sn, bm, relu, mbits);
                    } else {
                        if (hasmask) epi_cols_store<float, true>(rp, ncols, (float *)e.out + o0, (int)sn, bm, relu, mbits);
                        else epi_cols_store<float, false>(rp, ncols, (float *)e.out + o0, (int)sn, bm, relu, mbits);
                    }
                }
            } else if (!direct && narrow && n0 + c + 32 <= Nn && (red_ws ? ((Nn & 3) == 0 && ((uintptr_t)ws & 15) == 0)
                                                                         : (sn == 1 && (sm & 3) == 0 && ((uintptr_t)e.out & 15) == 0)) && ((n0 + c) & 3) == 0) {
                // split-K partial sums into the workspace / wgrad-style accumulation into the output, four columns per red
                float *bp = red_ws ? ws + (long long)row0 * Nn + n0 + c : (float *)e.out + (long long)row0 * sm + n0 + c;
                epi_rows_red_v4<LDC>(chunk, rows, bp, red_ws ? Nn : (int)sm, lane);
            } else if (n < Nn) {
                if (!direct && narrow) {
                    // split-K partial sums into the workspace, or wgrad-style accumulation into the output
                    float *bp = red_ws ? ws + (long long)row0 * Nn + n : (float *)e.out + (long long)row0 * sm + (long long)n * sn;
                    epi_rows_red<LDC>(chunk + lane, rows, bp, red_ws ? Nn : (int)sm);
                } else if (direct && simple && narrow && bias_mode != 2 && !trans) {
                    const float bv = bias_mode == 1 ? __ldg(e.bias + n) : 0.f;
                    const float rm = rmul ? ldg_elem(rmp, out_bf16, (long long)(row0 / rm_div) * rm_ld + n) : 1.f;
                    const long long o0 = (long long)row0 * sm + (long long)n * sn;
                    const float *cp = chunk + lane;
                    if (out_bf16) {
                        if (hasmask) epi_rows_store<bf16, LDC, true>(cp, rows, (bf16 *)e.out + o0, (int)sm, bv, relu, mbits, rm);
                        else epi_rows_store<bf16, LDC, false>(cp, rows, (bf16 *)e.out + o0, (int)sm, bv, relu, mbits, rm);
                    } else {
                        if (hasmask) epi_rows_store<float, LDC, true>(cp, rows, (float *)e.out + o0, (int)sm, bv, relu, mbits, rm);
                        else epi_rows_store<float, LDC, false>(cp, rows, (float *)e.out + o0, (int)sm, bv, relu, mbits, rm);
                    }
                } else if (!direct) {
                    float *bp = red_ws ? ws + (long long)row0 * Nn + n : (float *)e.out + (long long)row0 * sm + (long long)n * sn;
                    const long long bs = red_ws ? (long long)Nn : sm;
#pragma unroll 1
                    for (int rr = 0; rr < rows; ++rr) atomicAdd(bp + (long long)rr * bs, chunk[rr * ld + lane]);
                } else if (simple) {
#pragma unroll 1
                    for (int rr = 0; rr < rows; ++rr) {  // rare shapes (row bias without swap-AB, very large strides): one element at a time
                        const int m = row0 + rr;
                        float x = chunk[rr * ld + lane];
                        if (bias_mode == 1) x += __ldg(e.bias + n);
                        if (bias_mode == 2) x += __ldg(e.bias + m);
                        if (relu) x = fmaxf(x, 0.f);
                        if (maskp) x = ldg_elem(maskp, mask_bf16, (long long)m * msm + (long long)n * msn) > 0.f ? x : 0.f;
                        if (rmp) x *= ldg_elem(rmp, out_bf16, (long long)(m / rm_div) * rm_ld + n);
                        st_elem(e.out, out_bf16, (long long)m * sm + (long long)n * sn, x);
                    }
                } else {
#pragma unroll 1
                    for (int rr = 0; rr < rows; ++rr) epi_store(e, chunk[rr * ld + lane], row0 + rr, n);
                }
            }
            __syncwarp();
            if (c == 0 && threadIdx.x == 64 && !refill) dbg_stamp(P, 11);
        }
        if (!red_ws) break;
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        unsigned *ctr = P.ctr + blockIdx.y * gridDim.x + blockIdx.x;
        if (threadIdx.x == 64) {
            __threadfence();  // cumulative: the other 127 threads' reds, observed through the barrier, are ordered before the ticket
            const unsigned old = atomicAdd(ctr, 1u);
            const bool last = old == gridDim.z - 1;
            if (last) *ctr = 0u;
            *tmem_slot = last ? 1u : 0u;  // tmem_slot is dead after tmem_base was read
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        if (!*tmem_slot) break;
        __threadfence();
        red_ws = false;
        refill = true;
        }
        }
        }  // tma_store != 3
    }
    if (threadIdx.x == 64) dbg_stamp(P, 6);
    if (P.xmode & 32) pdl_trigger();  // (experiment: dependents become resident only when this CTA is about to leave)
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) dbg_stamp(P, 7);
    if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
    if (threadIdx.x == 32) dbg_stamp(P, 8);
}

// ---------------------------------------------------------------- launch templates ---------
static inline bool tc_wide_enabled() {  // ZOO_TC_WIDE=0: four epilogue / converter warps everywhere (A/B)
    static const bool on = [] { const char *e = getenv("ZOO_TC_WIDE"); return !(e && e[0] == '0'); }();
    return on;
}
static constexpr int TC_EPI_SMEM = 40 * 1024;  // the epilogue stages through the (idle) pipeline stages: never launch with less
template <int EB, int BN, bool A_K, bool B_K, bool X3>
static zoo_status launch_tc(const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &P, dim3 grid, cudaStream_t st) {
    using C = TcCfg<EB, BN, X3>;
    static bool attr_set = false;
    if (!attr_set) {
        ZOO_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<EB, BN, A_K, B_K, X3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr_set = true;
    }
    ZOO_CUDA(launch_pdl(tc_gemm_kernel<EB, BN, A_K, B_K, X3, 0>, grid, dim3(TcWarps<BN, X3, 0>::THREADS), (size_t)C::smem_for(P.stages), st, tmA, tmB, P));
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

template <int EB, int BN, bool X3>
static zoo_status launch_tc_major(bool ak, bool bk, const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &P, dim3 grid,
                                  cudaStream_t st) {
    if (ak && bk) return launch_tc<EB, BN, true, true, X3>(tmA, tmB, P, grid, st);
    if (ak && !bk) return launch_tc<EB, BN, true, false, X3>(tmA, tmB, P, grid, st);
    if (!ak && bk) return launch_tc<EB, BN, false, true, X3>(tmA, tmB, P, grid, st);
    return launch_tc<EB, BN, false, false, X3>(tmA, tmB, P, grid, st);
}

template <int EB, bool X3>
static zoo_status launch_tc_bn(int BN, bool ak, bool bk, const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &P, dim3 grid,
                               cudaStream_t st) {
    if (BN == 32) {
        if (EB == 2) {  // bf16 MN-major boxes are 64 wide: BN = 32 only with a K-major B
            if (ak) return launch_tc<EB, 32, true, true, X3>(tmA, tmB, P, grid, st);
            return launch_tc<EB, 32, false, true, X3>(tmA, tmB, P, grid, st);
        }
        return launch_tc_major<EB, 32, X3>(ak, bk, tmA, tmB, P, grid, st);
    }
    if (BN == 64) return launch_tc_major<EB, 64, X3>(ak, bk, tmA, tmB, P, grid, st);
    return launch_tc_major<EB, 128, X3>(ak, bk, tmA, tmB, P, grid, st);
}

template <int EB, int BN, bool B_K, bool X3>
static zoo_status launch_conv(const CUtensorMap &tmB, const TcParams &P, dim3 grid, cudaStream_t st) {
    using C = TcCfg<EB, BN, X3>;
    static bool attr_set = false;
    if (!attr_set) {
        ZOO_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<EB, BN, true, B_K, X3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr_set = true;
    }
    ZOO_CUDA(launch_pdl(tc_gemm_kernel<EB, BN, true, B_K, X3, 1>, grid, dim3(TcWarps<BN, X3, 1>::THREADS), (size_t)C::SMEM, st, tmB, tmB, P));
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

template <int EB, int BN, bool X3, bool B_K = true, int SPAN = 128>
static zoo_status launch_conv_tma(const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &P, dim3 grid, cudaStream_t st) {
    using C = TcCfg<EB, BN, X3, SPAN>;
    static bool attr_set = false;
    if (!attr_set) {
        ZOO_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<EB, BN, true, B_K, X3, 2, SPAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(C::SMEM, TC_EPI_SMEM)));
        if constexpr (X3 && BN >= 64)
            ZOO_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<EB, BN, true, B_K, X3, 2, SPAN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(C::SMEM, TC_EPI_SMEM)));
        attr_set = true;
    }
    if constexpr (X3 && BN >= 64) {
        if ((long long)grid.x * grid.y * grid.z <= g_sm_budget && tc_wide_enabled()) {  // one CTA per SM anyway: eight epilogue / converter warps
            ZOO_CUDA(launch_pdl(tc_gemm_kernel<EB, BN, true, B_K, X3, 2, SPAN, true>, grid, dim3(TcWarps<BN, X3, 2, true>::THREADS),
                                (size_t)std::max(C::smem_for(P.stages), TC_EPI_SMEM), st, tmA, tmB, P));
            ZOO_LAUNCH_CHECK();
            return ZOO_OK;
        }
    }
    ZOO_CUDA(launch_pdl(tc_gemm_kernel<EB, BN, true, B_K, X3, 2, SPAN>, grid, dim3(TcWarps<BN, X3, 2>::THREADS), (size_t)std::max(C::smem_for(P.stages), TC_EPI_SMEM), st, tmA, tmB, P));
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

template <int EB, bool X3, int SPAN>
static zoo_status launch_conv_wgrad(const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &P0, dim3 grid, cudaStream_t st) {
    using C = TcCfg<EB, 128, X3, SPAN, true>;
    static bool attr_set = false;
    if (!attr_set) {
        ZOO_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<EB, 128, false, false, X3, 4, SPAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        if constexpr (X3) ZOO_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<EB, 128, false, false, X3, 4, SPAN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr_set = true;
    }
    TcParams P = P0;
    const long long ctas = (long long)grid.x * grid.y * grid.z;
    int stg = ctas > g_sm_budget ? C::TWO_CTA_STAGES : C::MAX_STAGES;
    if (stg > P.kb_per_split + 1) stg = P.kb_per_split + 1;
    if (stg < 2) stg = 2;
    P.stages = stg;
    P.xmode = 0;
    if constexpr (X3) {
        if (ctas <= g_sm_budget && tc_wide_enabled()) {
            ZOO_CUDA(launch_pdl(tc_gemm_kernel<EB, 128, false, false, X3, 4, SPAN, true>, grid, dim3(TcWarps<128, X3, 4, true>::THREADS),
                                (size_t)std::max(C::smem_for(stg), TC_EPI_SMEM), st, tmA, tmB, P));
            ZOO_LAUNCH_CHECK();
            return ZOO_OK;
        }
    }
    ZOO_CUDA(launch_pdl(tc_gemm_kernel<EB, 128, false, false, X3, 4, SPAN>, grid, dim3(TcWarps<128, X3, 4>::THREADS), (size_t)std::max(C::smem_for(stg), TC_EPI_SMEM), st, tmA, tmB, P));
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

template <int EB, bool X3>
static zoo_status launch_conv_bn(int BN, bool bk, const CUtensorMap &tmB, const TcParams &P, dim3 grid, cudaStream_t st) {
    if (BN == 32) {
        if (bk) return launch_conv<EB, 32, true, X3>(tmB, P, grid, st);
        if (EB == 4) return launch_conv<4, 32, false, X3>(tmB, P, grid, st);
        set_err("tc_conv_gemm: bf16 MN-major B needs BN >= 64");
        return ZOO_BAD_ARG;
    }
    if (BN == 64) return bk ? launch_conv<EB, 64, true, X3>(tmB, P, grid, st) : launch_conv<EB, 64, false, X3>(tmB, P, grid, st);
    return bk ? launch_conv<EB, 128, true, X3>(tmB, P, grid, st) : launch_conv<EB, 128, false, X3>(tmB, P, grid, st);
}


// One translation unit per operand mode (gemm_tc_bf16.cu / gemm_tc_tf32.cu / gemm_tc_x3.cu) instantiates the kernels, so the
// three compile in parallel.  kind: 0 = plain GEMM (both operands by TMA), 1 = gathered-A conv, 2 = TMA-im2col conv, 3 = TMA-im2col data gradient.
template <int EB, bool X3, int SPAN = 128>
static int tc_pick_stages(int BN, long long ctas, int nkb) {
    const int mx = BN == 32 ? TcCfg<EB, 32, X3, SPAN>::MAX_STAGES : (BN == 64 ? TcCfg<EB, 64, X3, SPAN>::MAX_STAGES : TcCfg<EB, 128, X3, SPAN>::MAX_STAGES);
    const int two = BN == 32 ? TcCfg<EB, 32, X3, SPAN>::TWO_CTA_STAGES : (BN == 64 ? TcCfg<EB, 64, X3, SPAN>::TWO_CTA_STAGES : TcCfg<EB, 128, X3, SPAN>::TWO_CTA_STAGES);
    int stg = (ctas > g_sm_budget || g_tc_shallow) ? two : mx;
    // bf16 grids of many waves (the large-batch convolutions): a two-deep ring lets three or four CTAs share an SM and overlap
    // their prologues / epilogues -- Rainbow B = 1024 conv2 / conv3 forward 39 -> 30 us each, step 1187 -> 1100 us; IQN 512 x 64
    // 2434 -> 2388 us.  (The fp32-grade kernels lose with it: their converter warps need the deeper ring.)
    if (EB == 2 && !X3 && ctas > 4LL * g_sm_budget) stg = 2;
    const int force = g_tc_stages_force & 0xff;  // ZOO_TC_STAGES / zoo_test_set_tc_stages
    if (force >= 2) stg = force < mx ? force : mx;
    if (stg > nkb + 1) stg = nkb + 1 < 2 ? 2 : nkb + 1;  // no point in a ring deeper than the K loop
    if (SPAN == 64 && stg < 4) stg = 4;  // the epilogue stages ~35 KB through the ring: keep the barriers behind it
    return stg;
}

template <int EB, bool X3>
static zoo_status tc_launch_mode(int kind, int BN, bool ak, bool bk, const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &P0, dim3 grid,
                                 cudaStream_t st) {
    TcParams P = P0;
    P.stages = tc_pick_stages<EB, X3>(BN, (long long)grid.x * grid.y * grid.z, P0.kb_per_split);
    P.xmode = g_tc_stages_force >> 8;
    if (kind == 2 && P0.span == 64) {  // bf16 TMA-im2col convolution over 64-byte runs (SWIZZLE_64B tiles)
        if constexpr (EB == 2 && !X3) {
            P.stages = tc_pick_stages<2, false, 64>(BN, (long long)grid.x * grid.y * grid.z, P0.kb_per_split);
            if (BN == 32) return launch_conv_tma<2, 32, false, true, 64>(tmA, tmB, P, grid, st);
            if (BN == 64) return launch_conv_tma<2, 64, false, true, 64>(tmA, tmB, P, grid, st);
            set_err("tc_conv: 64-byte spans cover N <= 64");
            return ZOO_BAD_ARG;
        }
        set_err("tc_conv: 64-byte spans are a bf16 path");
        return ZOO_BAD_ARG;
    }
    if (kind == 4) {  // TMA-im2col weight gradient (both operands MN-major boxes)
        if (P0.span == 64) {
            if constexpr (EB == 2 && !X3) return launch_conv_wgrad<2, false, 64>(tmA, tmB, P0, grid, st);
            set_err("tc_conv_wgrad: 64-byte spans are a bf16 path");
            return ZOO_BAD_ARG;
        }
        return launch_conv_wgrad<EB, X3, 128>(tmA, tmB, P0, grid, st);
    }
    if (kind == 0) return launch_tc_bn<EB, X3>(BN, ak, bk, tmA, tmB, P, grid, st);
    if (kind == 1) return launch_conv_bn<EB, X3>(BN, bk, tmB, P, grid, st);
    if (kind == 3) {  // TMA-im2col data gradient: MN-major weight operand (Cin <= 64)
        if constexpr (EB == 4) {
            if (BN == 32) return launch_conv_tma<4, 32, X3, false>(tmA, tmB, P, grid, st);
            if (BN == 64) return launch_conv_tma<4, 64, X3, false>(tmA, tmB, P, grid, st);
        } else {
            if (BN == 64) return launch_conv_tma<2, 64, false, false>(tmA, tmB, P, grid, st);
        }
        set_err("tc_conv: the TMA data-gradient path covers Cin <= 64");
        return ZOO_BAD_ARG;
    }
    if (BN == 32) return launch_conv_tma<EB, 32, X3>(tmA, tmB, P, grid, st);
    if (BN == 64) return launch_conv_tma<EB, 64, X3>(tmA, tmB, P, grid, st);
    return launch_conv_tma<EB, 128, X3>(tmA, tmB, P, grid, st);
}
zoo_status tc_launch_bf16(int kind, int BN, bool ak, bool bk, const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &P, dim3 grid, cudaStream_t st);
zoo_status tc_launch_tf32(int kind, int BN, bool ak, bool bk, const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &P, dim3 grid, cudaStream_t st);
zoo_status tc_launch_x3(int kind, int BN, bool ak, bool bk, const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &P, dim3 grid, cudaStream_t st);

}  // namespace zoo
