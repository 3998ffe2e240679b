// Host side of the tcgen05 GEMM engine: tensor maps, tiling / split-K policy, dispatch to the per-mode translation units
// (the kernel itself lives in gemm_tc.cuh).
#include <algorithm>
#include <mutex>

#include "gemm_tc.cuh"

namespace zoo {

unsigned long long *g_tc_dbg = nullptr;  // set by zoo_test_gemm_trace
thread_local int g_tc_shallow = 0;

// Per-launch phase trace (zoo_test_trace_*): while armed, launch i of the engine gets its own 16-stamp slot, so an update that is
// captured into a CUDA graph while armed keeps writing CTA (0,0,0)'s phase stamps on every replay -- the graph's real timeline.
unsigned long long *g_tc_trace = nullptr;
int g_tc_trace_cap = 0, g_tc_trace_n = 0;
std::vector<std::string> g_tc_trace_names;
static unsigned long long *trace_slot(const char *kind, int M, int N, int K, dim3 grid) {
    if (!g_tc_trace) return g_tc_dbg;
    if (g_tc_trace_n >= g_tc_trace_cap) return nullptr;
    char buf[96];
    snprintf(buf, sizeof(buf), "%s[%dx%dx%d]<%u,%u,%u>", kind, M, N, K, grid.x, grid.y, grid.z);
    g_tc_trace_names.push_back(buf);
    return g_tc_trace + 16 * (g_tc_trace_n++);
}

static zoo_status tc_launch(int eb, bool x3, int kind, int BN, bool ak, bool bk, const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &P,
                            dim3 grid, cudaStream_t st) {
    if (eb == 2) return tc_launch_bf16(kind, BN, ak, bk, tmA, tmB, P, grid, st);
    if (x3) return tc_launch_x3(kind, BN, ak, bk, tmA, tmB, P, grid, st);
    return tc_launch_tf32(kind, BN, ak, bk, tmA, tmB, P, grid, st);
}

// ---------------------------------------------------------------- host side ---------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

static EncodeTiledFn get_encode() {
    std::call_once(g_encode_once, [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            g_encode = (EncodeTiledFn)fn;
    });
    return g_encode;
}

// 2D tensor map: inner (contiguous) extent d0, outer extent d1, row pitch ld elements, box {b0, b1}, SWIZZLE_128B
static zoo_status make_map(CUtensorMap *tm, const void *p, int eb, long long d0, long long d1, long long ld, int b0, int b1, bool mn_major, int span = 128) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_err("cuTensorMapEncodeTiled not available from the driver");
        return ZOO_CUDA_ERROR;
    }
    cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
    cuuint64_t strides[1] = {(cuuint64_t)ld * eb};
    cuuint32_t box[2] = {(cuuint32_t)b0, (cuuint32_t)b1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(p), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     span == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : ((eb == 4 && mn_major) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_err("cuTensorMapEncodeTiled failed (%d): dims %lld x %lld ld %lld box %d x %d ptr %p", (int)r, d0, d1, ld, b0, b1, p);
        return ZOO_CUDA_ERROR;
    }
    return ZOO_OK;
}

// 4-D tensor map over an NHWC activation tensor (or an overlapping-window view of it) for the CONV_TMA A operand
static zoo_status make_map4(CUtensorMap *tm, const void *p, int eb, const cuuint64_t dims[4], const cuuint64_t strides_bytes[3],
                            const cuuint32_t box[4], const cuuint32_t estr[4], int span = 128) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_err("cuTensorMapEncodeTiled not available from the driver"); return ZOO_CUDA_ERROR; }
    CUresult r = enc(tm, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void *>(p), dims, strides_bytes,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, span == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_err("cuTensorMapEncodeTiled(4D) failed (%d): dims %llu %llu %llu %llu box %u %u %u %u estr %u %u %u %u", (int)r, (unsigned long long)dims[0],
                (unsigned long long)dims[1], (unsigned long long)dims[2], (unsigned long long)dims[3], box[0], box[1], box[2], box[3], estr[0], estr[1],
                estr[2], estr[3]);
        return ZOO_CUDA_ERROR;
    }
    return ZOO_OK;
}

bool tc_gemm_supported(const Operand &A, const Operand &B, int M, int N, int K) {
    if (A.is_bf16 != B.is_bf16) return false;
    const int al = A.is_bf16 ? 8 : 4;  // elements per 16 bytes
    if (M < 1 || N < 8 || K < al) return false;
    if (A.ld % al || B.ld % al) return false;                                  // 16-byte global strides
    if (((uintptr_t)A.p & 15) || ((uintptr_t)B.p & 15)) return false;        // 16-byte base alignment
    return true;
}

static bool midk_split_enabled(const Epilogue &epi, int BN) {
    // OFF by default (it bought no end-to-end time).  Bit 0: unmasked GEMMs (forward), bit 1: masked ones (dgrad); bits 2 / 3
    // restrict it to BN >= 64 / BN <= 32.  With bit 0 set, test_rainbow_atari[fp32] in the explicit-conv mode (ZOO_TMA_CONV=0)
    // fails about one run in four, always by the same 9.5e-3 on conv1's gradient.  scripts/midk_anomaly_probe.py shows what that
    // is: the loss is bit-identical, fc1 / fc2 gradients stay at 3e-6, only conv3's tensors and everything below move -- one
    // conv3 output unit of that B = 8 batch has a pre-activation within rounding of zero, split-K sums its partials with fp32
    // atomics in a run-dependent order, and the unit takes either branch of relu'.  A property of atomics-based split-K meeting
    // that test vector, not a defect (the split products themselves: scripts/stress_split.py, 0 / 800 wrong).
    static const int on = [] { const char *e = getenv("ZOO_MIDK_SPLIT"); return e ? atoi(e) : 0; }();
    if ((on & 4) && BN < 64) return false;   // bisecting aids: bit 2 = only BN >= 64, bit 3 = only BN <= 32
    if ((on & 8) && BN > 32) return false;
    return (on >> (epi.mask ? 1 : 0)) & 1;
}

// tf32_passes: for fp32 operands 1 (plain tf32) or 3 (3xTF32); ignored for bf16 operands
zoo_status tc_gemm(const Operand &A, const Operand &B, int M, int N, int K, const Epilogue &epi, int split_k, float *ws,
                   size_t ws_elems, cudaStream_t st) {
    ZOO_REQUIRE(tc_gemm_supported(A, B, M, N, K), "tc_gemm: unsupported operands (M=%d N=%d K=%d)", M, N, K);
    const int eb = A.is_bf16 ? 2 : 4;
    const int bk = 128 / eb, mn_box = 128 / eb;
    const bool x3 = eb == 4 && epi.tf32_passes == 3;
    int BN = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
    // one M tile (skinny batch x wide layer): many narrow N tiles first, so fewer K splits have to be summed
    if (M <= BM && N >= 256) BN = 32;
    if (eb == 2 && BN == 32 && !B.kmajor) BN = 64;
    const int total_kb = cdiv(K, bk);
    // fp32-grade, shortish K, many tiles (IQN phi forward: K = 64, 15 616 tiles; fc1 data gradient: K = 512): a 128-wide tile takes all
    // 512 TMEM columns and 128 KB of shared memory, i.e. one CTA per SM with nothing to overlap its epilogue; 64-wide tiles fit two
    // per SM.  IQN 512 x 64 fp32: phi forward 1103 -> 780 us, fc1 dgrad 2770 -> 2408 us; no gain at K = 7744 (the main loop is
    // shared-memory-bandwidth bound there either way).  ZOO_X3_NARROW = longest K, in K-blocks of 32, that takes the narrow tile.
    static const int x3_narrow = [] { const char *e = getenv("ZOO_X3_NARROW"); return e ? atoi(e) : 16; }();
    static const int x3_narrow_q = [] { const char *e = getenv("ZOO_X3_NARROW_TILES_Q"); return e ? atoi(e) : 4; }();  // quarters of the SM count
    if (x3 && BN == 128 && total_kb <= x3_narrow && 4LL * cdiv(M, BM) * cdiv(N, 128) > (long long)x3_narrow_q * g_sm_budget) BN = 64;
    dim3 grid(cdiv(M, BM), cdiv(N, BN), 1);
    const long long tiles = (long long)grid.x * grid.y;
    static const int split_cap_pct = [] { const char *e = getenv("ZOO_SPLIT_CAP_PCT"); return e ? atoi(e) : 25; }();
    // (a quarter of the resident-CTA capacity: the update runs three lanes of kernels at once and every fp32-grade CTA owns an SM for
    // its whole life, so SM-time, not the latency of a kernel alone on the GPU, is what the step pays for -- fc1 forward of both nets
    // at 49 splits = 392 one-per-SM CTAs in 2.6 waves; at 18 splits both fit in one.  Measured on the B = 32 update: 100 % 0.2571 ms,
    // 50 % 0.2590, 25 % 0.2542)
    long long cap_ctas = g_sm_budget * ((x3 && BN >= 128) ? 1 : 2);  // resident CTAs (TcCfg::TWO_CTA_STAGES)
    if (split_k <= 0) {
        // split only when the output grid leaves most SMs idle AND one CTA would walk a long K: each split costs
        // atomics on the whole tile, so keep >= 4 K-blocks per CTA and stay within one resident wave
        split_k = 1;
        // (the reduced cap is for the skinny outputs of the Dopamine batch sizes -- a handful of tiles, K in the thousands; outputs
        // with more tiles keep the full cap: IQN B = 32 fc1, 64 tiles, still splits 4 ways)
        const long long split_ctas = tiles <= 8 ? cap_ctas * split_cap_pct / 100 : cap_ctas;
        if (tiles * 2 <= cap_ctas && total_kb >= 32) split_k = (int)std::min<long long>(total_kb / 4, std::max<long long>(1, split_ctas / tiles));
        else if (x3 && tiles * 2 <= cap_ctas && total_kb >= 16 && midk_split_enabled(epi, BN)) {
            // 3xTF32 K-blocks cost ~0.85 us each, a split ~7 us (BN = 32) to ~10 us (BN = 64) of reds + finish: measured
            // (scripts/sweep_split.py) it pays for 61 x BN32 tiles at 4 splits and for 31 x BN64 tiles at 6+, not for 61 x BN64 at 4
            const int s = (int)std::min<long long>(total_kb / 2, cap_ctas / tiles);
            if (s >= (BN <= 32 ? 3 : 6)) split_k = s;
        }
        if (split_k < 1) split_k = 1;
    }
    split_k = std::min(split_k, total_kb);
    int kbps = cdiv(total_kb, split_k);
    split_k = cdiv(total_kb, kbps);
    grid.z = split_k;

    TcParams P;
    P.M = M; P.N = N; P.K = K; P.kb_per_split = kbps; P.epi = epi; P.ws = nullptr; P.ctr = nullptr; P.dbg = g_tc_dbg;
    bool direct_acc = epi.accumulate && !epi.out_bf16;
    if (direct_acc && split_k == 1 && epi.accumulate == 2) P.epi.accumulate = 0;  // "overwrite or accumulate": nothing to sum
    if (split_k > 1 && !direct_acc) {
        ZOO_REQUIRE(ws && ws_elems >= (size_t)M * N + TC_WS_COUNTERS && tiles <= TC_WS_COUNTERS,
                    "tc_gemm: split-K workspace too small (%zu < %lld)", ws_elems, (long long)M * N + TC_WS_COUNTERS);
        P.ws = ws;  // invariant: all-zero on entry, left all-zero by the finishing CTAs
        P.ctr = reinterpret_cast<unsigned *>(ws + ws_elems - TC_WS_COUNTERS);
    }
    {
        // TMA-store epilogue for plain bf16 row-major outputs (see TcParams::tma_store); ZOO_TMA_STORE=0 falls back to the
        // per-row 128-byte stores
        static const bool ts_on = [] { const char *e = getenv("ZOO_TMA_STORE"); return !(e && e[0] == '0'); }();
        const Epilogue &pe = P.epi;
        const bool rm_ok = !pe.rowmul || ((pe.rowmul_div & 31) == 0);
        const bool mask_ok = !pe.mask || (pe.mask_bf16 && pe.mask_sn == 1 && (pe.mask_sm & 7) == 0 && ((uintptr_t)pe.mask & 15) == 0);
        P.tma_store = 0;
        if (ts_on && !x3 && BN >= 64 && split_k == 1 && pe.out_bf16 && !pe.accumulate && pe.bias_mode != 2 && pe.sn == 1 && (pe.sm & 7) == 0 &&
            ((uintptr_t)pe.out & 15) == 0 && rm_ok && mask_ok) {
            ZOO_TRY(make_map(&P.tmC, pe.out, 2, N, M, pe.sm, 64, 32, false));
            P.tma_store = 1;
        }
        // fp32-grade (3xTF32) kernels with a plain fp32 row-major output: same idea, 32-column boxes (tma_store == 3).  The phi forward
        // of IQN at 512 x 64 rows (1 GB of fp32 output, K = 64) spent most of its 1.19 ms in the slab + per-row stores.
        static const long long x3_ts_min = [] { const char *e = getenv("ZOO_X3_TS_MIN"); return e ? atoll(e) : (1ll << 21); }();
        const bool mask32_ok = !pe.mask || (!pe.mask_bf16 && pe.mask_sn == 1 && (pe.mask_sm & 3) == 0 && ((uintptr_t)pe.mask & 15) == 0);
        if (ts_on && x3 && BN >= 64 && split_k == 1 && !pe.out_bf16 && !pe.accumulate && pe.bias_mode != 2 && pe.sn == 1 && (pe.sm & 3) == 0 &&
            ((uintptr_t)pe.out & 15) == 0 && rm_ok && mask32_ok && (long long)M * N >= x3_ts_min) {
            ZOO_TRY(make_map(&P.tmC, pe.out, 4, N, M, pe.sm, 32, 32, false));
            P.tma_store = 3;
        }
    }
    CUtensorMap tmA, tmB;
    if (A.kmajor) ZOO_TRY(make_map(&tmA, A.p, eb, K, M, A.ld, bk, BM, false));
    else ZOO_TRY(make_map(&tmA, A.p, eb, M, K, A.ld, mn_box, bk, true));
    if (B.kmajor) ZOO_TRY(make_map(&tmB, B.p, eb, K, N, B.ld, bk, BN, false));
    else ZOO_TRY(make_map(&tmB, B.p, eb, N, K, B.ld, mn_box, bk, true));

    P.dbg = trace_slot("gemm", M, N, K, grid);
    ZOO_TRY(tc_launch(eb, x3, 0, BN, A.kmajor, B.kmajor, tmA, tmB, P, grid, st));
    return ZOO_OK;
}

// ---------------------------------------------------------------- implicit-GEMM convolution --------
bool tc_conv_supported(const ConvGeom &g, int is_bf16) {
    const int eb = is_bf16 ? 2 : 4, ch = 16 / eb, bk = 128 / eb;
    if (((uintptr_t)g.x & 15) && g.mode != CONV_RAW_U8) return false;
    if (g.mode == CONV_RAW_U8 || g.mode == CONV_RAW_F32) return g.ks % ch == 0 && ((long long)g.C * g.ks * g.ks * eb) % 16 == 0;
    if (g.mode == CONV_NHWC) return g.C % ch == 0;
    if (g.mode == CONV_DGRAD) return g.Cout % bk == 0 && g.C % 8 == 0 && ((long long)g.ks * g.ks * g.C * eb) % 16 == 0;
    if (g.mode == CONV_TMA_DGRAD) {  // rows = the pixels of one residue class of the input grid; boxes over dY [B][OH][OW][Cout]
        if (g.s < 1 || g.s > 2 || g.W < 1 || g.H < 1) return false;  // s*s <= 4 class maps (TcParams::tmC4)
        const int cols = cdiv(g.W, g.s), rows = cdiv(g.H, g.s);
        if (cols > 128) return false;
        const int R = std::min(rows, 128 / cols);
        // dY channel blocks of 128 bytes (K-major A), C <= 64 output columns whose rows are whole 16-byte chunks, W boxes of MN_BOX
        return R >= 1 && R <= 256 && cols <= 256 && (g.Cout * eb) % 128 == 0 && g.C <= 64 && (g.C * eb) % 64 == 0 && g.ks >= g.s;
    }
    if (g.mode == CONV_TMA) {
        if (g.OW > 128 || g.OW < 1) return false;
        // the innermost contiguous run (one kernel row of the padded frame stack, or one pixel's channels) must be a whole number
        // of 128-byte spans -- or, bf16 only, exactly one 64-byte span (SWIZZLE_64B variant; N <= 64)
        const int run = (g.first ? g.ks * g.C : g.C) * eb;
        const bool span_ok = g.first ? (run == 128 || (eb == 2 && run == 64 && g.Cout <= 64)) : (run % 128 == 0 || (eb == 2 && run == 64 && g.Cout <= 64));
        if (g.first) return span_ok && g.p == 0 && (g.s * g.C * eb) % 16 == 0 && g.OW <= 256 && ((128 / g.OW - 1) * g.s + 1) <= 256;
        const int R = std::min(g.OH, 128 / g.OW);
        return span_ok && (g.OW - 1) * g.s + 1 <= 256 && (R - 1) * g.s + 1 <= 256;
    }
    return false;
}

// forward conv with TMA-fetched im2col rows.  g.x: NHWC activations [images][H][W][C] (for `first`: already zero-padded, p = 0)
static zoo_status tc_conv_tma(ConvGeom g, const Operand &Wop, int images, int N, int K, const Epilogue &epi, float *ws, size_t ws_elems,
                              cudaStream_t st) {
    const int eb = Wop.is_bf16 ? 2 : 4;
    const int span = ((g.first ? g.ks * g.C : g.C) * eb) % 128 == 0 ? 128 : 64;
    const int bk = span / eb;
    const bool x3 = eb == 4 && epi.tf32_passes == 3;
    ZOO_REQUIRE(Wop.kmajor && K % bk == 0, "tc_conv_tma: K %d is not a whole number of K-blocks", K);
    int BN = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
    g.R = std::min(g.OH, 128 / g.OW);
    g.tpi = cdiv(g.OH, g.R);
    g.cblocks = g.first ? 1 : (g.C * eb) / span;
    dim3 grid(images * g.tpi, cdiv(N, BN), 1);
    TcParams P;
    P.M = images * g.OH * g.OW; P.N = N; P.K = K; P.kb_per_split = K / bk; P.epi = epi; P.ws = nullptr; P.ctr = nullptr; P.dbg = nullptr; P.conv = g;
    P.span = span;
    if (P.epi.accumulate == 2) P.epi.accumulate = 0;
    // optional split over kernel taps (ZOO_CONV_SPLIT=n).  Measured at B = 32 (one tile per image for the 9 x 9 / 7 x 7 maps, 32-64 CTAs):
    // 2 / 4 / 8 splits cost 8 / 12 / 22 % of the whole update -- the extra CTAs and reds take the SMs the other two lanes were
    // using -- so it stays off.
    static const int conv_split = [] { const char *e = getenv("ZOO_CONV_SPLIT"); return e ? atoi(e) : 0; }();
    const long long tiles = (long long)grid.x * grid.y;
    const int total_kb = K / bk;
    const int split = conv_split > 0 ? conv_split : 1;
    const size_t need = (size_t)P.M * N + TC_WS_COUNTERS;
    if (split > 1 && !P.epi.accumulate && ws && ws_elems >= need && tiles <= TC_WS_COUNTERS) {
        const int kbps = cdiv(total_kb, std::min(split, total_kb));
        grid.z = cdiv(total_kb, kbps);
        if (grid.z > 1) {
            P.kb_per_split = kbps;
            P.ws = ws;
            P.ctr = reinterpret_cast<unsigned *>(ws + ws_elems - TC_WS_COUNTERS);
        }
    }
    CUtensorMap tmA, tmB;
    const cuuint64_t sC = (cuuint64_t)g.C * eb, sW = (cuuint64_t)g.W * sC, sH = (cuuint64_t)g.H * sW;
    if (g.first) {
        // view of the padded frame stack: d0 = one kernel row (ks taps x C channels, contiguous), d1 = output column (overlapping
        // windows, stride s pixels), d2 = input row (the box steps s rows per output row), d3 = image
        const cuuint64_t dims[4] = {(cuuint64_t)(g.ks * g.C), (cuuint64_t)g.OW, (cuuint64_t)g.H, (cuuint64_t)images};
        const cuuint64_t str[3] = {(cuuint64_t)g.s * sC, sW, sH};
        const cuuint32_t box[4] = {(cuuint32_t)(g.ks * g.C), (cuuint32_t)g.OW, (cuuint32_t)((g.R - 1) * g.s + 1), 1};
        const cuuint32_t es[4] = {1, 1, (cuuint32_t)g.s, 1};
        ZOO_TRY(make_map4(&tmA, g.x, eb, dims, str, box, es, span));
    } else {
        const cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)images};
        const cuuint64_t str[3] = {sC, sW, sH};
        const cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)((g.OW - 1) * g.s + 1), (cuuint32_t)((g.R - 1) * g.s + 1), 1};
        const cuuint32_t es[4] = {1, (cuuint32_t)g.s, (cuuint32_t)g.s, 1};
        ZOO_TRY(make_map4(&tmA, g.x, eb, dims, str, box, es, span));
    }
    ZOO_TRY(make_map(&tmB, Wop.p, eb, K, N, Wop.ld, bk, BN, false, span));
    P.dbg = trace_slot("conv", P.M, N, K, grid);
    return tc_launch(eb, x3, 2, BN, true, true, tmA, tmB, P, grid, st);
}

// 3-D tensor map {d0, d1, d2} with explicit swizzle (the dY operand of the weight gradient)
static zoo_status make_map3(CUtensorMap *tm, const void *p, int eb, const cuuint64_t dims[3], const cuuint64_t strides_bytes[2], const cuuint32_t box[3],
                            CUtensorMapSwizzle sw) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_err("cuTensorMapEncodeTiled not available from the driver"); return ZOO_CUDA_ERROR; }
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(p), dims, strides_bytes, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_err("cuTensorMapEncodeTiled(3D) failed (%d): dims %llu %llu %llu box %u %u %u", (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
                (unsigned long long)dims[2], box[0], box[1], box[2]);
        return ZOO_CUDA_ERROR;
    }
    return ZOO_OK;
}

static int wgrad_span(const ConvGeom &g, int eb) {
    const int run = (g.first ? g.ks * g.C : g.C) * eb;
    if (run % 128 == 0) return 128;
    return (eb == 2 && run == 64) ? 64 : 0;
}

bool tc_conv_wgrad_supported(const ConvGeom &g, int is_bf16) {
    const int eb = is_bf16 ? 2 : 4, span = wgrad_span(g, eb);
    if (!span) return false;
    if (g.first && !(g.ks * g.C * eb == span && g.p == 0)) return false;
    const int mn_rows = 128 / eb, mn_box = span / eb;
    if (g.OW < 1 || g.OW > mn_rows || g.Cout > BM || g.Cout < 8) return false;
    const int R = std::min(g.OH, mn_rows / g.OW);
    if ((g.OW - 1) * g.s + 1 > 256 || (R - 1) * g.s + 1 > 256) return false;
    if ((g.ks * g.ks * g.C) % mn_box) return false;
    if ((g.Cout * eb) % 16 || (g.s * g.C * eb) % 16 || ((uintptr_t)g.x & 15)) return false;
    return true;
}

// dW[Cout][ks*ks*C] += dY^T im2col(X) with both operands fetched as TMA boxes (kernel GA = 4): g.x = NHWC activations the forward
// conv read (`first`: the zero-padded frame stack, p = 0), dy = [images][OH][OW][Cout] in the same element type; epi.out = fp32
// [Cout][ks*ks*C], all-zero on entry (partial sums of the image splits are red-added).
zoo_status tc_conv_wgrad(ConvGeom g, const void *dy, int is_bf16, int images, const Epilogue &epi, cudaStream_t st) {
    ZOO_REQUIRE(tc_conv_wgrad_supported(g, is_bf16) && images > 0 && ((uintptr_t)dy & 15) == 0 && !epi.out_bf16, "tc_conv_wgrad: unsupported geometry");
    const int eb = is_bf16 ? 2 : 4, span = wgrad_span(g, eb);
    const int mn_rows = 128 / eb, mn_box = span / eb, umma_k = 32 / eb;
    const bool x3 = eb == 4 && epi.tf32_passes == 3;
    g.R = std::min(g.OH, mn_rows / g.OW);
    g.tpi = cdiv(g.OH, g.R);
    g.ksteps = cdiv(g.R * g.OW, umma_k);
    g.mode = CONV_TMA_WGRAD;
    const int N = g.ks * g.ks * g.C, BN = 128;
    const int total_kb = images * g.tpi;
    const int ntiles = cdiv(N, BN);
    static const int wgrad_cap_pct = [] { const char *e = getenv("ZOO_WGRAD_CAP_PCT"); return e ? atoi(e) : 100; }();
    const long long cap = g_sm_budget * (x3 ? 1 : 2) * (images <= 128 ? wgrad_cap_pct : 100) / 100;  // (large batches keep the full cap)
    int split = (int)std::max<long long>(1, std::min<long long>(total_kb, cap / ntiles));
    const int kbps = cdiv(total_kb, split);
    split = cdiv(total_kb, kbps);
    dim3 grid(1, ntiles, split);
    TcParams P;
    P.M = g.Cout; P.N = N; P.K = total_kb; P.kb_per_split = kbps; P.epi = epi; P.ws = nullptr; P.ctr = nullptr; P.dbg = nullptr; P.conv = g; P.span = span;
    P.epi.sm = N; P.epi.sn = 1; P.epi.out_bf16 = 0; P.epi.bias_mode = 0; P.epi.relu = 0; P.epi.mask = nullptr; P.epi.rowmul = nullptr;
    P.epi.accumulate = split > 1 ? 1 : 0;
    const CUtensorMapSwizzle sw = eb == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : (span == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    CUtensorMap tmA, tmB;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)g.Cout, (cuuint64_t)(g.OH * g.OW), (cuuint64_t)images};
        const cuuint64_t str[2] = {(cuuint64_t)g.Cout * eb, (cuuint64_t)g.OH * g.OW * g.Cout * eb};
        const cuuint32_t box[3] = {(cuuint32_t)mn_box, (cuuint32_t)(g.R * g.OW), 1};
        ZOO_TRY(make_map3(&tmA, dy, eb, dims, str, box, sw));
    }
    {
        EncodeTiledFn enc = get_encode();
        ZOO_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
        const cuuint64_t sC = (cuuint64_t)g.C * eb, sW = (cuuint64_t)g.W * sC, sH = (cuuint64_t)g.H * sW;
        cuuint64_t dims[4], str[3];
        cuuint32_t box[4], es[4];
        if (g.first) {
            dims[0] = (cuuint64_t)(g.ks * g.C); dims[1] = (cuuint64_t)g.OW; dims[2] = (cuuint64_t)g.H; dims[3] = (cuuint64_t)images;
            str[0] = (cuuint64_t)g.s * sC; str[1] = sW; str[2] = sH;
            box[0] = (cuuint32_t)(g.ks * g.C); box[1] = (cuuint32_t)g.OW; box[2] = (cuuint32_t)((g.R - 1) * g.s + 1); box[3] = 1;
            es[0] = 1; es[1] = 1; es[2] = (cuuint32_t)g.s; es[3] = 1;
        } else {
            dims[0] = (cuuint64_t)g.C; dims[1] = (cuuint64_t)g.W; dims[2] = (cuuint64_t)g.H; dims[3] = (cuuint64_t)images;
            str[0] = sC; str[1] = sW; str[2] = sH;
            box[0] = (cuuint32_t)mn_box; box[1] = (cuuint32_t)((g.OW - 1) * g.s + 1); box[2] = (cuuint32_t)((g.R - 1) * g.s + 1); box[3] = 1;
            es[0] = 1; es[1] = (cuuint32_t)g.s; es[2] = (cuuint32_t)g.s; es[3] = 1;
        }
        CUresult r = enc(&tmB, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void *>(g.x), dims, str, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_err("tc_conv_wgrad: cuTensorMapEncodeTiled(4D) failed (%d)", (int)r); return ZOO_CUDA_ERROR; }
    }
    P.dbg = trace_slot("cwgrad", P.M, N, total_kb, grid);
    return tc_launch(eb, x3, 4, BN, false, false, tmA, tmB, P, grid, st);
}

zoo_status tc_conv_gemm(const ConvGeom &g, const Operand &Wop, int M, int N, int K, const Epilogue &epi, float *ws, size_t ws_elems,
                        cudaStream_t st) {
    ZOO_REQUIRE(tc_conv_supported(g, Wop.is_bf16) && M > 0 && N >= 8 && ((uintptr_t)Wop.p & 15) == 0, "tc_conv_gemm: unsupported geometry");
    if (g.mode == CONV_TMA) return tc_conv_tma(g, Wop, M / (g.OH * g.OW), N, K, epi, ws, ws_elems, st);
    if (g.mode == CONV_TMA_DGRAD) {
        // data gradient by sub-pixel decomposition (kernel: GA = 2 with mode CONV_TMA_DGRAD): grid.z = the s*s residue classes
        const int images = M / (g.H * g.W);
        const int eb = Wop.is_bf16 ? 2 : 4, bk = 128 / eb, mn_box = 128 / eb;
        const bool x3 = eb == 4 && epi.tf32_passes == 3;
        ZOO_REQUIRE(!Wop.kmajor && K == g.ks * g.ks * g.Cout && N == g.C && epi.out && ((uintptr_t)epi.out & 15) == 0 && (epi.out_bf16 != 0) == (eb == 2),
                    "tc_conv_gemm: bad operands for the TMA data gradient");
        ConvGeom d = g;
        const int cols = cdiv(g.W, g.s), rows = cdiv(g.H, g.s);
        d.R = std::min(rows, 128 / cols); d.tpi = cdiv(rows, d.R); d.cblocks = (g.Cout * eb) / 128; d.first = 0;
        const int BN = eb == 2 ? 64 : (N <= 32 ? 32 : 64);  // bf16 MN-major W boxes are 64 columns wide (N = 32: the upper half is never stored)
        dim3 grid(images * d.tpi, 1, g.s * g.s);
        TcParams P;
        const int kt = cdiv(g.ks, g.s);
        P.M = M; P.N = N; P.K = K; P.kb_per_split = kt * kt * d.cblocks; P.epi = epi; P.ws = nullptr; P.ctr = nullptr; P.dbg = nullptr; P.conv = d;
        P.tpi_max = d.tpi; P.tma_store = 2;
        const int esz = eb;
        P.ts_span = std::min(128, N * esz);
        const CUtensorMapSwizzle osw = P.ts_span == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (P.ts_span == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        EncodeTiledFn enc = get_encode();
        ZOO_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
        for (int py = 0; py < g.s; ++py)
            for (int px = 0; px < g.s; ++px) {  // the class's strided view of dX [images][H][W][C]
                const int rows_c = (g.H - py + g.s - 1) / g.s, cols_c = (g.W - px + g.s - 1) / g.s;
                const cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)cols_c, (cuuint64_t)rows_c, (cuuint64_t)images};
                const cuuint64_t str[3] = {(cuuint64_t)g.s * g.C * esz, (cuuint64_t)g.s * g.W * g.C * esz, (cuuint64_t)g.H * g.W * g.C * esz};
                const cuuint32_t box[4] = {(cuuint32_t)(P.ts_span / esz), (cuuint32_t)cols, (cuuint32_t)d.R, 1};
                const cuuint32_t es[4] = {1, 1, 1, 1};
                void *base = (uint8_t *)epi.out + ((size_t)py * g.W + px) * g.C * esz;
                CUresult r = enc(&P.tmC4[py * g.s + px], eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, str, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, osw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { set_err("tc_conv_gemm: output map of class (%d,%d) failed (%d)", py, px, (int)r); return ZOO_CUDA_ERROR; }
            }
        CUtensorMap tmA, tmB;
        const cuuint64_t sC = (cuuint64_t)g.Cout * eb, sW = (cuuint64_t)g.OW * sC, sH = (cuuint64_t)g.OH * sW;
        const cuuint64_t dims[4] = {(cuuint64_t)g.Cout, (cuuint64_t)g.OW, (cuuint64_t)g.OH, (cuuint64_t)images};
        const cuuint64_t str[3] = {sC, sW, sH};
        const cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)cols, (cuuint32_t)d.R, 1};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        ZOO_TRY(make_map4(&tmA, g.x, eb, dims, str, box, es));
        ZOO_TRY(make_map(&tmB, Wop.p, eb, Wop.ld, g.Cout, Wop.ld, mn_box, bk, true));  // W as [Cout][ks*ks*C]: boxes at tap offsets
        P.dbg = trace_slot("cdgrad", M, N, K, grid);
        return tc_launch(eb, x3, 3, BN, true, false, tmA, tmB, P, grid, st);
    }
    const int eb = Wop.is_bf16 ? 2 : 4;
    const int bk = 128 / eb, mn_box = 128 / eb;
    const bool x3 = eb == 4 && epi.tf32_passes == 3;
    ZOO_REQUIRE(g.mode == CONV_DGRAD ? !Wop.kmajor : Wop.kmajor, "tc_conv_gemm: weight operand has the wrong major");
    int BN = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
    if (eb == 2 && BN == 32 && !Wop.kmajor) BN = 64;
    dim3 grid(cdiv(M, BM), cdiv(N, BN), 1);
    const int total_kb = cdiv(K, bk);
    const long long tiles = (long long)grid.x * grid.y;
    const long long cap_ctas = g_sm_budget * ((x3 && BN >= 128) ? 1 : 2);
    int split_k = 1;
    if (tiles * 2 <= cap_ctas && total_kb >= 32) split_k = (int)std::min<long long>(total_kb / 4, cap_ctas / tiles);
    split_k = std::max(1, std::min(split_k, total_kb));
    int kbps = cdiv(total_kb, split_k);
    split_k = cdiv(total_kb, kbps);
    grid.z = split_k;
    TcParams P;
    P.M = M; P.N = N; P.K = K; P.kb_per_split = kbps; P.epi = epi; P.ws = nullptr; P.ctr = nullptr; P.dbg = nullptr; P.conv = g;
    const bool direct_acc = epi.accumulate && !epi.out_bf16;
    if (direct_acc && split_k == 1 && epi.accumulate == 2) P.epi.accumulate = 0;
    if (split_k > 1 && !direct_acc) {
        ZOO_REQUIRE(ws && ws_elems >= (size_t)M * N + TC_WS_COUNTERS && tiles <= TC_WS_COUNTERS, "tc_conv_gemm: split-K workspace too small");
        P.ws = ws;
        P.ctr = reinterpret_cast<unsigned *>(ws + ws_elems - TC_WS_COUNTERS);
    }
    CUtensorMap tmB;
    if (Wop.kmajor) ZOO_TRY(make_map(&tmB, Wop.p, eb, K, N, Wop.ld, bk, BN, false));
    else ZOO_TRY(make_map(&tmB, Wop.p, eb, Wop.ld, g.Cout, Wop.ld, mn_box, bk, true));  // W as [Cout][ks*ks*C]: boxes at tap offsets
    return tc_launch(eb, x3, 1, BN, true, Wop.kmajor, tmB, tmB, P, grid, st);
}

}  // namespace zoo
