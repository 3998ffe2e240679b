// Shared declarations of libzoo_sm100 (sm_100a only; no other architecture is compiled).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include "../../include/zoo_sm100.h"

namespace zoo {

typedef __nv_bfloat16 bf16;

void set_err(const char *fmt, ...);
extern thread_local int g_launches;  // kernels enqueued since the counter was last reset

#define ZOO_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            zoo::set_err("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return ZOO_CUDA_ERROR;                                                            \
        }                                                                                     \
    } while (0)

#define ZOO_TRY(call)                    \
    do {                                 \
        zoo_status s__ = (call);         \
        if (s__ != ZOO_OK) return s__;   \
    } while (0)

#define ZOO_REQUIRE(cond, ...)           \
    do {                                 \
        if (!(cond)) {                   \
            zoo::set_err(__VA_ARGS__);   \
            return ZOO_BAD_ARG;          \
        }                                \
    } while (0)

// Built-in launch profiler (zoo_prof_*): when armed, an event is recorded after every kernel launch on the profiled
// stream, so consecutive differences give warm, in-situ per-kernel durations (bench.py's roofline leg).
extern thread_local bool g_prof_on;
void prof_mark(const char *func, int line);
void prof_label(const char *fmt, ...);

#define ZOO_LAUNCH_CHECK()                                                                   \
    do {                                                                                     \
        zoo::g_launches++;                                                                   \
        if (zoo::g_prof_on) zoo::prof_mark(__func__, __LINE__);                              \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess) {                                                            \
            zoo::set_err("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return ZOO_CUDA_ERROR;                                                           \
        }                                                                                    \
    } while (0)

zoo_status require_sm100();
// SMs the GEMM split-K heuristics may plan for: 148, or fewer while a data-parallel communicator is alive (the NCCL kernel
// of the overlapped late bucket holds SMs; a 148-CTA one-CTA-per-SM GEMM would then fall into a second wave)
extern int g_sm_budget;
// set around launches that run beside the critical path (the target net's forward on the side lane): the engine then sizes its
// ring so that two CTAs share an SM, leaving SMs to the lane that matters instead of owning 128 of them for the kernel's whole life
extern thread_local int g_tc_shallow;
extern int g_tc_stages_force;  // 0 = per-launch choice (tc_pick_stages); >= 2 forces the ring depth (tuning hook)

// Programmatic dependent launch: a kernel launched through launch_pdl may become resident while its stream predecessor is
// still draining, and runs its prologue (barrier init, TMEM allocation, index setup) there; it must call pdl_wait() before it
// touches global memory, and calls pdl_trigger() once its main loop is done (triggering at kernel entry was measured slower:
// early-resident dependents take SMs from the predecessor's second wave).  +4 % on the B=32 update; ZOO_PDL=0 disables it.
bool pdl_enabled();
// timeline profiler (zoo_prof_timeline): with the lanes running concurrently, every launch_pdl launch is bracketed by two
// events on its own stream; zoo_prof_timeline_get returns (label, stream, start, end) relative to the first launch
extern thread_local bool g_timeline_on;
void timeline_mark(cudaStream_t st, int end);
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    if (g_timeline_on) {
        timeline_mark(st, 0);
        cudaLaunchConfig_t c0 = {};
        c0.gridDim = grid; c0.blockDim = block; c0.dynamicSmemBytes = smem; c0.stream = st;
        cudaError_t e = cudaLaunchKernelEx(&c0, kernel, std::forward<Args>(args)...);
        timeline_mark(st, 1);
        return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- GEMM engines ------------
// D(m,n) = sum_k A(m,k) * B(n,k)
struct Operand {
    const void *p;
    int is_bf16;   // element type: 1 = bf16, 0 = fp32
    long long ld;  // leading dimension in elements
    int kmajor;    // 1: (i,k) at p[i*ld + k];  0: (i,k) at p[k*ld + i]
};

struct Epilogue {
    void *out = nullptr;
    int out_bf16 = 0;
    long long sm = 0, sn = 1;  // D(m,n) stored at out[m*sm + n*sn]
    const float *bias = nullptr;
    int bias_mode = 0;  // 0 none, 1 per-n, 2 per-m
    int relu = 0;
    const void *mask = nullptr;  // multiply by (mask(m,n) > 0): relu' of the layer this gradient enters
    int mask_bf16 = 0;
    long long mask_sm = 0, mask_sn = 1;
    const void *rowmul = nullptr;  // multiply by rowmul[(m / rowmul_div)][n] (IQN Hadamard merge), same dtype as out
    int rowmul_div = 1;
    long long rowmul_ld = 0;
    int rowmul_nonneg = 0;  // the multiplier is a relu output (>= 0): relu(x) * rm may be evaluated as relu(x * rm)
    int accumulate = 0;  // 1: atomically add into fp32 out (out must be fp32; bias/relu/mask not allowed);
                         // 2: out is pre-zeroed fp32: plain store when the GEMM is not K-split, atomic add when it is
    int tf32_passes = 0; // fp32 operands on tcgen05: 0 = never (CUDA cores), 1 = plain tf32, 3 = 3xTF32 (fp32-grade)
};

// Implicit-GEMM convolution: the A operand (im2col matrix) is never materialised; the tcgen05 kernel's producer warps
// gather each 128-row x 128-byte K-block straight from the activation tensor into the swizzled shared-memory tile.
enum { CONV_RAW_U8 = 1,   // first conv: observation [B][C][H][W] uint8, K order (c,ky,kx), optional x/255 (IEEE divide)
       CONV_RAW_F32 = 2,  // first conv: observation [B][C][H][W] float
       CONV_NHWC = 3,     // inner conv: activations [B][H][W][C] in the GEMM dtype, K order (ky,kx,c)
       CONV_DGRAD = 4,    // data gradient: A = dY [B][OH][OW][Cout] gathered transposed-conv style, K order (ky,kx,o),
                          //   M = B*H*W input pixels, N = C;  B operand = W viewed [Cout][ks*ks*C] (MN-major)
       CONV_TMA = 5,      // forward conv on NHWC activations with TMA-fetched im2col rows (4-D boxes), K order (ky,kx,c)
       CONV_TMA_WGRAD = 7,    // weight gradient: contraction over the output pixels, dY and the activations as MN-major TMA boxes of
                          //   R whole output rows of one image (see gemm_tc.cuh, GA = 4)
       CONV_TMA_DGRAD = 6 };  // data gradient the same way, any stride (sub-pixel decomposition): grid.z = the s*s residue classes
                          //   (py, px) of the input pixel grid; class pixels (s*yy + py, s*xx + px) only receive the taps
                          //   ky = (py + p) % s + a*s, i.e. a stride-1 convolution of dY with the class's sub-kernel: A = dY boxes,
                          //   K order (a, b, o), M = class pixels (tiles = whole class rows of one image), N = C, W operand as for
                          //   CONV_DGRAD; the tile goes back to the strided class view of dX with one TMA store
struct ConvGeom {
    const void *x = nullptr;
    int mode = 0;
    int C = 0, H = 0, W = 0;      // input channels / spatial size (the tensor the forward conv reads)
    int ks = 0, s = 1, p = 0;     // kernel, stride, symmetric zero pad
    int OH = 0, OW = 0, Cout = 0; // forward output size / channels
    int scale255 = 0;
    // CONV_TMA: M tiles are R whole output rows of one image (tpi tiles per image); one K-block is one tap x one 128-byte
    // channel block (cblocks per tap), or -- `first`: zero-padded 4-channel frame stack -- one kernel row (ks taps x C)
    int tpi = 0, R = 0, cblocks = 1, first = 0;
    int ksteps = 4;  // CONV_TMA_WGRAD: tcgen05.mma per K-block = ceil(R * OW / UMMA_K)
};
bool tc_conv_supported(const ConvGeom &g, int is_bf16);
// weight gradient of a convolution with TMA-fetched operands (no im2col matrix): dW[Cout][ks*ks*C] += dY^T im2col(X)
bool tc_conv_wgrad_supported(const ConvGeom &g, int is_bf16);
zoo_status tc_conv_wgrad(ConvGeom g, const void *dy, int is_bf16, int images, const Epilogue &epi, cudaStream_t st);
// D[M][N] = sum_k A(m,k) W(n,k) with A gathered per ConvGeom; Wop: K-major [N][K] (forward) or MN-major (dgrad)
zoo_status tc_conv_gemm(const ConvGeom &g, const Operand &Wop, int M, int N, int K, const Epilogue &epi, float *ws, size_t ws_elems,
                        cudaStream_t st);

// fp32 CUDA-core engine (any shapes, fp32 or bf16 inputs, fp32 accumulate)
zoo_status simt_gemm(const Operand &A, const Operand &B, int M, int N, int K, const Epilogue &epi, int split_k,
                     float *ws, size_t ws_elems, cudaStream_t st);
// tcgen05 engine (bf16 inputs).  tc_gemm_supported tells whether the shapes/alignments qualify.
bool tc_gemm_supported(const Operand &A, const Operand &B, int M, int N, int K);
zoo_status tc_gemm(const Operand &A, const Operand &B, int M, int N, int K, const Epilogue &epi, int split_k,
                   float *ws, size_t ws_elems, cudaStream_t st);
// persistent tcgen05 engine for the large-batch bf16 contractions (gemm_persist.cu): 128 x 256 tiles, one CTA per SM, double-buffered
// TMEM accumulators, TMA-store epilogue
bool tc_gemm_persist_ok(const Operand &A, const Operand &B, int M, int N, int K, const Epilogue &epi);
zoo_status tc_gemm_persist(const Operand &A, const Operand &B, int M, int N, int K, const Epilogue &epi, cudaStream_t st);
// dispatch: tcgen05 when supported (bf16 x bf16, or fp32 x fp32 with epi.tf32_passes > 0), otherwise SIMT
zoo_status gemm(const Operand &A, const Operand &B, int M, int N, int K, const Epilogue &epi, int split_k, float *ws,
                size_t ws_elems, cudaStream_t st);

__device__ __forceinline__ float ld_elem(const void *p, int is_bf16, long long i) {
    return is_bf16 ? __bfloat162float(((const bf16 *)p)[i]) : ((const float *)p)[i];
}
// read-only (non-coherent) variant: lets the compiler batch these loads ahead of unrelated global stores
__device__ __forceinline__ float ldg_elem(const void *p, int is_bf16, long long i) {
    return is_bf16 ? __bfloat162float(__ldg(((const bf16 *)p) + i)) : __ldg(((const float *)p) + i);
}
__device__ __forceinline__ void st_elem(void *p, int is_bf16, long long i, float v) {
    if (is_bf16) ((bf16 *)p)[i] = __float2bfloat16_rn(v);
    else ((float *)p)[i] = v;
}

// applies everything of the epilogue except the store decision; shared by both engines + finish kernel
__device__ __forceinline__ float epi_apply(const Epilogue &e, float v, int m, int n) {
    if (e.bias_mode == 1) v += e.bias[n];
    else if (e.bias_mode == 2) v += e.bias[m];
    if (e.relu) v = fmaxf(v, 0.f);
    if (e.mask) v = ld_elem(e.mask, e.mask_bf16, (long long)m * e.mask_sm + (long long)n * e.mask_sn) > 0.f ? v : 0.f;
    if (e.rowmul) v *= ld_elem(e.rowmul, e.out_bf16, (long long)(m / e.rowmul_div) * e.rowmul_ld + n);
    return v;
}
__device__ __forceinline__ void epi_store(const Epilogue &e, float v, int m, int n) {
    long long o = (long long)m * e.sm + (long long)n * e.sn;
    if (e.accumulate) atomicAdd((float *)e.out + o, v);
    else st_elem(e.out, e.out_bf16, o, epi_apply(e, v, m, n));
}

}  // namespace zoo
