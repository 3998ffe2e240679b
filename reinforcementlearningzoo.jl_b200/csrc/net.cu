// Q-network engine (see net.cuh).  Activations are NHWC internally ([rows*OH*OW][C]), which makes every conv a
// GEMM over an im2col matrix and makes reshape(x,:,B) before the first Dense a no-op; the feature permutation this
// implies relative to the reference's (c,h,w) flatten order is applied once, when parameters cross the C ABI.
#include "net.cuh"

namespace zoo {
void dp_free_registered(void *ptr);  // dp.cu
// narrow.cu: the Q head (n_actions <= 8 outputs) as one streaming pass forward and one backward
bool narrow_supported(int N, long long K, int x_bf16);
zoo_status narrow_forward(const void *X, int x_bf16, const float *W, const float *bias, float *out, long long M, int N, long long K, cudaStream_t st);
zoo_status narrow_backward(const void *X, int x_bf16, const float *dY, const float *W, void *dX, float *dW, float *db, float *db_prev, long long M,
                           int N, long long K, int relu_prev, cudaStream_t st);
static bool narrow_enabled() {
    static const bool on = [] { const char *e = getenv("ZOO_NARROW"); return !(e && e[0] == '0'); }();
    return on;
}

// Implicit-GEMM convolutions (tc_conv_gemm: the A operand is gathered in-kernel, no im2col / col2im) are parity-tested but
// this round still lose to explicit im2col + TMA GEMM at these layer sizes (the 4-warp gather is instruction-latency bound,
// see DESIGN.md section 4), so they are opt-in: ZOO_IMPLICIT_CONV=1.
// forward convs with TMA-fetched im2col rows (CONV_TMA) are the default where the layer geometry allows; ZOO_TMA_CONV=0 disables
static bool tma_dgrad_enabled() {
    // data gradients of the convolutions by sub-pixel decomposition with TMA-fetched dY boxes and a TMA store into the strided
    // class view of dX: no dcols matrix, no col2im pass.  ZOO_TMA_DGRAD=0 restores GEMM + col2im.
    static const bool on = [] { const char *e = getenv("ZOO_TMA_DGRAD"); return !(e && e[0] == '0'); }();
    return on;
}
static bool tma_conv_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("ZOO_TMA_CONV"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}
// weight gradients of the convolutions with TMA-fetched operands (no im2col matrix at all); ZOO_TMA_WGRAD=0 restores the explicit
// im2col + GEMM path
static bool tma_wgrad_enabled() {
    static const bool on = [] { const char *e = getenv("ZOO_TMA_WGRAD"); return !(e && e[0] == '0'); }();
    return on && tma_conv_enabled();
}
// geometry of layer L's weight gradient as seen by tc_conv_wgrad; x = the NHWC tensor the forward conv read
static bool wgrad_geom(const zoo_net *net, const LayerExec &L, const void *x, ConvGeom &g) {
    if (!tma_wgrad_enabled() || net->prec == ZOO_PREC_FP32_SIMT || L.kind != ZOO_LAYER_CONV || !x) return false;
    if (L.first_raw && !net->obs_nhwc) return false;
    g = ConvGeom();
    g.x = x; g.C = L.cin; g.H = L.ih; g.W = L.iw; g.ks = L.ksize; g.s = L.stride; g.p = L.pad; g.OH = L.oh; g.OW = L.ow; g.Cout = L.cout;
    g.first = (L.first_raw && net->obs_nhwc) ? 1 : 0;
    return tc_conv_wgrad_supported(g, net->prec == ZOO_PREC_BF16);
}
static bool implicit_conv_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("ZOO_IMPLICIT_CONV"); on = (e && e[0] == '1') ? 1 : 0; }
    return on == 1;
}

// ---------------------------------------------------------------- kernels -----------------
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f(uint8_t v) { return (float)v; }

// first conv: raw NCHW observation -> cols[M][C*k*k] with k index (c,ky,kx); folds `x ./ 255` (IEEE divide,
// Dopamine_DQN_Atari.jl:28) into the load.  One thread per (m, c, ky) writes k contiguous outputs.
template <typename Tin, typename T>
__global__ void im2col_first_kernel(const Tin *__restrict__ x, T *__restrict__ cols, long long total, int C, int H, int W,
                                    int k, int s, int p, int OH, int OW, int scale255) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int ky = (int)(idx % k);
    long long t = idx / k;
    int c = (int)(t % C);
    long long m = t / C;
    int ow = (int)(m % OW);
    long long t2 = m / OW;
    int oh = (int)(t2 % OH);
    long long b = t2 / OH;
    int iy = oh * s - p + ky;
    T *dst = cols + idx * k;
    const Tin *src = x + ((b * C + c) * H + iy) * (long long)W;
    bool rowok = iy >= 0 && iy < H;
    for (int kx = 0; kx < k; ++kx) {
        int ix = ow * s - p + kx;
        float v = 0.f;
        if (rowok && ix >= 0 && ix < W) {
            v = to_f(src[ix]);
            if (scale255) v = __fdiv_rn(v, 255.f);
        }
        dst[kx] = from_f<T>(v);
    }
}

template <typename T, int VEC> struct alignas(sizeof(T) * VEC) VecT { T v[VEC]; };

// raw observation [B][C][H][W] (u8 | f32) -> NHWC in the net's activation type [B][H][W][C] with `x ./ 255` (IEEE divide) folded in, so the first
// conv runs through the same implicit-GEMM gather as the inner convs.  One thread per pixel, C = 4 channels (16 bytes).
template <typename Tin, typename Tout>
__global__ void obs_to_nhwc4_kernel(const Tin *__restrict__ x, Tout *__restrict__ out, long long pixels, int H, int W, int pad,
                                    int scale255) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    const int HW = H * W;
    const long long b = i / HW, p = i - b * HW;
    const Tin *src = x + b * 4 * HW + p;
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        v[c] = to_f(src[(long long)c * HW]);
        if (scale255) v[c] = __fdiv_rn(v[c], 255.f);
    }
    VecT<Tout, 4> o;
#pragma unroll
    for (int c = 0; c < 4; ++c) o.v[c] = from_f<Tout>(v[c]);
    pdl_trigger();
    // the copy is physically zero-padded ([B][H + 2 pad][W + 2 pad][4], border written once at creation): the first conv then
    // needs no bounds checks, which is what lets its im2col rows be expressed as plain (overlapping) TMA boxes
    const int y = (int)(p / W), xx = (int)(p - (long long)y * W);
    const long long o_idx = (b * (H + 2 * pad) + y + pad) * (W + 2 * pad) + xx + pad;
    *reinterpret_cast<VecT<Tout, 4> *>(out + o_idx * 4) = o;
}

// Same, four consecutive pixels of one image row per thread (W % 4 == 0, 4-pixel-aligned planes): each of the four channel planes is
// read as one 4-pixel vector (a warp covers 128 consecutive pixels of a plane instead of 32 single bytes), and the thread writes
// its 4 x 4 interleaved values as contiguous 8- or 16-byte pieces.  At batch 2048 (Rainbow, 57.8 MB of frames) the one-pixel
// kernel ran at 1.3 TB/s.
template <typename Tin, typename Tout>
__global__ void obs_to_nhwc4_x4_kernel(const Tin *__restrict__ x, Tout *__restrict__ out, long long quads, int H, int W, int pad, int scale255) {
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= quads) return;
    const int HW = H * W;
    const long long pix = i * 4;
    const long long b = pix / HW, p = pix - b * HW;
    const Tin *src = x + b * 4 * HW + p;
    VecT<Tin, 4> in[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) in[c] = *reinterpret_cast<const VecT<Tin, 4> *>(src + (long long)c * HW);
    pdl_trigger();
    const int y = (int)(p / W), xx = (int)(p - (long long)y * W);
    Tout *dst = out + ((b * (H + 2 * pad) + y + pad) * (W + 2 * pad) + xx + pad) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        VecT<Tout, 4> o;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float v = to_f(in[c].v[j]);
            if (scale255) v = __fdiv_rn(v, 255.f);
            o.v[c] = from_f<Tout>(v);
        }
        *reinterpret_cast<VecT<Tout, 4> *>(dst + j * 4) = o;
    }
}

// conv reading internal NHWC activations: cols[M][k*k*C] with k index (ky,kx,c); VEC channels per thread
template <typename T, int VEC>
__global__ void im2col_nhwc_kernel(const T *__restrict__ x, T *__restrict__ cols, long long total, int C, int H, int W, int k,
                                   int s, int p, int OH, int OW) {
    pdl_wait();
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = C / VEC;
    int cc = (int)(idx % cv);
    long long t = idx / cv;
    int kx = (int)(t % k); t /= k;
    int ky = (int)(t % k);
    long long m = t / k;
    int ow = (int)(m % OW);
    long long t2 = m / OW;
    int oh = (int)(t2 % OH);
    long long b = t2 / OH;
    int iy = oh * s - p + ky, ix = ow * s - p + kx;
    VecT<T, VEC> v;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        v = *reinterpret_cast<const VecT<T, VEC> *>(x + (((b * H + iy) * W + ix) * (long long)C + cc * VEC));
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) v.v[i] = from_f<T>(0.f);
    }
    pdl_trigger();
    *reinterpret_cast<VecT<T, VEC> *>(cols + idx * VEC) = v;
}

// gather-form col2im (no atomics): dx[b,h,w,c] = sum over (ky,kx) of dcols[(b,oh,ow)][(ky,kx,c)], then relu' mask
template <typename T, int VEC>
__global__ void col2im_nhwc_kernel(const T *__restrict__ dcols, T *__restrict__ dx, const T *__restrict__ act, long long total,
                                   int C, int H, int W, int k, int s, int p, int OH, int OW) {
    pdl_wait();
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = C / VEC;
    int cc = (int)(idx % cv);
    long long t = idx / cv;
    int w = (int)(t % W); t /= W;
    int h = (int)(t % H);
    long long b = t / H;
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    const long long KK = (long long)k * k * C;
    for (int ky = 0; ky < k; ++ky) {
        int ty = h + p - ky;
        if (ty < 0 || ty % s) continue;
        int oh = ty / s;
        if (oh >= OH) continue;
        for (int kx = 0; kx < k; ++kx) {
            int tx = w + p - kx;
            if (tx < 0 || tx % s) continue;
            int ow = tx / s;
            if (ow >= OW) continue;
            long long m = (b * OH + oh) * OW + ow;
            VecT<T, VEC> v = *reinterpret_cast<const VecT<T, VEC> *>(dcols + m * KK + (long long)(ky * k + kx) * C + cc * VEC);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] += to_f(v.v[i]);
        }
    }
    VecT<T, VEC> o;
    if (act) {
        VecT<T, VEC> a = *reinterpret_cast<const VecT<T, VEC> *>(act + idx * VEC);
#pragma unroll
        for (int i = 0; i < VEC; ++i) o.v[i] = from_f<T>(to_f(a.v[i]) > 0.f ? acc[i] : 0.f);
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) o.v[i] = from_f<T>(acc[i]);
    }
    pdl_trigger();
    *reinterpret_cast<VecT<T, VEC> *>(dx + idx * VEC) = o;
}

// bias gradient: db[n] += sum_m dy[m][n]
template <typename T>
__global__ void colsum_kernel(const T *__restrict__ dy, long long M, int N, float *__restrict__ db) {
    pdl_wait();
    pdl_trigger();
    __shared__ float red[8][33];
    int n = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (n < N)
        for (long long m = (long long)blockIdx.y * 8 + threadIdx.y; m < M; m += (long long)gridDim.y * 8) acc += to_f(dy[m * N + n]);
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
        atomicAdd(db + n, s);
    }
}

// embed(tau, N_em) = cos.(Float32(pi) .* (1:N_em) .* tau)  -- iqn.jl:157
template <typename T>
__global__ void embed_kernel(const float *__restrict__ tau, T *__restrict__ emb, long long rows, int nem) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * nem) return;
    int i = (int)(idx % nem);
    long long r = idx / nem;
    float a = __fmul_rn(__fmul_rn(3.14159274101257324f, (float)(i + 1)), tau[r]);
    emb[idx] = from_f<T>(cosf(a));
}

// merged[r][f] = feat[r / N][f] * ea[r][f]  -- iqn.jl:28-30
template <typename T>
__global__ void iqn_merge_kernel(const T *__restrict__ feat, const T *__restrict__ ea, T *__restrict__ merged, long long rows,
                                 int F, int N) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * F) return;
    int f = (int)(idx % F);
    long long r = idx / F;
    merged[idx] = from_f<T>(to_f(feat[(r / N) * F + f]) * to_f(ea[idx]));
}

template <typename T, int VEC>
__global__ void iqn_merge_vec_kernel(const T *__restrict__ feat, const T *__restrict__ ea, T *__restrict__ merged, long long rows, int F, int N) {
    const int fv = F / VEC;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * fv) return;
    const int f = (int)(idx % fv);
    const long long r = idx / fv;
    const VecT<T, VEC> a = *reinterpret_cast<const VecT<T, VEC> *>(feat + (r / N) * F + (long long)f * VEC);
    const VecT<T, VEC> e = *reinterpret_cast<const VecT<T, VEC> *>(ea + idx * VEC);
    VecT<T, VEC> o;
#pragma unroll
    for (int i = 0; i < VEC; ++i) o.v[i] = from_f<T>(to_f(a.v[i]) * to_f(e.v[i]));
    *reinterpret_cast<VecT<T, VEC> *>(merged + idx * VEC) = o;
}

template <typename T, int VEC>
__global__ void iqn_merge_bwd_vec_kernel(const T *__restrict__ dm, const T *__restrict__ feat, const T *__restrict__ ea, T *__restrict__ dea,
                                         T *__restrict__ dfeat, int B, int F, int N, int phi_relu, int psi_relu) {
    const int fv = F / VEC;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * fv) return;
    const int f = (int)(idx % fv);
    const long long b = idx / fv;
    const VecT<T, VEC> ftv = *reinterpret_cast<const VecT<T, VEC> *>(feat + b * F + (long long)f * VEC);
    float ft[VEC], acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { ft[i] = to_f(ftv.v[i]); acc[i] = 0.f; }
    for (int n = 0; n < N; ++n) {
        const long long o = (b * N + n) * F + (long long)f * VEC;
        const VecT<T, VEC> d = *reinterpret_cast<const VecT<T, VEC> *>(dm + o);
        const VecT<T, VEC> e = *reinterpret_cast<const VecT<T, VEC> *>(ea + o);
        VecT<T, VEC> out;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float dd = to_f(d.v[i]), ee = to_f(e.v[i]);
            out.v[i] = from_f<T>((!phi_relu || ee > 0.f) ? dd * ft[i] : 0.f);
            acc[i] += dd * ee;
        }
        *reinterpret_cast<VecT<T, VEC> *>(dea + o) = out;
    }
    VecT<T, VEC> of;
#pragma unroll
    for (int i = 0; i < VEC; ++i) of.v[i] = from_f<T>((!psi_relu || ft[i] > 0.f) ? acc[i] : 0.f);
    *reinterpret_cast<VecT<T, VEC> *>(dfeat + b * F + (long long)f * VEC) = of;
}

// Adjoint of the FUSED merge (phi's output was never stored; merged = psi .* relu(phi), both factors >= 0):
//   relu'(phi) = [merged > 0] wherever psi > 0, and where psi == 0 the gradient into phi is zero anyway;
//   relu(phi)  = merged / psi wherever psi > 0, and where psi == 0 relu'(psi) = 0 kills d psi anyway.
// so  dphi = dm .* psi .* [merged > 0],   dpsi[b] = [psi > 0] / psi .* sum_n dm .* merged,   dbias_phi += sum over rows of dphi
// (the bias gradient of phi's last layer is folded in: it was a separate column-sum pass over the 0.5 GB dphi tensor).
// CTA = 32 feature vectors (x) x 8 samples (y): one thread per (feature vector, sample) walks the N quantile rows; the bias sums
// of the 8 samples are combined in shared memory, so the CTA issues one atomic per feature.
template <typename T, int VEC, int G, int MINB>
__global__ void __launch_bounds__(256, MINB) iqn_merge_bwd_fused_kernel(const T *__restrict__ dm, const T *__restrict__ feat, const T *__restrict__ merged,
                                                                  T *__restrict__ dea, T *__restrict__ dfeat, float *__restrict__ dbias, int B, int F, int N) {
    __shared__ float red[8][32 * VEC + 1];
    const int fv = F / VEC;
    const int f = blockIdx.x * 32 + threadIdx.x;
    const int b = blockIdx.y * 8 + threadIdx.y;
    float bsum[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) bsum[i] = 0.f;
    if (f < fv && b < B) {
        const VecT<T, VEC> ftv = *reinterpret_cast<const VecT<T, VEC> *>(feat + (long long)b * F + (long long)f * VEC);
        float ft[VEC], acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) { ft[i] = to_f(ftv.v[i]); acc[i] = 0.f; }
        const long long o0 = (long long)b * N * F + (long long)f * VEC;
        auto one = [&](const VecT<T, VEC> &d, const VecT<T, VEC> &m, long long o) {
            VecT<T, VEC> out;
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float dd = to_f(d.v[i]), mm = to_f(m.v[i]);
                const float de = mm > 0.f ? dd * ft[i] : 0.f;
                out.v[i] = from_f<T>(de);
                bsum[i] += to_f(out.v[i]);  // the bias gradient sums what the weight gradient will read
                acc[i] = fmaf(dd, mm, acc[i]);
            }
            *reinterpret_cast<VecT<T, VEC> *>(dea + o) = out;
        };
        if (G > 0 && (N % (2 * (G > 0 ? G : 1))) == 0) {
            // software pipeline over groups of 4 quantile rows: the 8 loads of group g + 1 are issued before group g is reduced and
            // stored, so the thread always has a group in flight (the plain unrolled loop drained its loads before every compute phase:
            // 80 % of its stall samples were long-scoreboard, 4.67 TB/s)
            constexpr int GG = G > 0 ? G : 1;
            VecT<T, VEC> d0[GG], m0[GG], d1[GG], m1[GG];
            auto ldg = [&](VecT<T, VEC> (&d)[GG], VecT<T, VEC> (&m)[GG], int n) {
#pragma unroll
                for (int u = 0; u < GG; ++u) {
                    d[u] = *reinterpret_cast<const VecT<T, VEC> *>(dm + o0 + (long long)(n + u) * F);
                    m[u] = *reinterpret_cast<const VecT<T, VEC> *>(merged + o0 + (long long)(n + u) * F);
                }
            };
            ldg(d0, m0, 0);
#pragma unroll 1
            for (int n = 0; n < N; n += 2 * GG) {
                ldg(d1, m1, n + GG);
#pragma unroll
                for (int u = 0; u < GG; ++u) one(d0[u], m0[u], o0 + (long long)(n + u) * F);
                if (n + 2 * GG < N) ldg(d0, m0, n + 2 * GG);
#pragma unroll
                for (int u = 0; u < GG; ++u) one(d1[u], m1[u], o0 + (long long)(n + GG + u) * F);
            }
        } else {
#pragma unroll 4
            for (int n = 0; n < N; ++n) {
                const long long o = o0 + (long long)n * F;
                one(*reinterpret_cast<const VecT<T, VEC> *>(dm + o), *reinterpret_cast<const VecT<T, VEC> *>(merged + o), o);
            }
        }
        VecT<T, VEC> of;
#pragma unroll
        for (int i = 0; i < VEC; ++i) of.v[i] = from_f<T>(ft[i] > 0.f ? acc[i] / ft[i] : 0.f);
        *reinterpret_cast<VecT<T, VEC> *>(dfeat + (long long)b * F + (long long)f * VEC) = of;
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) red[threadIdx.y][threadIdx.x * VEC + i] = bsum[i];
    __syncthreads();
    for (int j = threadIdx.y * 32 + threadIdx.x; j < 32 * VEC; j += 256) {
        const int fcol = blockIdx.x * 32 * VEC + j;
        if (fcol < F) {
            float s = 0.f;
#pragma unroll
            for (int y = 0; y < 8; ++y) s += red[y][j];
            atomicAdd(dbias + fcol, s);
        }
    }
}

// adjoint of the merge: dea = dm * feat * relu'(ea);  dfeat[b] = relu'(feat) * sum_n dm * ea
template <typename T>
__global__ void iqn_merge_bwd_kernel(const T *__restrict__ dm, const T *__restrict__ feat, const T *__restrict__ ea,
                                     T *__restrict__ dea, T *__restrict__ dfeat, int B, int F, int N, int phi_relu, int psi_relu) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * F) return;
    int f = (int)(idx % F);
    long long b = idx / F;
    float ft = to_f(feat[idx]);
    float acc = 0.f;
    for (int n = 0; n < N; ++n) {
        long long o = (b * N + n) * F + f;
        float d = to_f(dm[o]), e = to_f(ea[o]);
        dea[o] = from_f<T>((!phi_relu || e > 0.f) ? d * ft : 0.f);
        acc += d * e;
    }
    dfeat[idx] = from_f<T>((!psi_relu || ft > 0.f) ? acc : 0.f);
}

// bf16 nets whose output layer is wide (Rainbow: Dense(512 => n_actions * n_atoms) = 306 columns): the head kernels leave d(loss)/d(logits)
// in fp32 [rows][N] -- an odd row pitch that TMA cannot address and an operand type the bf16 engine does not take, so both gradient
// GEMMs of that layer used to fall back to the CUDA-core engine (67 us of a 0.68 ms update at B = 1024).  One pass writes the bf16 copy
// with rows padded to a multiple of 8 columns (zeros), which both GEMMs read by TMA.
__global__ void head_dy_to_bf16_kernel(const float *__restrict__ dy, bf16 *__restrict__ out, long long rows, int N, int Np) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one 8-column group
    const int gpr = Np >> 3;
    if (i >= rows * gpr) return;
    const long long m = i / gpr;
    const int c0 = (int)(i - m * gpr) << 3;
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + 2 * k;
        const float a = c < N ? dy[m * N + c] : 0.f, b = c + 1 < N ? dy[m * N + c + 1] : 0.f;
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        w[k] = *reinterpret_cast<const uint32_t *>(&h);
    }
    *reinterpret_cast<uint4 *>(out + m * Np + c0) = make_uint4(w[0], w[1], w[2], w[3]);
}

// DuelingNetwork: q = val .+ adv .- mean(adv, dims=1)  -- common.jl:51-56
__global__ void dueling_combine_kernel(const float *__restrict__ val, const float *__restrict__ adv, float *__restrict__ q, int B, int A) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float s = 0.f;
    for (int a = 0; a < A; ++a) s += adv[b * A + a];
    float mean = s / (float)A;
    for (int a = 0; a < A; ++a) q[b * A + a] = val[b] + adv[b * A + a] - mean;
}
__global__ void dueling_combine_bwd_kernel(const float *__restrict__ dq, float *__restrict__ dval, float *__restrict__ dadv, int B, int A) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float s = 0.f;
    for (int a = 0; a < A; ++a) s += dq[b * A + a];
    dval[b] = s;
    float mean = s / (float)A;
    for (int a = 0; a < A; ++a) dadv[b * A + a] = dq[b * A + a] - mean;
}
template <typename T>
__global__ void add_mask_kernel(const T *__restrict__ a, const T *__restrict__ b, const T *__restrict__ act, T *__restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = to_f(a[i]) + to_f(b[i]);
    if (act && !(to_f(act[i]) > 0.f)) v = 0.f;
    out[i] = from_f<T>(v);
}

__global__ void cast_bf16_kernel(const float *__restrict__ in, bf16 *__restrict__ out, long long n) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 v = *reinterpret_cast<const float4 *>(in + i);
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<__nv_bfloat162 *>(out + i) = lo;
        *reinterpret_cast<__nv_bfloat162 *>(out + i + 2) = hi;
    } else {
        for (; i < n; ++i) out[i] = __float2bfloat16_rn(in[i]);
    }
}

// ABI (Julia bytes) <-> internal layout.  dir 0: abi -> internal, 1: internal -> abi.
__global__ void permute_param_kernel(float *__restrict__ internal, float *__restrict__ abi, TensorInfo ti, int dir) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // internal linear index
    if (i >= ti.count) return;
    long long src;
    const int hw = ti.fh * ti.fw;
    if (ti.kind == 2) {  // bias
        long long o = i;
        if (ti.perm_out) { int c = (int)(o % ti.fc); long long p = o / ti.fc; o = (long long)c * hw + p; }
        src = o;
    } else if (ti.kind == 0) {  // dense: internal [out][in] <- abi [in][out]
        long long o = i / ti.n_in, in = i % ti.n_in;
        if (ti.perm_out) { int c = (int)(o % ti.fc); long long p = o / ti.fc; o = (long long)c * hw + p; }
        if (ti.perm_in) { int c = (int)(in % ti.fc); long long p = in / ti.fc; in = (long long)c * hw + p; }
        src = in * ti.n_out + o;
    } else {  // conv
        if (ti.conv_nhwc) {  // internal [O][ky][kx][C] <- abi [O][C][ky][kx]
            int c = (int)(i % ti.n_in);
            long long t = i / ti.n_in;
            int kx = (int)(t % ti.ksize); t /= ti.ksize;
            int ky = (int)(t % ti.ksize);
            long long o = t / ti.ksize;
            src = ((o * ti.n_in + c) * ti.ksize + ky) * ti.ksize + kx;
        } else src = i;
    }
    if (dir == 0) internal[i] = abi[src];
    else abi[src] = internal[i];
}

// ---------------------------------------------------------------- helpers -----------------
static inline bool is_bf(const zoo_net *n) { return n->prec == ZOO_PREC_BF16; }

template <typename T>
static zoo_status launch_im2col_nhwc(const void *x, void *cols, long long M, const LayerExec &L, cudaStream_t st) {
    constexpr int V = 16 / sizeof(T);
    if (L.cin % V != 0 && L.cin % (V / 2) == 0) {  // e.g. the 4-channel bf16 frame stack: 8-byte chunks
        long long total = M * L.ksize * L.ksize * (L.cin / (V / 2));
        ZOO_CUDA(launch_pdl(im2col_nhwc_kernel<T, V / 2>, dim3(cdiv(total, 256)), dim3(256), (size_t)0, st, (const T *)x, (T *)cols, total, L.cin, L.ih,
                            L.iw, L.ksize, L.stride, L.pad, L.oh, L.ow));
        ZOO_LAUNCH_CHECK();
        return ZOO_OK;
    }
    if (L.cin % V == 0) {
        long long total = M * L.ksize * L.ksize * (L.cin / V);
        ZOO_CUDA(launch_pdl(im2col_nhwc_kernel<T, V>, dim3(cdiv(total, 256)), dim3(256), (size_t)0, st, (const T *)x, (T *)cols, total, L.cin, L.ih, L.iw,
                            L.ksize, L.stride, L.pad, L.oh, L.ow));
    } else {
        long long total = M * L.ksize * L.ksize * L.cin;
        ZOO_CUDA(launch_pdl(im2col_nhwc_kernel<T, 1>, dim3(cdiv(total, 256)), dim3(256), (size_t)0, st, (const T *)x, (T *)cols, total, L.cin, L.ih, L.iw,
                            L.ksize, L.stride, L.pad, L.oh, L.ow));
    }
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

template <typename T>
static zoo_status launch_col2im_nhwc(const void *dcols, void *dx, const void *act, long long rows, const LayerExec &L, cudaStream_t st) {
    constexpr int V = 16 / sizeof(T);
    if (L.cin % V == 0) {
        long long total = rows * L.ih * L.iw * (L.cin / V);
        ZOO_CUDA(launch_pdl(col2im_nhwc_kernel<T, V>, dim3(cdiv(total, 256)), dim3(256), (size_t)0, st, (const T *)dcols, (T *)dx, (const T *)act, total,
                            L.cin, L.ih, L.iw, L.ksize, L.stride, L.pad, L.oh, L.ow));
    } else {
        long long total = rows * L.ih * L.iw * L.cin;
        ZOO_CUDA(launch_pdl(col2im_nhwc_kernel<T, 1>, dim3(cdiv(total, 256)), dim3(256), (size_t)0, st, (const T *)dcols, (T *)dx, (const T *)act, total,
                            L.cin, L.ih, L.iw, L.ksize, L.stride, L.pad, L.oh, L.ow));
    }
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

static zoo_status launch_colsum(const void *dy, int dy_bf16, long long M, int N, float *db, cudaStream_t st) {
    const int xb = cdiv(N, 32);
    dim3 grid(xb, (unsigned)std::max<long long>(1, std::min<long long>(std::max(64, 2368 / xb), M / 64)));
    dim3 block(32, 8);
    if (dy_bf16) ZOO_CUDA(launch_pdl(colsum_kernel<bf16>, grid, block, (size_t)0, st, (const bf16 *)dy, M, N, db));
    else ZOO_CUDA(launch_pdl(colsum_kernel<float>, grid, block, (size_t)0, st, (const float *)dy, M, N, db));
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

// weight as the B operand; bf16 shadow only when the tcgen05 engine will take the GEMM, fp32 master otherwise
static Operand weight_operand(const zoo_net *net, const TensorInfo &W, const Operand &A, int M, int N, int K, long long ld, int kmajor) {
    Operand B{net->params + W.offset, 0, ld, kmajor};
    if (is_bf(net) && A.is_bf16) {
        Operand Bs{net->shadow + W.offset, 1, ld, kmajor};
        if (tc_gemm_supported(A, Bs, M, N, K)) return Bs;
    }
    return B;
}

// Every tensor at offset >= late_off now has its gradient enqueued on the side lane, and this layer's dgrad -- the last
// reader of those parameters -- is enqueued on `st`: once both are done the late bucket (all-reduce + optimiser) may run.
static zoo_status net_late_bucket(zoo_net *net, cudaStream_t st, cudaStream_t sw, bool par) {
    cudaStream_t sc = sw;
    if (par) {  // third lane, so the side lane carries on with the remaining bias / weight gradients meanwhile
        sc = net->comm;
        ZOO_CUDA(cudaEventRecord(net->ev_late, st));
        ZOO_CUDA(cudaStreamWaitEvent(sc, net->ev_late, 0));
        ZOO_CUDA(cudaEventRecord(net->ev_late2, sw));
        ZOO_CUDA(cudaStreamWaitEvent(sc, net->ev_late2, 0));
        net->comm_used = true;
    }
    zoo_status s = net->late_cb(net->late_ctx, sc);
    net->late_cb = nullptr;
    return s;
}

// in_kind: 0 raw u8, 1 raw f32, 2 internal T
// cols_rows: leading sample rows whose im2col matrices the later wgrad needs (0 = forward only, e.g. the target net)
static zoo_status chain_forward(zoo_net *net, Chain &ch, const void *in, int in_kind, long long rows, cudaStream_t st, long long cols_rows = 0) {
    ZOO_REQUIRE(rows <= ch.rows_cap, "forward: %lld rows exceed the capacity %lld this net was created with", rows, ch.rows_cap);
    const int bf = is_bf(net);
    const void *cur = in;
    int cur_kind = in_kind;
    for (auto &L : ch.layers) {
        if (net->skip_head && &ch == &net->a && &L == &ch.layers.back()) break;  // the fused tail evaluates the head layer
        const long long M = rows * L.opr;
        const TensorInfo &W = net->tensors[L.wt];
        const TensorInfo &Bt = net->tensors[L.bt];
        prof_label("fwd:t%d[%lldx%dx%lld]", L.wt, M, L.cout, L.K);
        Operand A;
        Epilogue e;
        e.out = L.out; e.out_bf16 = L.out_fp32 ? 0 : bf; e.sm = L.cout; e.sn = 1;
        e.bias = net->params + Bt.offset; e.bias_mode = 1; e.relu = L.relu; e.tf32_passes = net->tf32_passes;
        if (L.rowmul) { e.rowmul = L.rowmul; e.rowmul_div = L.rowmul_div; e.rowmul_ld = L.rowmul_ld; e.rowmul_nonneg = L.rowmul_nonneg; }
        if (L.kind == ZOO_LAYER_CONV && L.first_raw && net->obs_nhwc) {
            if (cur_kind == 2) { set_err("conv: first layer expects a raw observation"); return ZOO_BAD_ARG; }
            const int oh_ = net->desc.in_h, ow_ = net->desc.in_w, op_ = (L.ih - oh_) / 2;  // L.ih / L.iw describe the padded copy
            const long long pixels = rows * oh_ * ow_;
            if (net->prestaged_src == cur && rows <= net->prestaged_rows) {  // the staging launch already wrote the padded NHWC copy
                cur = net->obs_nhwc_buf;
                cur_kind = 2;
            } else {
            prof_label("obs2nhwc[%lld]", pixels);
            const bool x4 = (ow_ % 4) == 0 && ((uintptr_t)cur % (cur_kind == 0 ? 4 : 16)) == 0;
#define OBS2NHWC(Tin, Tout)                                                                                                                        \
    do {                                                                                                                                           \
        if (x4) obs_to_nhwc4_x4_kernel<Tin, Tout><<<cdiv(pixels / 4, 256), 256, 0, st>>>((const Tin *)cur, (Tout *)net->obs_nhwc_buf, pixels / 4, oh_, ow_, op_, L.scale255); \
        else obs_to_nhwc4_kernel<Tin, Tout><<<cdiv(pixels, 256), 256, 0, st>>>((const Tin *)cur, (Tout *)net->obs_nhwc_buf, pixels, oh_, ow_, op_, L.scale255);              \
    } while (0)
            if (cur_kind == 0) { if (bf) OBS2NHWC(uint8_t, bf16); else OBS2NHWC(uint8_t, float); }
            else { if (bf) OBS2NHWC(float, bf16); else OBS2NHWC(float, float); }
#undef OBS2NHWC
            ZOO_LAUNCH_CHECK();
            cur = net->obs_nhwc_buf;
            cur_kind = 2;
            }
        }
        const bool raw_conv = L.kind == ZOO_LAYER_CONV && L.first_raw && !net->obs_nhwc;
        if (L.kind == ZOO_LAYER_CONV) {
            // implicit GEMM on tcgen05 (the producer warps gather the im2col rows): no cols matrix on the critical path.
            // The explicit cols that wgrad consumes are built for the first cols_rows samples on the side lane.
            ConvGeom g;
            g.x = cur; g.C = L.cin; g.H = L.ih; g.W = L.iw; g.ks = L.ksize; g.s = L.stride; g.p = L.pad; g.OH = L.oh; g.OW = L.ow;
            g.Cout = L.cout; g.scale255 = L.scale255;
            g.mode = raw_conv ? (cur_kind == 0 ? CONV_RAW_U8 : CONV_RAW_F32) : CONV_NHWC;
            if (raw_conv && cur_kind == 2) { set_err("conv: first layer expects a raw observation"); return ZOO_BAD_ARG; }
            // (a raw first conv -- bf16 nets -- keeps the explicit im2col + TMA GEMM path: its register gather is slower)
            bool implicit = implicit_conv_enabled() && net->prec != ZOO_PREC_FP32_SIMT && !raw_conv && L.cout >= 8 && tc_conv_supported(g, bf);
            if (!raw_conv && net->prec != ZOO_PREC_FP32_SIMT && L.cout >= 8 && tma_conv_enabled()) {
                // preferred: im2col rows fetched by TMA (4-D boxes) -- the producer stays one asynchronous thread
                ConvGeom gt = g;
                gt.mode = CONV_TMA;
                gt.first = (L.first_raw && net->obs_nhwc) ? 1 : 0;
                if (tc_conv_supported(gt, bf)) { g = gt; implicit = true; }
            }
            ConvGeom gw;
            const bool wg_implicit = implicit && wgrad_geom(net, L, cur, gw);
            const long long Mc = implicit ? (wg_implicit ? 0 : cols_rows * L.opr) : M;  // rows of the explicit cols matrix to build
            cudaStream_t sc = st;
            const bool par = implicit && net->side && !g_prof_on;
            if (par && Mc > 0) { ZOO_TRY(net_side_fork(net, st)); sc = net->side; }
            if (Mc > 0) {
                prof_label("im2col:t%d[%lldx%lld]", L.wt, Mc, L.K);
                if (raw_conv) {
                    long long total = Mc * L.cin * L.ksize;
                    int blocks = cdiv(total, 256);
#define IM2COL_FIRST(Tin, T)                                                                                              \
    im2col_first_kernel<Tin, T><<<blocks, 256, 0, sc>>>((const Tin *)cur, (T *)L.cols, total, L.cin, L.ih, L.iw, L.ksize, \
                                                        L.stride, L.pad, L.oh, L.ow, L.scale255)
                    if (cur_kind == 0) { if (bf) IM2COL_FIRST(uint8_t, bf16); else IM2COL_FIRST(uint8_t, float); }
                    else { if (bf) IM2COL_FIRST(float, bf16); else IM2COL_FIRST(float, float); }
#undef IM2COL_FIRST
                    ZOO_LAUNCH_CHECK();
                } else {
                    if (bf) ZOO_TRY(launch_im2col_nhwc<bf16>(cur, L.cols, Mc, L, sc));
                    else ZOO_TRY(launch_im2col_nhwc<float>(cur, L.cols, Mc, L, sc));
                }
            }
            if (implicit) {
                prof_label("fwd:t%d[%lldx%dx%lld]", L.wt, M, L.cout, L.K);
                Operand Ad{nullptr, bf, L.K, 1};
                Operand Wk = weight_operand(net, W, Ad, (int)M, L.cout, (int)L.K, L.K, 1);
                if (Wk.is_bf16 != bf) { set_err("conv: weight shadow missing"); return ZOO_CUDA_ERROR; }
                ZOO_TRY(tc_conv_gemm(g, Wk, (int)M, L.cout, (int)L.K, e, net->ws, net->ws_elems, st));
                cur = L.out;
                cur_kind = 2;
                continue;
            }
            A = Operand{L.cols, bf, L.K, 1};
        } else {
            ZOO_REQUIRE(cur_kind != 0, "dense: uint8 observations need a conv or scale layer first");
            const int xbf = cur_kind == 2 ? bf : 0;
            if (L.out_fp32 && !L.relu && narrow_enabled() && narrow_supported(L.cout, L.K, xbf)) {
                ZOO_TRY(narrow_forward(cur, xbf, net->params + W.offset, net->params + Bt.offset, (float *)L.out, M, L.cout, L.K, st));
                cur = L.out;
                cur_kind = 2;
                continue;
            }
            A = Operand{cur, xbf, L.K, 1};
        }
        Operand B = weight_operand(net, W, A, (int)M, L.cout, (int)L.K, L.K, 1);
        ZOO_TRY(gemm(A, B, (int)M, L.cout, (int)L.K, e, 0, net->ws, net->ws_elems, st));
        cur = L.out;
        cur_kind = 2;
    }
    return ZOO_OK;
}

static bool head_bf16_enabled() {  // ZOO_HEAD_BF16=0: the wide output layer's gradients on the CUDA-core engine again (A/B, parity tests)
    static const bool on = [] { const char *e = getenv("ZOO_HEAD_BF16"); return !(e && e[0] == '0'); }();
    return on;
}

// Each layer's dout must already hold d(loss)/d(out) with the layer's own relu' applied.
// dx (optional): gradient w.r.t. the chain input, internal T [rows][in_feat]; dx_mask: relu' source for it.
static zoo_status chain_backward(zoo_net *net, Chain &ch, const void *in, int in_kind, long long rows, void *dx, const void *dx_mask,
                                 cudaStream_t st, int bias_done = -1) {  // bias_done: layer whose bias gradient a fused kernel already produced
    const int bf = is_bf(net);
    int top = (int)ch.layers.size() - 1;
    if (net->skip_head && &ch == &net->a) { --top; bias_done = top; }  // head layer + the bias gradient below it: done by the fused tail
    for (int li = top; li >= 0; --li) {
        LayerExec &L = ch.layers[li];
        const long long M = rows * L.opr;
        const TensorInfo &W = net->tensors[L.wt];
        const TensorInfo &Bt = net->tensors[L.bt];
        int dy_bf = L.out_fp32 ? 0 : bf;
        const void *dy_p = L.dout;   // what the two gradient GEMMs read (the bias gradient always sums L.dout itself)
        long long dy_ld = L.cout;
        if (li > 0 && L.kind == ZOO_LAYER_DENSE && L.out_fp32 && !L.relu && narrow_enabled() && narrow_supported(L.cout, L.K, bf) &&
            !(net->late_cb && W.offset == net->late_off)) {
            // the Q head: data, weight and bias gradient plus the previous layer's bias gradient in one pass over its input
            LayerExec &P = ch.layers[li - 1];
            prof_label("dgrad+wgrad:t%d[%lldx%lldx%d]", L.wt, M, L.K, L.cout);
            ZOO_TRY(narrow_backward(P.out, bf, (const float *)L.dout, net->params + W.offset, P.dout, net->grads + W.offset, net->grads + Bt.offset,
                                    net->grads + net->tensors[P.bt].offset, M, L.cout, L.K, P.relu, st));
            bias_done = li - 1;
            continue;
        }
        if (bf && L.out_fp32 && L.kind == ZOO_LAYER_DENSE && L.dcols && head_bf16_enabled()) {
            const int Np = (L.cout + 7) & ~7;
            prof_label("dy2bf16:t%d[%lldx%d]", L.wt, M, L.cout);
            head_dy_to_bf16_kernel<<<cdiv(M * (Np >> 3), 256), 256, 0, st>>>((const float *)L.dout, (bf16 *)L.dcols, M, L.cout, Np);
            ZOO_LAUNCH_CHECK();
            dy_p = L.dcols; dy_ld = Np; dy_bf = 1;
        }
        // bias grad + wgrad only need this layer's dout, which is complete on `st` here: they go to the side lane and
        // overlap the dgrad chain (the critical path) that continues on `st`
        // (the bottom layer has no data gradient: its bias / weight gradient ARE the rest of the critical path and stay on `st`
        // instead of queueing behind the previous layer's weight gradient on the side lane)
        static const bool last_on_main = [] { const char *e = getenv("ZOO_LAST_WGRAD_MAIN"); return !(e && e[0] == '0'); }();
        const bool par = net->side && !g_prof_on && !(last_on_main && li == 0 && !dx && &ch == &net->a && ch.layers.size() > 1);
        cudaStream_t sw = par ? net->side : st;
        float *wsw = par ? net->ws_side : net->ws;
        const size_t wsw_elems = par ? net->ws_side_elems : net->ws_elems;
        if (par) ZOO_TRY(net_side_fork(net, st));
        // (bottom layer on the main lane: its small bias gradient goes to the side lane instead, so that the weight gradient -- the
        // end of the critical path -- starts 5 us earlier)
        const bool bias_aside = !par && net->side && !g_prof_on && last_on_main && li == 0 && !dx && &ch == &net->a && ch.layers.size() > 1;
        if (bias_aside) ZOO_TRY(net_side_fork(net, st));
        if (bias_done != li) {
            prof_label("bgrad:t%d", L.bt);
            ZOO_TRY(launch_colsum(L.dout, L.out_fp32 ? 0 : bf, M, L.cout, net->grads + Bt.offset, bias_aside ? net->side : sw));
        }
        // wgrad: dW[cout][K] += dY^T X   (both operands MN-major: stored [rows][features])
        ConvGeom gw;
        const void *xin = L.kind != ZOO_LAYER_CONV ? nullptr : (li == 0 ? ((L.first_raw && net->obs_nhwc) ? net->obs_nhwc_buf : nullptr) : ch.layers[li - 1].out);
        prof_label("wgrad:t%d[%dx%lldx%lld]", L.wt, L.cout, L.K, M);
        if (dy_bf == bf && wgrad_geom(net, L, xin, gw)) {
            // convolution: dY and the activations arrive as TMA boxes, the im2col matrix is never built (tc_conv_wgrad)
            Epilogue ew;
            ew.out = net->grads + W.offset; ew.tf32_passes = net->tf32_passes;
            ZOO_TRY(tc_conv_wgrad(gw, L.dout, bf, (int)rows, ew, sw));
        } else {
            Operand X;
            if (L.kind == ZOO_LAYER_CONV) X = Operand{L.cols, bf, L.K, 0};
            else if (li == 0) X = Operand{in, in_kind == 2 ? bf : 0, L.K, 0};
            else X = Operand{ch.layers[li - 1].out, bf, L.K, 0};
            Operand dYt{dy_p, dy_bf, dy_ld, 0};
            Epilogue ew;
            ew.out = net->grads + W.offset; ew.sm = L.K; ew.sn = 1; ew.tf32_passes = net->tf32_passes;
            ew.accumulate = 2;  // grads are zeroed per update: plain store, or atomic add when the GEMM is K-split
            ZOO_TRY(gemm(dYt, X, L.cout, (int)L.K, (int)M, ew, 0, wsw, wsw_elems, sw));
        }
        const bool late_here = net->late_cb && W.offset == net->late_off;
        // dgrad
        if (li == 0 && !dx) { if (late_here) ZOO_TRY(net_late_bucket(net, st, sw, par)); continue; }
        Operand dY{dy_p, dy_bf, dy_ld, 1};
        prof_label("dgrad:t%d[%lldx%lldx%d]", L.wt, M, L.K, L.cout);
        Operand Wn = weight_operand(net, W, dY, (int)M, (int)L.K, L.cout, L.K, 0);
        Epilogue ed;
        ed.sm = L.K; ed.sn = 1; ed.out_bf16 = bf; ed.tf32_passes = net->tf32_passes;
        if (L.kind == ZOO_LAYER_CONV) {
            ZOO_REQUIRE(li > 0, "backward: gradient w.r.t. the observation is not supported");
            LayerExec &P = ch.layers[li - 1];
            ConvGeom g;
            g.x = L.dout; g.mode = CONV_DGRAD; g.C = L.cin; g.H = L.ih; g.W = L.iw; g.ks = L.ksize; g.s = L.stride; g.p = L.pad;
            g.OH = L.oh; g.OW = L.ow; g.Cout = L.cout;
            if (tma_conv_enabled() && tma_dgrad_enabled() && net->prec != ZOO_PREC_FP32_SIMT && Wn.is_bf16 == bf) {
                ConvGeom gt = g;
                gt.mode = CONV_TMA_DGRAD;
                if (tc_conv_supported(gt, bf)) g = gt;  // dY boxes fetched by TMA per residue class, no dcols, no col2im
            }
            if ((g.mode == CONV_TMA_DGRAD || implicit_conv_enabled()) && net->prec != ZOO_PREC_FP32_SIMT && Wn.is_bf16 == dy_bf && dy_bf == bf &&
                L.cin >= 8 && tc_conv_supported(g, bf)) {
                // implicit transposed-conv GEMM: dX[pixel][c] = sum_(tap,o) dY[pixel_out(tap)][o] W[o][tap][c], relu' in the
                // epilogue -- no dcols matrix, no col2im pass
                Operand Wt{Wn.p, Wn.is_bf16, L.K, 0};
                ed.out = P.dout; ed.sm = L.cin; ed.sn = 1;
                if (P.relu) { ed.mask = P.out; ed.mask_bf16 = bf; ed.mask_sm = L.cin; ed.mask_sn = 1; }
                const long long Mi = rows * L.ih * L.iw;
                prof_label("dgrad:t%d[%lldx%dx%lld]", L.wt, Mi, L.cin, (long long)L.ksize * L.ksize * L.cout);
                ZOO_TRY(tc_conv_gemm(g, Wt, (int)Mi, L.cin, L.ksize * L.ksize * L.cout, ed, net->ws, net->ws_elems, st));
                continue;
            }
            ed.out = L.dcols;
            ZOO_TRY(gemm(dY, Wn, (int)M, (int)L.K, L.cout, ed, 0, net->ws, net->ws_elems, st));
            if (bf) ZOO_TRY(launch_col2im_nhwc<bf16>(L.dcols, P.dout, P.relu ? P.out : nullptr, rows, L, st));
            else ZOO_TRY(launch_col2im_nhwc<float>(L.dcols, P.dout, P.relu ? P.out : nullptr, rows, L, st));
        } else {
            if (li > 0) {
                LayerExec &P = ch.layers[li - 1];
                ed.out = P.dout;
                if (P.relu) { ed.mask = P.out; ed.mask_bf16 = bf; ed.mask_sm = L.K; ed.mask_sn = 1; }
            } else {
                ed.out = dx;
                if (dx_mask) { ed.mask = dx_mask; ed.mask_bf16 = bf; ed.mask_sm = L.K; ed.mask_sn = 1; }
            }
            ZOO_TRY(gemm(dY, Wn, (int)M, (int)L.K, L.cout, ed, 0, net->ws, net->ws_elems, st));
        }
        if (late_here) ZOO_TRY(net_late_bucket(net, st, sw, par));
    }
    return ZOO_OK;
}

zoo_status net_side_fork(zoo_net *net, cudaStream_t main) {
    ZOO_CUDA(cudaEventRecord(net->ev_fork, main));
    ZOO_CUDA(cudaStreamWaitEvent(net->side, net->ev_fork, 0));
    return ZOO_OK;
}
zoo_status net_side_join(zoo_net *net, cudaStream_t main) {
    ZOO_CUDA(cudaEventRecord(net->ev_join, net->side));
    ZOO_CUDA(cudaStreamWaitEvent(main, net->ev_join, 0));
    return ZOO_OK;
}

zoo_status net_mark_busy(zoo_net *net, cudaStream_t user) {
    if (!net->ev_busy) ZOO_CUDA(cudaEventCreateWithFlags(&net->ev_busy, cudaEventDisableTiming));
    ZOO_CUDA(cudaEventRecord(net->ev_busy, user));
    net->busy_pending = true;
    return ZOO_OK;
}
zoo_status net_wait_busy(zoo_net *net) {
    if (net->busy_pending) {
        ZOO_CUDA(cudaStreamWaitEvent(net->st, net->ev_busy, 0));
        net->busy_pending = false;
    }
    return ZOO_OK;
}

long long obs_elems(const zoo_net *net) { return (long long)net->desc.in_c * net->desc.in_h * net->desc.in_w; }

zoo_status net_refresh_shadow(zoo_net *net, cudaStream_t st) {
    if (!net->shadow) return ZOO_OK;
    cast_bf16_kernel<<<cdiv(net->n_flat / 4 + 1, 256), 256, 0, st>>>(net->params, net->shadow, net->n_flat);
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

zoo_status net_forward(zoo_net *net, const void *obs, int obs_dtype, int rows, const float *tau, int n_tau, cudaStream_t st, int bwd_rows) {
    const int bf = is_bf(net);
    const int in_kind = obs_dtype == ZOO_OBS_U8 ? 0 : 1;
    if (net->desc.kind == ZOO_NET_CHAIN) return chain_forward(net, net->a, obs, in_kind, rows, st, bwd_rows);
    if (net->desc.kind == ZOO_NET_DUELING) {
        ZOO_TRY(chain_forward(net, net->a, obs, in_kind, rows, st, bwd_rows));
        ZOO_TRY(chain_forward(net, net->b, net->a.out(), 2, rows, st));
        ZOO_TRY(chain_forward(net, net->c, net->a.out(), 2, rows, st));
        dueling_combine_kernel<<<cdiv(rows, 128), 128, 0, st>>>((const float *)net->b.out(), (const float *)net->c.out(),
                                                                net->out_final, rows, net->n_out);
        ZOO_LAUNCH_CHECK();
        return ZOO_OK;
    }
    // IQN -- iqn.jl:25-33
    ZOO_REQUIRE(tau && n_tau > 0 && n_tau <= net->max_q, "IQN forward: tau missing or n_tau %d > max_quantiles %lld", n_tau, net->max_q);
    const long long R = (long long)rows * n_tau;
    const int nem = (int)net->b.in_feat, F = (int)net->a.out_feat;
    ZOO_TRY(chain_forward(net, net->a, obs, in_kind, rows, st, bwd_rows));
    if (bf) embed_kernel<bf16><<<cdiv(R * nem, 256), 256, 0, st>>>(tau, (bf16 *)net->emb, R, nem);
    else embed_kernel<float><<<cdiv(R * nem, 256), 256, 0, st>>>(tau, (float *)net->emb, R, nem);
    ZOO_LAUNCH_CHECK();
    if (net->iqn_fused) {  // phi's last GEMM writes merged = psi[b] .* relu(W cos + b) from its epilogue (iqn.jl:27-30)
        LayerExec &Lp = net->b.layers.back();
        Lp.rowmul = net->a.out(); Lp.rowmul_div = n_tau; Lp.rowmul_ld = F; Lp.rowmul_nonneg = net->a.layers.back().relu ? 1 : 0;
        ZOO_TRY(chain_forward(net, net->b, net->emb, 2, R, st));
        return chain_forward(net, net->c, net->merged, 2, R, st);
    }
    ZOO_TRY(chain_forward(net, net->b, net->emb, 2, R, st));
    prof_label("iqn_merge[%lldx%d]", R, F);
    if (bf && F % 8 == 0) iqn_merge_vec_kernel<bf16, 8><<<cdiv(R * (F / 8), 256), 256, 0, st>>>((const bf16 *)net->a.out(), (const bf16 *)net->b.out(), (bf16 *)net->merged, R, F, n_tau);
    else if (!bf && F % 4 == 0) iqn_merge_vec_kernel<float, 4><<<cdiv(R * (F / 4), 256), 256, 0, st>>>((const float *)net->a.out(), (const float *)net->b.out(), (float *)net->merged, R, F, n_tau);
    else if (bf) iqn_merge_kernel<bf16><<<cdiv(R * F, 256), 256, 0, st>>>((const bf16 *)net->a.out(), (const bf16 *)net->b.out(), (bf16 *)net->merged, R, F, n_tau);
    else iqn_merge_kernel<float><<<cdiv(R * F, 256), 256, 0, st>>>((const float *)net->a.out(), (const float *)net->b.out(), (float *)net->merged, R, F, n_tau);
    ZOO_LAUNCH_CHECK();
    return chain_forward(net, net->c, net->merged, 2, R, st);
}

static zoo_status net_backward_impl(zoo_net *net, int obs_dtype, const void *obs, int rows, int n_tau, cudaStream_t st);
zoo_status net_backward(zoo_net *net, int obs_dtype, const void *obs, int rows, int n_tau, cudaStream_t st) {
    ZOO_REQUIRE(net->bwd_ready && net->grads, "backward: net has no gradient buffers (attach an optimiser/learner first)");
    zoo_status s = net_backward_impl(net, obs_dtype, obs, rows, n_tau, st);
    if (net->side && !g_prof_on) {  // the side lane always rejoins, also on error (an open fork would poison a stream capture)
        zoo_status j = net_side_join(net, st);
        if (s == ZOO_OK) s = j;
        if (net->comm_used) {
            net->comm_used = false;
            if (cudaEventRecord(net->ev_comm, net->comm) != cudaSuccess || cudaStreamWaitEvent(st, net->ev_comm, 0) != cudaSuccess) {
                set_err("backward: cannot rejoin the communication lane");
                if (s == ZOO_OK) s = ZOO_CUDA_ERROR;
            }
        }
    }
    return s;
}

static zoo_status net_backward_impl(zoo_net *net, int obs_dtype, const void *obs, int rows, int n_tau, cudaStream_t st) {
    const int bf = is_bf(net);
    const int in_kind = obs_dtype == ZOO_OBS_U8 ? 0 : 1;
    if (net->desc.kind == ZOO_NET_CHAIN) return chain_backward(net, net->a, obs, in_kind, rows, nullptr, nullptr, st);
    if (net->desc.kind == ZOO_NET_DUELING) {
        dueling_combine_bwd_kernel<<<cdiv(rows, 128), 128, 0, st>>>(net->dout_final, (float *)net->b.layers.back().dout,
                                                                    (float *)net->c.layers.back().dout, rows, net->n_out);
        ZOO_LAUNCH_CHECK();
        ZOO_TRY(chain_backward(net, net->b, net->a.out(), 2, rows, net->dh1, nullptr, st));
        ZOO_TRY(chain_backward(net, net->c, net->a.out(), 2, rows, net->dh2, nullptr, st));
        LayerExec &P = net->a.layers.back();
        long long n = (long long)rows * net->a.out_feat;
        if (bf) add_mask_kernel<bf16><<<cdiv(n, 256), 256, 0, st>>>((const bf16 *)net->dh1, (const bf16 *)net->dh2, P.relu ? (const bf16 *)P.out : nullptr, (bf16 *)P.dout, n);
        else add_mask_kernel<float><<<cdiv(n, 256), 256, 0, st>>>((const float *)net->dh1, (const float *)net->dh2, P.relu ? (const float *)P.out : nullptr, (float *)P.dout, n);
        ZOO_LAUNCH_CHECK();
        return chain_backward(net, net->a, obs, in_kind, rows, nullptr, nullptr, st);
    }
    const long long R = (long long)rows * n_tau;
    const int F = (int)net->a.out_feat;
    ZOO_TRY(chain_backward(net, net->c, net->merged, 2, R, net->dmerged, nullptr, st));
    LayerExec &Pphi = net->b.layers.back();
    LayerExec &Ppsi = net->a.layers.back();
    long long n = (long long)rows * F;
    prof_label("iqn_merge_bwd[%lldx%d]", (long long)rows * n_tau, F);
    if (net->iqn_fused) {
        float *dbias = net->grads + net->tensors[Pphi.bt].offset;
        const dim3 blk(32, 8);
        // ZOO_MERGE_BWD_VAR: 0 = the plain unrolled loop (round-2 first half), 1 = pipelined groups of 2 rows at 2 CTAs / SM (default: 258 us = 5.9 TB/s against 335-359 us for 0),
        // 2 = groups of 4 rows at 2 CTAs / SM (279 us, spills), 3 = groups of 2 rows at 3 CTAs / SM (314 us, spills)
        static const int var = [] { const char *e = getenv("ZOO_MERGE_BWD_VAR"); return e ? atoi(e) : 1; }();
#define MBWD(T, V, G, MB) iqn_merge_bwd_fused_kernel<T, V, G, MB><<<dim3(cdiv(F / V, 32), cdiv(rows, 8)), blk, 0, st>>>((const T *)net->dmerged, (const T *)Ppsi.out, (const T *)net->merged, (T *)Pphi.dout, (T *)Ppsi.dout, dbias, rows, F, n_tau)
        if (bf && F % 8 == 0) { if (var == 0) MBWD(bf16, 8, 0, 3); else if (var == 1) MBWD(bf16, 8, 2, 2); else if (var == 3) MBWD(bf16, 8, 2, 3); else MBWD(bf16, 8, 4, 2); }
        else if (!bf && F % 4 == 0) { if (var == 0) MBWD(float, 4, 0, 3); else MBWD(float, 4, 4, 2); }
        else if (bf) MBWD(bf16, 1, 0, 3);
        else MBWD(float, 1, 0, 3);
#undef MBWD
        ZOO_LAUNCH_CHECK();
        ZOO_TRY(chain_backward(net, net->b, net->emb, 2, R, nullptr, nullptr, st, (int)net->b.layers.size() - 1));
        return chain_backward(net, net->a, obs, in_kind, rows, nullptr, nullptr, st);
    }
    if (bf && F % 8 == 0) iqn_merge_bwd_vec_kernel<bf16, 8><<<cdiv(n / 8, 128), 128, 0, st>>>((const bf16 *)net->dmerged, (const bf16 *)Ppsi.out, (const bf16 *)Pphi.out, (bf16 *)Pphi.dout, (bf16 *)Ppsi.dout, rows, F, n_tau, Pphi.relu, Ppsi.relu);
    else if (!bf && F % 4 == 0) iqn_merge_bwd_vec_kernel<float, 4><<<cdiv(n / 4, 128), 128, 0, st>>>((const float *)net->dmerged, (const float *)Ppsi.out, (const float *)Pphi.out, (float *)Pphi.dout, (float *)Ppsi.dout, rows, F, n_tau, Pphi.relu, Ppsi.relu);
    else if (bf) iqn_merge_bwd_kernel<bf16><<<cdiv(n, 256), 256, 0, st>>>((const bf16 *)net->dmerged, (const bf16 *)Ppsi.out, (const bf16 *)Pphi.out, (bf16 *)Pphi.dout, (bf16 *)Ppsi.dout, rows, F, n_tau, Pphi.relu, Ppsi.relu);
    else iqn_merge_bwd_kernel<float><<<cdiv(n, 256), 256, 0, st>>>((const float *)net->dmerged, (const float *)Ppsi.out, (const float *)Pphi.out, (float *)Pphi.dout, (float *)Ppsi.dout, rows, F, n_tau, Pphi.relu, Ppsi.relu);
    ZOO_LAUNCH_CHECK();
    ZOO_TRY(chain_backward(net, net->b, net->emb, 2, R, nullptr, nullptr, st));
    return chain_backward(net, net->a, obs, in_kind, rows, nullptr, nullptr, st);
}

// ---------------------------------------------------------------- construction ------------
struct ParseState {
    int c, h, w;        // spatial dims while spatial
    bool spatial;
    long long feat;     // feature count once flat
    bool flat_pending;  // the current feature axis is a flattened conv output (needs the (c,h,w)->(h,w,c) permutation)
    bool raw;           // next layer reads the raw observation
    bool scale;         // pending x ./ 255
};

static long long align_up(long long x, long long a) { return (x + a - 1) / a * a; }

static zoo_status parse_chain(zoo_net *net, const zoo_layer *Ls, int n, ParseState &S, Chain &ch, bool is_output_chain, const char *name) {
    ch.in_feat = S.spatial ? (long long)S.c * S.h * S.w : S.feat;
    for (int i = 0; i < n; ++i) {
        const zoo_layer &l = Ls[i];
        if (l.kind == ZOO_LAYER_SCALE255) {
            ZOO_REQUIRE(S.raw && i == 0, "%s: the x./255 layer must be the first layer of the first chain", name);
            S.scale = true;
        } else if (l.kind == ZOO_LAYER_CONV) {
            ZOO_REQUIRE(S.spatial, "%s[%d]: conv after flatten", name, i);
            ZOO_REQUIRE(l.n_in == S.c, "%s[%d]: conv expects %d input channels, got %d", name, i, l.n_in, S.c);
            ZOO_REQUIRE(l.ksize > 0 && l.stride > 0 && l.pad >= 0 && l.n_out > 0, "%s[%d]: bad conv geometry", name, i);
            LayerExec L{};
            L.kind = ZOO_LAYER_CONV; L.cin = l.n_in; L.cout = l.n_out; L.ksize = l.ksize; L.stride = l.stride; L.pad = l.pad;
            L.relu = l.relu; L.ih = S.h; L.iw = S.w;
            L.oh = (S.h + 2 * l.pad - l.ksize) / l.stride + 1;
            L.ow = (S.w + 2 * l.pad - l.ksize) / l.stride + 1;
            ZOO_REQUIRE(L.oh > 0 && L.ow > 0, "%s[%d]: conv output is empty", name, i);
            L.first_raw = S.raw; L.scale255 = S.scale; L.K = (long long)l.ksize * l.ksize * l.n_in; L.opr = (long long)L.oh * L.ow;
            if (S.raw && net->obs_nhwc) { L.ih += 2 * l.pad; L.iw += 2 * l.pad; L.pad = 0; }  // reads the zero-padded NHWC copy
            TensorInfo w{}; w.kind = 1; w.n_out = l.n_out; w.n_in = l.n_in; w.ksize = l.ksize; w.conv_nhwc = !S.raw || net->obs_nhwc;
            w.count = (long long)l.n_out * L.K;
            TensorInfo b{}; b.kind = 2; b.n_out = l.n_out; b.count = l.n_out;
            L.wt = (int)net->tensors.size(); net->tensors.push_back(w);
            L.bt = (int)net->tensors.size(); net->tensors.push_back(b);
            ch.layers.push_back(L);
            S.c = l.n_out; S.h = L.oh; S.w = L.ow; S.raw = false; S.scale = false;
        } else if (l.kind == ZOO_LAYER_FLATTEN) {
            if (S.spatial) {
                ZOO_REQUIRE(!S.raw, "%s[%d]: flatten of the raw observation is not supported (use in_h = in_w = 1)", name, i);
                S.spatial = false; S.feat = (long long)S.c * S.h * S.w; S.flat_pending = (S.h * S.w > 1);
                net->fc = S.c; net->fh = S.h; net->fw = S.w;
            }
        } else if (l.kind == ZOO_LAYER_DENSE) {
            ZOO_REQUIRE(!S.spatial, "%s[%d]: dense needs a flatten after the convs", name, i);
            ZOO_REQUIRE(l.n_in == S.feat, "%s[%d]: dense expects %lld inputs, got %d", name, i, S.feat, l.n_in);
            ZOO_REQUIRE(l.n_out > 0, "%s[%d]: bad dense size", name, i);
            LayerExec L{};
            L.kind = ZOO_LAYER_DENSE; L.cin = l.n_in; L.cout = l.n_out; L.relu = l.relu; L.first_raw = S.raw; L.scale255 = 0;
            ZOO_REQUIRE(!S.scale, "%s[%d]: x./255 directly before a dense layer is not supported", name, i);
            L.K = l.n_in; L.opr = 1;
            TensorInfo w{}; w.kind = 0; w.n_out = l.n_out; w.n_in = l.n_in; w.perm_in = S.flat_pending; w.fc = net->fc; w.fh = net->fh; w.fw = net->fw;
            w.count = (long long)l.n_out * l.n_in;
            TensorInfo b{}; b.kind = 2; b.n_out = l.n_out; b.count = l.n_out; b.fc = net->fc; b.fh = net->fh; b.fw = net->fw;
            L.wt = (int)net->tensors.size(); net->tensors.push_back(w);
            L.bt = (int)net->tensors.size(); net->tensors.push_back(b);
            ch.layers.push_back(L);
            S.feat = l.n_out; S.flat_pending = false; S.raw = false;
        } else {
            set_err("%s[%d]: unknown layer kind %d", name, i, l.kind);
            return ZOO_BAD_ARG;
        }
    }
    ZOO_REQUIRE(!ch.layers.empty(), "%s: chain has no parameter layers", name);
    ch.out_feat = S.spatial ? (long long)S.c * S.h * S.w : S.feat;
    if (is_output_chain) {
        ZOO_REQUIRE(ch.layers.back().kind == ZOO_LAYER_DENSE, "%s: the output layer must be dense", name);
        ch.layers.back().out_fp32 = 1;
    }
    return ZOO_OK;
}

static zoo_status alloc_chain(zoo_net *net, Chain &ch, long long rows_cap, float *final_out, void *last_out_alias = nullptr) {
    ch.rows_cap = rows_cap;
    for (auto &L : ch.layers) {
        long long M = rows_cap * L.opr;
        if (L.kind == ZOO_LAYER_CONV) ZOO_CUDA(cudaMalloc(&L.cols, (size_t)M * L.K * net->esz));
        if (last_out_alias && &L == &ch.layers.back()) { L.out = last_out_alias; continue; }
        if (L.out_fp32 && final_out) L.out = final_out;
        else ZOO_CUDA(cudaMalloc(&L.out, (size_t)M * L.cout * (L.out_fp32 ? 4 : net->esz)));
    }
    return ZOO_OK;
}

static zoo_status alloc_chain_bwd(zoo_net *net, Chain &ch, float *final_dout) {
    for (size_t i = 0; i < ch.layers.size(); ++i) {
        auto &L = ch.layers[i];
        long long M = ch.rows_cap * L.opr;
        if (L.out_fp32 && final_dout) L.dout = final_dout;
        else ZOO_CUDA(cudaMalloc(&L.dout, (size_t)M * L.cout * (L.out_fp32 ? 4 : net->esz)));
        if (L.kind == ZOO_LAYER_CONV && i > 0) ZOO_CUDA(cudaMalloc(&L.dcols, (size_t)M * L.K * net->esz));
        // wide fp32 output layer of a bf16 net: the bf16, row-padded copy of its incoming gradient (head_dy_to_bf16_kernel)
        if (L.kind == ZOO_LAYER_DENSE && L.out_fp32 && is_bf(net) && L.cout > 8) ZOO_CUDA(cudaMalloc(&L.dcols, (size_t)M * ((L.cout + 7) & ~7) * sizeof(bf16)));
    }
    return ZOO_OK;
}

zoo_status net_alloc_backward(zoo_net *net) {
    if (net->bwd_ready) return ZOO_OK;
    // 64 extra floats: grads[n_flat] carries the loss so data-parallel needs one all-reduce for both
    ZOO_CUDA(cudaMalloc(&net->grads, (net->n_flat + 64) * sizeof(float)));
    ZOO_CUDA(cudaMemset(net->grads, 0, (net->n_flat + 64) * sizeof(float)));
    const long long rq = net->max_batch * std::max<long long>(1, net->max_q);
    const long long out_rows = net->desc.kind == ZOO_NET_IQN ? rq : net->a.rows_cap;
    ZOO_CUDA(cudaMalloc(&net->dout_final, (size_t)out_rows * net->n_out * sizeof(float)));
    ZOO_CUDA(cudaMemset(net->dout_final, 0, (size_t)out_rows * net->n_out * sizeof(float)));
    if (net->desc.kind == ZOO_NET_CHAIN) ZOO_TRY(alloc_chain_bwd(net, net->a, net->dout_final));
    else if (net->desc.kind == ZOO_NET_IQN) {
        ZOO_TRY(alloc_chain_bwd(net, net->a, nullptr));
        ZOO_TRY(alloc_chain_bwd(net, net->b, nullptr));
        ZOO_TRY(alloc_chain_bwd(net, net->c, net->dout_final));
        ZOO_CUDA(cudaMalloc(&net->dmerged, (size_t)rq * net->a.out_feat * net->esz));
    } else {
        ZOO_TRY(alloc_chain_bwd(net, net->a, nullptr));
        ZOO_TRY(alloc_chain_bwd(net, net->b, nullptr));
        ZOO_TRY(alloc_chain_bwd(net, net->c, nullptr));
        ZOO_CUDA(cudaMalloc(&net->dh1, (size_t)net->a.rows_cap * net->a.out_feat * net->esz));
        ZOO_CUDA(cudaMalloc(&net->dh2, (size_t)net->a.rows_cap * net->a.out_feat * net->esz));
    }
    ZOO_CUDA(cudaStreamCreateWithFlags(&net->side, cudaStreamNonBlocking));
    ZOO_CUDA(cudaEventCreateWithFlags(&net->ev_fork, cudaEventDisableTiming));
    ZOO_CUDA(cudaEventCreateWithFlags(&net->ev_join, cudaEventDisableTiming));
    ZOO_CUDA(cudaEventCreateWithFlags(&net->ev_late, cudaEventDisableTiming));
    ZOO_CUDA(cudaEventCreateWithFlags(&net->ev_late2, cudaEventDisableTiming));
    ZOO_CUDA(cudaEventCreateWithFlags(&net->ev_comm, cudaEventDisableTiming));
    ZOO_CUDA(cudaStreamCreateWithFlags(&net->comm, cudaStreamNonBlocking));
    net->ws_side_elems = ((size_t)2 << 20) + 4096;
    ZOO_CUDA(cudaMalloc(&net->ws_side, net->ws_side_elems * sizeof(float)));
    ZOO_CUDA(cudaMemset(net->ws_side, 0, net->ws_side_elems * sizeof(float)));
    net->bwd_ready = true;
    return ZOO_OK;
}

}  // namespace zoo

using namespace zoo;

static zoo_status ensure_stage(zoo_net *net, size_t n) {
    if (net->stage_elems >= n) return ZOO_OK;
    if (net->stage) cudaFree(net->stage);
    ZOO_CUDA(cudaMalloc(&net->stage, n * sizeof(float)));
    net->stage_elems = n;
    return ZOO_OK;
}

extern "C" {

zoo_status zoo_net_create(const zoo_net_desc *d, zoo_net **out) {
    ZOO_REQUIRE(d && out, "zoo_net_create: null argument");
    ZOO_TRY(require_sm100());
    ZOO_REQUIRE(d->precision >= ZOO_PREC_FP32 && d->precision <= ZOO_PREC_FP32_SIMT, "zoo_net_create: bad precision %d", d->precision);
    ZOO_REQUIRE(d->max_batch > 0 && d->in_c > 0 && d->in_h > 0 && d->in_w > 0, "zoo_net_create: bad observation shape / max_batch");
    ZOO_REQUIRE(d->n_a >= 1 && d->n_a <= ZOO_MAX_LAYERS && d->n_b >= 0 && d->n_b <= ZOO_MAX_LAYERS && d->n_c >= 0 && d->n_c <= ZOO_MAX_LAYERS,
                "zoo_net_create: bad layer counts");
    zoo_net *net = new zoo_net();
    net->desc = *d;
    net->prec = d->precision;
    net->esz = d->precision == ZOO_PREC_BF16 ? 2 : 4;
    net->tf32_passes = d->precision == ZOO_PREC_FP32 ? 3 : (d->precision == ZOO_PREC_TF32 ? 1 : 0);
    net->max_batch = d->max_batch;
    net->max_q = d->kind == ZOO_NET_IQN ? d->max_quantiles : 0;
    // 4-channel frame stacks are first turned into NHWC in the activation type (x ./ 255 folded in), so the first conv is
    // an ordinary inner conv
    net->obs_nhwc = d->in_c == 4 && !(d->in_h == 1 && d->in_w == 1);
    ParseState S{d->in_c, d->in_h, d->in_w, !(d->in_h == 1 && d->in_w == 1), (long long)d->in_c, false, true, false};
    zoo_status s = ZOO_OK;
    if (d->kind == ZOO_NET_CHAIN) {
        s = parse_chain(net, d->a, d->n_a, S, net->a, true, "chain");
        if (s == ZOO_OK) net->n_out = (int)net->a.out_feat;
    } else if (d->kind == ZOO_NET_IQN) {
        if (d->max_quantiles <= 0 || d->n_b < 1 || d->n_c < 1) { set_err("zoo_net_create: IQN needs psi, phi, header and max_quantiles"); s = ZOO_BAD_ARG; }
        if (s == ZOO_OK) s = parse_chain(net, d->a, d->n_a, S, net->a, false, "psi");
        if (s == ZOO_OK) {
            if (S.spatial) {  // psi ending without an explicit flatten: flatten implicitly
                S.spatial = false; S.feat = (long long)S.c * S.h * S.w; S.flat_pending = S.h * S.w > 1;
                net->fc = S.c; net->fh = S.h; net->fw = S.w;
            }
            const bool pend = S.flat_pending;
            const long long F = S.feat;
            ParseState Sb{0, 1, 1, false, (long long)d->b[0].n_in, false, false, false};
            s = parse_chain(net, d->b, d->n_b, Sb, net->b, false, "phi");
            if (s == ZOO_OK && net->b.out_feat != F) { set_err("IQN: phi output %lld != psi output %lld", net->b.out_feat, F); s = ZOO_BAD_ARG; }
            if (s == ZOO_OK && pend) {  // phi's output axis lives in the flatten space
                net->tensors[net->b.layers.back().wt].perm_out = 1;
                TensorInfo &wb = net->tensors[net->b.layers.back().wt]; wb.fc = net->fc; wb.fh = net->fh; wb.fw = net->fw;
                TensorInfo &bb = net->tensors[net->b.layers.back().bt]; bb.perm_out = 1; bb.fc = net->fc; bb.fh = net->fh; bb.fw = net->fw;
            }
            ParseState Sc{0, 1, 1, false, F, pend, false, false};
            if (s == ZOO_OK) s = parse_chain(net, d->c, d->n_c, Sc, net->c, true, "header");
            if (s == ZOO_OK) net->n_out = (int)net->c.out_feat;
        }
    } else if (d->kind == ZOO_NET_DUELING) {
        if (d->n_b < 1 || d->n_c < 1) { set_err("zoo_net_create: dueling needs base, val, adv"); s = ZOO_BAD_ARG; }
        if (s == ZOO_OK) s = parse_chain(net, d->a, d->n_a, S, net->a, false, "base");
        if (s == ZOO_OK) {
            if (S.spatial) { S.spatial = false; S.feat = (long long)S.c * S.h * S.w; S.flat_pending = S.h * S.w > 1; net->fc = S.c; net->fh = S.h; net->fw = S.w; }
            ParseState Sb = S, Sc = S;
            s = parse_chain(net, d->b, d->n_b, Sb, net->b, true, "val");
            if (s == ZOO_OK) s = parse_chain(net, d->c, d->n_c, Sc, net->c, true, "adv");
            if (s == ZOO_OK && net->b.out_feat != 1) { set_err("dueling: val must have one output"); s = ZOO_BAD_ARG; }
            if (s == ZOO_OK) net->n_out = (int)net->c.out_feat;
        }
    } else {
        set_err("zoo_net_create: unknown net kind %d", d->kind);
        s = ZOO_BAD_ARG;
    }
    if (s != ZOO_OK) { delete net; return s; }

    long long off = 0;
    for (auto &t : net->tensors) { t.offset = off; off = align_up(off + t.count, 64); net->n_params += t.count; }
    net->n_flat = align_up(off, 64);
    ZOO_CUDA(cudaStreamCreateWithFlags(&net->st, cudaStreamNonBlocking));
    ZOO_CUDA(cudaMalloc(&net->params, net->n_flat * sizeof(float)));
    ZOO_CUDA(cudaMemset(net->params, 0, net->n_flat * sizeof(float)));
    if (net->prec == ZOO_PREC_BF16) {
        ZOO_CUDA(cudaMalloc(&net->shadow, net->n_flat * sizeof(bf16)));
        ZOO_CUDA(cudaMemset(net->shadow, 0, net->n_flat * sizeof(bf16)));
    }
    const long long rq = net->max_batch * std::max<long long>(1, net->max_q);
    const long long rows_a = d->kind == ZOO_NET_IQN ? net->max_batch : 2 * net->max_batch;  // online nets evaluate [s; s'] in one pass
    const long long out_rows = d->kind == ZOO_NET_IQN ? rq : rows_a;
    ZOO_CUDA(cudaMalloc(&net->out_final, (size_t)out_rows * net->n_out * sizeof(float)));
    if (d->kind == ZOO_NET_CHAIN) ZOO_TRY(alloc_chain(net, net->a, rows_a, net->out_final));
    else if (d->kind == ZOO_NET_IQN) {
        ZOO_CUDA(cudaMalloc(&net->merged, (size_t)rq * net->a.out_feat * net->esz));
        {   // fused merge: needs relu on both factors (so that merged / psi and [merged > 0] recover phi's output and mask)
            const char *e = getenv("ZOO_IQN_FUSE");
            net->iqn_fused = net->a.layers.back().relu && net->b.layers.back().relu && net->b.layers.back().kind == ZOO_LAYER_DENSE && !(e && e[0] == '0');
        }
        ZOO_TRY(alloc_chain(net, net->a, rows_a, nullptr));
        ZOO_TRY(alloc_chain(net, net->b, rq, nullptr, net->iqn_fused ? net->merged : nullptr));
        ZOO_TRY(alloc_chain(net, net->c, rq, net->out_final));
        ZOO_CUDA(cudaMalloc(&net->emb, (size_t)rq * net->b.in_feat * net->esz));
        ZOO_CUDA(cudaMalloc(&net->tau_dev, (size_t)rq * sizeof(float)));
    } else {
        ZOO_TRY(alloc_chain(net, net->a, rows_a, nullptr));
        ZOO_TRY(alloc_chain(net, net->b, rows_a, nullptr));
        ZOO_TRY(alloc_chain(net, net->c, rows_a, nullptr));
    }
    // split-K workspace (all-zero between launches): voluntary splits cover < 74 output tiles of 128 x 128 fp32; the
    // fp32-grade mode also splits every contraction longer than 1024 (dense layers fed by the flattened conv stack)
    size_t ws_need = (size_t)2 << 20;
    for (Chain *ch : {&net->a, &net->b, &net->c})
        for (auto &L : ch->layers)
            if (L.K > 1024) ws_need = std::max<size_t>(ws_need, (size_t)(ch->rows_cap * L.opr * L.cout));
    net->ws_elems = ws_need + 4096;
    ZOO_CUDA(cudaMalloc(&net->ws, net->ws_elems * sizeof(float)));
    ZOO_CUDA(cudaMemset(net->ws, 0, net->ws_elems * sizeof(float)));
    ZOO_CUDA(cudaMalloc(&net->obs, (size_t)rows_a * obs_elems(net) * sizeof(float)));
    if (net->obs_nhwc) {
        const LayerExec &L0 = net->a.layers.front();
        const size_t nb = (size_t)rows_a * L0.ih * L0.iw * 4 * sizeof(float);
        ZOO_CUDA(cudaMalloc(&net->obs_nhwc_buf, nb));
        ZOO_CUDA(cudaMemset(net->obs_nhwc_buf, 0, nb));  // the zero border is never written again
    }
    *out = net;
    return ZOO_OK;
}

static void free_chain(Chain &ch, float *fo, float *fd, void *alias = nullptr) {
    for (auto &L : ch.layers) {
        if (L.cols) cudaFree(L.cols);
        if (L.out && L.out != fo && L.out != alias) cudaFree(L.out);
        if (L.dout && L.dout != fd) cudaFree(L.dout);
        if (L.dcols) cudaFree(L.dcols);
    }
}

zoo_status zoo_net_destroy(zoo_net *net) {
    if (!net) return ZOO_OK;
    cudaStreamSynchronize(net->st);
    free_chain(net->a, net->out_final, net->dout_final);
    free_chain(net->b, net->out_final, net->dout_final, net->merged);
    free_chain(net->c, net->out_final, net->dout_final);
    if (net->grads_nccl) { dp_free_registered(net->grads); net->grads = nullptr; }
    void *ptrs[] = {net->params, net->shadow, net->grads, net->out_final, net->dout_final, net->emb, net->merged, net->dmerged,
                    net->tau_dev, net->dh1, net->dh2, net->ws, net->obs, net->stage, net->ws_side, net->obs_nhwc_buf};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (net->side) { cudaStreamSynchronize(net->side); cudaStreamDestroy(net->side); }
    if (net->ev_fork) cudaEventDestroy(net->ev_fork);
    if (net->ev_join) cudaEventDestroy(net->ev_join);
    if (net->ev_late) cudaEventDestroy(net->ev_late);
    if (net->ev_late2) cudaEventDestroy(net->ev_late2);
    if (net->ev_comm) cudaEventDestroy(net->ev_comm);
    if (net->ev_busy) cudaEventDestroy(net->ev_busy);
    if (net->comm) { cudaStreamSynchronize(net->comm); cudaStreamDestroy(net->comm); }
    cudaStreamDestroy(net->st);
    delete net;
    return ZOO_OK;
}

zoo_status zoo_net_num_tensors(const zoo_net *net, int32_t *n) {
    ZOO_REQUIRE(net && n, "zoo_net_num_tensors: null argument");
    *n = (int32_t)net->tensors.size();
    return ZOO_OK;
}
zoo_status zoo_net_tensor_size(const zoo_net *net, int32_t id, int64_t *n) {
    ZOO_REQUIRE(net && n && id >= 0 && id < (int)net->tensors.size(), "zoo_net_tensor_size: bad tensor id %d", id);
    *n = net->tensors[id].count;
    return ZOO_OK;
}
zoo_status zoo_net_num_params(const zoo_net *net, int64_t *n) {
    ZOO_REQUIRE(net && n, "zoo_net_num_params: null argument");
    *n = net->n_params;
    return ZOO_OK;
}
zoo_status zoo_net_out_dim(const zoo_net *net, int32_t *n) {
    ZOO_REQUIRE(net && n, "zoo_net_out_dim: null argument");
    *n = net->n_out;
    return ZOO_OK;
}

// shared by params / grads / optimiser state: flat = device buffer in the internal layout
zoo_status zoo_internal_tensor_io(zoo_net *net, float *flat, int32_t id, float *host, int64_t n, int to_device) {
    ZOO_REQUIRE(net && flat && host && id >= 0 && id < (int)net->tensors.size(), "tensor io: bad tensor id %d", id);
    const TensorInfo &t = net->tensors[id];
    ZOO_REQUIRE(n == t.count, "tensor io: tensor %d has %lld elements, caller passed %lld", id, t.count, (long long)n);
    ZOO_TRY(ensure_stage(net, (size_t)t.count));
    ZOO_TRY(net_wait_busy(net));
    if (to_device) {
        ZOO_CUDA(cudaMemcpyAsync(net->stage, host, t.count * sizeof(float), cudaMemcpyHostToDevice, net->st));
        permute_param_kernel<<<cdiv(t.count, 256), 256, 0, net->st>>>(flat + t.offset, net->stage, t, 0);
        ZOO_LAUNCH_CHECK();
    } else {
        permute_param_kernel<<<cdiv(t.count, 256), 256, 0, net->st>>>(flat + t.offset, net->stage, t, 1);
        ZOO_LAUNCH_CHECK();
        ZOO_CUDA(cudaMemcpyAsync(host, net->stage, t.count * sizeof(float), cudaMemcpyDeviceToHost, net->st));
    }
    ZOO_CUDA(cudaStreamSynchronize(net->st));
    return ZOO_OK;
}

zoo_status zoo_net_set_params(zoo_net *net, int32_t id, const float *host, int64_t n) {
    ZOO_REQUIRE(net, "zoo_net_set_params: null net");
    ZOO_TRY(zoo_internal_tensor_io(net, net->params, id, const_cast<float *>(host), n, 1));
    ZOO_TRY(net_refresh_shadow(net, net->st));
    ZOO_CUDA(cudaStreamSynchronize(net->st));
    return ZOO_OK;
}
zoo_status zoo_net_get_params(const zoo_net *net, int32_t id, float *host, int64_t n) {
    ZOO_REQUIRE(net, "zoo_net_get_params: null net");
    zoo_net *m = const_cast<zoo_net *>(net);
    return zoo_internal_tensor_io(m, m->params, id, host, n, 0);
}

zoo_status zoo_net_copy(zoo_net *dst, const zoo_net *src) {
    ZOO_REQUIRE(dst && src, "zoo_net_copy: null argument");
    ZOO_REQUIRE(dst->n_flat == src->n_flat && dst->tensors.size() == src->tensors.size(), "zoo_net_copy: nets have different structure");
    ZOO_TRY(net_wait_busy(dst));
    if (src->busy_pending) ZOO_CUDA(cudaStreamWaitEvent(dst->st, src->ev_busy, 0));  // (src stays marked: its own stream has not waited)
    ZOO_CUDA(cudaMemcpyAsync(dst->params, src->params, src->n_flat * sizeof(float), cudaMemcpyDeviceToDevice, dst->st));
    ZOO_TRY(net_refresh_shadow(dst, dst->st));
    ZOO_CUDA(cudaStreamSynchronize(dst->st));
    return ZOO_OK;
}

zoo_status zoo_net_forward(zoo_net *net, const void *obs_host, int32_t obs_dtype, int32_t B, const float *tau_host, int32_t n_tau, float *out_host) {
    ZOO_REQUIRE(net && obs_host && out_host && B > 0, "zoo_net_forward: bad argument");
    ZOO_REQUIRE(obs_dtype == ZOO_OBS_U8 || obs_dtype == ZOO_OBS_F32, "zoo_net_forward: bad obs dtype");
    ZOO_REQUIRE(B <= net->a.rows_cap, "zoo_net_forward: B %d exceeds capacity %lld", B, net->a.rows_cap);
    size_t ob = (size_t)B * obs_elems(net) * (obs_dtype == ZOO_OBS_U8 ? 1 : 4);
    ZOO_TRY(net_wait_busy(net));
    ZOO_CUDA(cudaMemcpyAsync(net->obs, obs_host, ob, cudaMemcpyHostToDevice, net->st));
    long long R = B;
    if (net->desc.kind == ZOO_NET_IQN) {
        ZOO_REQUIRE(tau_host && n_tau > 0 && n_tau <= net->max_q && B <= net->max_batch, "zoo_net_forward: IQN needs tau [B][n_tau], n_tau <= max_quantiles");
        R = (long long)B * n_tau;
        ZOO_CUDA(cudaMemcpyAsync(net->tau_dev, tau_host, R * sizeof(float), cudaMemcpyHostToDevice, net->st));
    }
    ZOO_TRY(net_forward(net, net->obs, obs_dtype, B, net->tau_dev, n_tau, net->st, 0));
    ZOO_CUDA(cudaMemcpyAsync(out_host, net->out_final, (size_t)R * net->n_out * sizeof(float), cudaMemcpyDeviceToHost, net->st));
    ZOO_CUDA(cudaStreamSynchronize(net->st));
    return ZOO_OK;
}

zoo_status zoo_test_net_activation(zoo_net *net, int32_t chain, int32_t layer, int64_t rows, void *host, int64_t host_bytes, int32_t *is_bf16) {
    ZOO_REQUIRE(net && host && is_bf16 && chain >= 0 && chain <= 2, "zoo_test_net_activation: bad argument");
    const Chain &ch = chain == 0 ? net->a : (chain == 1 ? net->b : net->c);
    ZOO_REQUIRE(layer >= 0 && layer < (int)ch.layers.size() && rows > 0 && rows <= ch.rows_cap, "zoo_test_net_activation: chain %d has %d layers, %lld rows",
                chain, (int)ch.layers.size(), ch.rows_cap);
    const LayerExec &L = ch.layers[layer];
    const int esz = L.out_fp32 ? 4 : net->esz;
    const size_t need = (size_t)rows * L.opr * L.cout * esz;
    ZOO_REQUIRE((size_t)host_bytes == need, "zoo_test_net_activation: layer holds %zu bytes for %lld rows, caller passed %lld", need, (long long)rows, (long long)host_bytes);
    ZOO_CUDA(cudaDeviceSynchronize());
    ZOO_CUDA(cudaMemcpy(host, L.out, need, cudaMemcpyDeviceToHost));
    *is_bf16 = esz == 2;
    return ZOO_OK;
}

}  // extern "C"
