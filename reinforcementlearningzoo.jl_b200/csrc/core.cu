// Error reporting, device checks and stand-alone test hooks of libzoo_sm100.
#include "common.cuh"

namespace zoo {

static thread_local std::string g_err;
thread_local int g_launches = 0;

void set_err(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

thread_local bool g_prof_on = false;
static thread_local cudaStream_t g_prof_stream = nullptr;
static thread_local std::vector<cudaEvent_t> g_prof_events;
static thread_local std::vector<std::string> g_prof_names;
static thread_local std::vector<float> g_prof_ms;
static thread_local char g_prof_lbl[96] = "";

void prof_label(const char *fmt, ...) {
    if (!g_prof_on) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_prof_lbl, sizeof(g_prof_lbl), fmt, ap);
    va_end(ap);
}

void prof_mark(const char *func, int line) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, g_prof_stream);
    g_prof_events.push_back(e);
    char buf[192];
    snprintf(buf, sizeof(buf), "%s%s%s:%d", g_prof_lbl, g_prof_lbl[0] ? "|" : "", func, line);
    g_prof_names.push_back(buf);
}

__global__ void prof_delay_kernel(unsigned long long ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(1000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < ns);
}

// No CPU fallback and no other architecture: everything here is sm_100a SASS.
zoo_status require_sm100() {
    static int cached = -1;
    if (cached < 0) {
        int dev = 0;
        cudaDeviceProp prop;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
            cudaGetLastError();
            set_err("libzoo_sm100: no CUDA device available (there is no CPU fallback)");
            return ZOO_CUDA_ERROR;
        }
        cached = prop.major * 10 + prop.minor;
    }
    if (cached / 10 != 10) {
        set_err("libzoo_sm100: device is sm_%d, this library only contains sm_100a code", cached);
        return ZOO_UNSUPPORTED_ARCH;
    }
    return ZOO_OK;
}

__global__ void f32_to_bf16_kernel(const float *in, bf16 *out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}

}  // namespace zoo

using namespace zoo;

extern "C" {

const char *zoo_last_error(void) { return g_err.c_str(); }
int32_t zoo_version(void) { return 100; }

int32_t zoo_device_arch(void) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return prop.major * 10 + prop.minor;
}

zoo_status zoo_set_device(int32_t device) {
    ZOO_CUDA(cudaSetDevice(device));
    return ZOO_OK;
}

zoo_status zoo_prof_begin(void *stream, int32_t delay_us) {
    for (cudaEvent_t e : g_prof_events) cudaEventDestroy(e);
    g_prof_events.clear(); g_prof_names.clear(); g_prof_ms.clear();
    g_prof_stream = (cudaStream_t)stream;
    g_prof_lbl[0] = 0;
    // hold the stream busy while the host enqueues, so event differences are device time, not launch-rate gaps
    if (delay_us > 0) prof_delay_kernel<<<1, 1, 0, g_prof_stream>>>((unsigned long long)delay_us * 1000ull);
    g_prof_on = true;
    prof_mark("begin", 0);
    return ZOO_OK;
}

zoo_status zoo_prof_end(int32_t *n) {
    ZOO_REQUIRE(n, "zoo_prof_end: null argument");
    g_prof_on = false;
    ZOO_CUDA(cudaStreamSynchronize(g_prof_stream));
    g_prof_ms.assign(g_prof_events.size(), 0.f);
    for (size_t i = 1; i < g_prof_events.size(); ++i) cudaEventElapsedTime(&g_prof_ms[i], g_prof_events[i - 1], g_prof_events[i]);
    *n = (int32_t)g_prof_events.size() - 1;
    return ZOO_OK;
}

zoo_status zoo_prof_get(int32_t i, char *name, int32_t cap, float *ms) {
    ZOO_REQUIRE(name && ms && i >= 0 && i + 1 < (int)g_prof_events.size() && cap > 0, "zoo_prof_get: bad index");
    snprintf(name, cap, "%s", g_prof_names[i + 1].c_str());
    *ms = g_prof_ms[i + 1];
    return ZOO_OK;
}

zoo_status zoo_test_gemm(int32_t engine, int32_t M, int32_t N, int32_t K, const float *A, int32_t a_kmajor, const float *B,
                         int32_t b_kmajor, float *D, int32_t split_k) {
    ZOO_REQUIRE(A && B && D && M > 0 && N > 0 && K > 0, "zoo_test_gemm: bad argument");
    ZOO_TRY(require_sm100());
    cudaStream_t st;
    ZOO_CUDA(cudaStreamCreate(&st));
    float *dA, *dB, *dD, *ws;
    size_t na = (size_t)M * K, nb = (size_t)N * K, nd = (size_t)M * N;
    ZOO_CUDA(cudaMalloc(&dA, na * 4));
    ZOO_CUDA(cudaMalloc(&dB, nb * 4));
    ZOO_CUDA(cudaMalloc(&dD, nd * 4));
    ZOO_CUDA(cudaMalloc(&ws, nd * 4));
    ZOO_CUDA(cudaMemcpyAsync(dA, A, na * 4, cudaMemcpyHostToDevice, st));
    ZOO_CUDA(cudaMemcpyAsync(dB, B, nb * 4, cudaMemcpyHostToDevice, st));
    Operand oa{dA, 0, a_kmajor ? K : M, a_kmajor}, ob{dB, 0, b_kmajor ? K : N, b_kmajor};
    bf16 *hA = nullptr, *hB = nullptr;
    Epilogue e;
    e.out = dD; e.sm = N; e.sn = 1;
    zoo_status s;
    if (engine == 1) {
        ZOO_CUDA(cudaMalloc(&hA, na * 2));
        ZOO_CUDA(cudaMalloc(&hB, nb * 2));
        f32_to_bf16_kernel<<<cdiv(na, 256), 256, 0, st>>>(dA, hA, na);
        f32_to_bf16_kernel<<<cdiv(nb, 256), 256, 0, st>>>(dB, hB, nb);
        oa.p = hA; oa.is_bf16 = 1; ob.p = hB; ob.is_bf16 = 1;
        ZOO_REQUIRE(tc_gemm_supported(oa, ob, M, N, K), "zoo_test_gemm: shape not supported by the tcgen05 engine");
        s = tc_gemm(oa, ob, M, N, K, e, split_k, ws, nd, st);
    } else if (engine == 2 || engine == 3) {
        e.tf32_passes = engine == 2 ? 1 : 3;
        ZOO_REQUIRE(tc_gemm_supported(oa, ob, M, N, K), "zoo_test_gemm: shape not supported by the tcgen05 engine");
        s = tc_gemm(oa, ob, M, N, K, e, split_k, ws, nd, st);
    } else {
        s = simt_gemm(oa, ob, M, N, K, e, split_k, ws, nd, st);
    }
    if (s != ZOO_OK) return s;
    ZOO_CUDA(cudaMemcpyAsync(D, dD, nd * 4, cudaMemcpyDeviceToHost, st));
    ZOO_CUDA(cudaStreamSynchronize(st));
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(ws);
    if (hA) cudaFree(hA);
    if (hB) cudaFree(hB);
    cudaStreamDestroy(st);
    return ZOO_OK;
}

}  // extern "C"
