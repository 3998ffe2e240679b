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

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("ZOO_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

int g_sm_budget = 148;
int g_tc_stages_force = [] { const char *e = getenv("ZOO_TC_STAGES"); return e ? atoi(e) : 0; }();
thread_local bool g_timeline_on = false;
struct TlRec { cudaEvent_t e0, e1; cudaStream_t st; std::string label; };
static thread_local std::vector<TlRec> g_tl;

thread_local bool g_prof_on = false;
static thread_local cudaStream_t g_prof_stream = nullptr;
static thread_local std::vector<cudaEvent_t> g_prof_events;
static thread_local std::vector<std::string> g_prof_names;
static thread_local std::vector<float> g_prof_ms;
static thread_local char g_prof_lbl[96] = "";

void prof_label(const char *fmt, ...) {
    if (!g_prof_on && !g_timeline_on) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_prof_lbl, sizeof(g_prof_lbl), fmt, ap);
    va_end(ap);
}

void timeline_mark(cudaStream_t st, int end) {
    if (!end) {
        TlRec r;
        cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
        r.st = st; r.label = g_prof_lbl;
        cudaEventRecord(r.e0, st);
        g_tl.push_back(r);
    } else if (!g_tl.empty()) {
        cudaEventRecord(g_tl.back().e1, st);
    }
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
extern unsigned long long *g_tc_dbg;
extern unsigned long long *g_tc_trace;
extern int g_tc_trace_cap, g_tc_trace_n;
extern std::vector<std::string> g_tc_trace_names;

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

// timeline mode: zoo_prof_timeline(1) ... run ONE eager update ... zoo_prof_timeline(0) returns the number of records
zoo_status zoo_prof_timeline(int32_t on, int32_t *n_records) {
    if (on) {
        for (auto &r : g_tl) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
        g_tl.clear();
        g_prof_lbl[0] = 0;
        g_timeline_on = true;
    } else {
        g_timeline_on = false;
        ZOO_CUDA(cudaDeviceSynchronize());
        if (n_records) *n_records = (int32_t)g_tl.size();
    }
    return ZOO_OK;
}
zoo_status zoo_prof_timeline_get(int32_t i, char *name, int32_t cap, uint64_t *stream, float *start_us, float *end_us) {
    ZOO_REQUIRE(name && stream && start_us && end_us && i >= 0 && i < (int)g_tl.size() && cap > 0, "zoo_prof_timeline_get: bad index");
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, g_tl[0].e0, g_tl[i].e0);
    cudaEventElapsedTime(&b, g_tl[0].e0, g_tl[i].e1);
    snprintf(name, cap, "%s", g_tl[i].label.c_str());
    *stream = (uint64_t)(uintptr_t)g_tl[i].st; *start_us = a * 1e3f; *end_us = b * 1e3f;
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
    ZOO_CUDA(cudaMalloc(&ws, (nd + 4096) * 4));
    ZOO_CUDA(cudaMemsetAsync(ws, 0, (nd + 4096) * 4, st));
    ZOO_CUDA(cudaMemcpyAsync(dA, A, na * 4, cudaMemcpyHostToDevice, st));
    ZOO_CUDA(cudaMemcpyAsync(dB, B, nb * 4, cudaMemcpyHostToDevice, st));
    Operand oa{dA, 0, a_kmajor ? K : M, a_kmajor}, ob{dB, 0, b_kmajor ? K : N, b_kmajor};
    bf16 *hA = nullptr, *hB = nullptr;
    Epilogue e;
    e.out = dD; e.sm = N; e.sn = 1;
    zoo_status s;
    if (engine == 1 || engine == 4) {
        ZOO_CUDA(cudaMalloc(&hA, na * 2));
        ZOO_CUDA(cudaMalloc(&hB, nb * 2));
        f32_to_bf16_kernel<<<cdiv(na, 256), 256, 0, st>>>(dA, hA, na);
        f32_to_bf16_kernel<<<cdiv(nb, 256), 256, 0, st>>>(dB, hB, nb);
        oa.p = hA; oa.is_bf16 = 1; ob.p = hB; ob.is_bf16 = 1;
        ZOO_REQUIRE(tc_gemm_supported(oa, ob, M, N, K), "zoo_test_gemm: shape not supported by the tcgen05 engine");
        if (engine == 4) {  // the persistent engine (large-batch bf16 GEMMs)
            ZOO_REQUIRE(tc_gemm_persist_ok(oa, ob, M, N, K, e), "zoo_test_gemm: shape not taken by the persistent engine");
            s = tc_gemm_persist(oa, ob, M, N, K, e, st);
        } else {
            s = tc_gemm(oa, ob, M, N, K, e, split_k, ws, nd + 4096, st);
        }
    } else if (engine == 2 || engine == 3) {
        e.tf32_passes = engine == 2 ? 1 : 3;
        ZOO_REQUIRE(tc_gemm_supported(oa, ob, M, N, K), "zoo_test_gemm: shape not supported by the tcgen05 engine");
        s = tc_gemm(oa, ob, M, N, K, e, split_k, ws, nd + 4096, st);
    } else {
        s = simt_gemm(oa, ob, M, N, K, e, split_k, ws, nd + 4096, st);
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

// Phase timestamps (ns, %globaltimer) of CTA (0,0,0) of one tcgen05 GEMM launch: 0 entry, 1 setup done, 2 first TMA issued,
// 3 first stage landed, 4 last MMA committed, 5 accumulator ready (epilogue), 6 epilogue done, 7 CTA joined, 8 TMEM freed.
zoo_status zoo_test_gemm_trace(int32_t engine, int32_t M, int32_t N, int32_t K, int32_t a_kmajor, int32_t b_kmajor, int32_t split_k,
                               uint64_t *stamps9) {
    ZOO_REQUIRE(stamps9, "zoo_test_gemm_trace: null argument");
    unsigned long long *d;
    ZOO_CUDA(cudaMalloc(&d, 16 * 8));
    ZOO_CUDA(cudaMemset(d, 0, 16 * 8));
    float us;
    zoo_status s = zoo_test_gemm_bench(engine, M, N, K, a_kmajor, b_kmajor, split_k, 0, 3, &us);  // warm
    g_tc_dbg = d;
    if (s == ZOO_OK) s = zoo_test_gemm_bench(engine, M, N, K, a_kmajor, b_kmajor, split_k, 0, 1, &us);
    g_tc_dbg = nullptr;
    ZOO_CUDA(cudaMemcpy(stamps9, d, 16 * 8, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return s;
}

// Per-launch phase trace of the tcgen05 engine (see gemm_tc.cu): begin(cap) arms it, every engine launch from then on -- also
// the ones captured into a CUDA graph -- stamps CTA (0,0,0)'s phases into its own slot; get(i) reads slot i back.
zoo_status zoo_test_trace_begin(int32_t cap) {
    ZOO_REQUIRE(cap > 0 && cap <= 4096, "zoo_test_trace_begin: bad capacity");
    if (g_tc_trace) cudaFree(g_tc_trace);
    g_tc_trace = nullptr;
    unsigned long long *d;
    ZOO_CUDA(cudaMalloc(&d, (size_t)cap * 16 * 8));
    ZOO_CUDA(cudaMemset(d, 0, (size_t)cap * 16 * 8));
    g_tc_trace_names.clear();
    g_tc_trace_cap = cap; g_tc_trace_n = 0; g_tc_trace = d;
    return ZOO_OK;
}
zoo_status zoo_test_trace_stop(int32_t *n) {  // disarms; the slots stay readable (and graph replays keep stamping them)
    ZOO_REQUIRE(n, "zoo_test_trace_stop: null argument");
    *n = g_tc_trace_n;
    g_tc_trace_cap = g_tc_trace_n;
    return ZOO_OK;
}
zoo_status zoo_test_trace_get(int32_t i, char *name, int32_t name_cap, uint64_t *stamps16) {
    ZOO_REQUIRE(g_tc_trace && name && stamps16 && i >= 0 && i < g_tc_trace_n && name_cap > 0, "zoo_test_trace_get: bad index");
    snprintf(name, name_cap, "%s", g_tc_trace_names[i].c_str());
    ZOO_CUDA(cudaMemcpy(stamps16, g_tc_trace + 16 * i, 16 * 8, cudaMemcpyDeviceToHost));
    return ZOO_OK;
}

// Forces the tcgen05 GEMM's smem ring depth (0 = per-launch choice); a tuning / test hook.
zoo_status zoo_test_set_tc_stages(int32_t stages) {
    g_tc_stages_force = stages;
    return ZOO_OK;
}

// Device-time benchmark of one GEMM shape (test hook): random-free (zeros) operands resident on the device, `iters`
// back-to-back launches between two events; returns the mean microseconds per GEMM (all its launches).
zoo_status zoo_test_gemm_bench(int32_t engine, int32_t M, int32_t N, int32_t K, int32_t a_kmajor, int32_t b_kmajor, int32_t split_k,
                               int32_t mode, int32_t iters, float *us_out) {
    ZOO_REQUIRE(us_out && M > 0 && N > 0 && K > 0 && iters > 0, "zoo_test_gemm_bench: bad argument");
    ZOO_TRY(require_sm100());
    cudaStream_t st;
    ZOO_CUDA(cudaStreamCreate(&st));
    const int eb = engine == 1 ? 2 : 4;
    void *dA, *dB; float *dD, *ws;
    size_t na = (size_t)M * K, nb = (size_t)N * K, nd = (size_t)M * N;
    // mode & 8: cold operands -- every launch reads its own copy of A and B, the copies together larger than L2 (126 MB)
    const bool cold = (mode & 8) != 0;
    mode &= 7;
    int copies = 1;
    if (cold) copies = (int)std::max<size_t>(1, std::min<size_t>((size_t)iters, (320u << 20) / ((na + nb) * eb) + 1));
    ZOO_CUDA(cudaMalloc(&dA, na * eb * copies));
    ZOO_CUDA(cudaMalloc(&dB, nb * eb * copies));
    ZOO_CUDA(cudaMalloc(&dD, nd * 4));
    ZOO_CUDA(cudaMalloc(&ws, (nd + 4096) * 4));
    ZOO_CUDA(cudaMemsetAsync(dA, 0, na * eb * copies, st));
    ZOO_CUDA(cudaMemsetAsync(dB, 0, nb * eb * copies, st));
    ZOO_CUDA(cudaMemsetAsync(dD, 0, nd * 4, st));
    ZOO_CUDA(cudaMemsetAsync(ws, 0, (nd + 4096) * 4, st));
    Operand oa{dA, eb == 2, a_kmajor ? K : M, a_kmajor}, ob{dB, eb == 2, b_kmajor ? K : N, b_kmajor};
    Epilogue e;
    e.out = dD; e.sm = N; e.sn = 1;
    if (mode == 1) e.accumulate = 1;            // wgrad-style atomic accumulate
    if (mode == 2) { e.relu = 1; e.bias = dD; e.bias_mode = 1; }  // forward-style
    e.tf32_passes = engine == 2 ? 1 : (engine == 3 ? 3 : 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int turn = 0;
    auto one = [&]() {
        Operand a = oa, b = ob;
        a.p = (char *)dA + (size_t)(turn % copies) * na * eb;
        b.p = (char *)dB + (size_t)(turn % copies) * nb * eb;
        ++turn;
        return engine == 0 ? simt_gemm(a, b, M, N, K, e, split_k, ws, nd + 4096, st) : tc_gemm(a, b, M, N, K, e, split_k, ws, nd + 4096, st);
    };
    for (int it = 0; it < 3; ++it) ZOO_TRY(one());
    turn = 0;
    // the timed launches are replayed from a CUDA graph so the number is device time, not host launch rate
    cudaGraph_t graph;
    cudaGraphExec_t gexec;
    ZOO_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int it = 0; it < iters; ++it) {
        zoo_status s = one();
        if (s != ZOO_OK) { cudaStreamEndCapture(st, &graph); return s; }
    }
    ZOO_CUDA(cudaStreamEndCapture(st, &graph));
    ZOO_CUDA(cudaGraphInstantiate(&gexec, graph, 0));
    ZOO_CUDA(cudaGraphLaunch(gexec, st));
    cudaEventRecord(e0, st);
    ZOO_CUDA(cudaGraphLaunch(gexec, st));
    cudaEventRecord(e1, st);
    ZOO_CUDA(cudaStreamSynchronize(st));
    cudaGraphExecDestroy(gexec);
    cudaGraphDestroy(graph);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *us_out = ms * 1e3f / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(ws);
    cudaStreamDestroy(st);
    return ZOO_OK;
}

}  // extern "C"
