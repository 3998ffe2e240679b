// Prioritized-replay SumTree on the device (bit-exact with RLCore 0.7.2's host SumTree).
// Reference call sites: sample  src/algorithms/dqns/common.jl:18
//                       update  src/algorithms/dqns/common.jl:22   (t[:priority][inds] .= priorities)
//                       push    src/algorithms/dqns/common.jl:28-37
// The tree (fp32, 1-based heap, leaves at nparents+1..) lives in HBM and is L2-resident (8.2 MB at 1M).
//  * sample: one thread per draw; the top TOP_LEVELS levels are staged in shared memory once per CTA so
//    only the lower levels pay an L2 round trip.  Arithmetic runs in the dtype of the draw (f32 | f64)
//    with explicit round-to-nearest intrinsics so it is bit-identical to the host semantics.
//  * update: the reference applies `tree[node] += change` sequentially in batch order (fp32, order
//    dependent, duplicates see earlier writes).  Nodes are independent of each other, so per tree
//    level one CTA sorts (node, batch position) keys in shared memory and the head of every node's run
//    applies that node's changes serially in batch order: parallel across nodes and levels, bit-exact.
#include "sumtree.cuh"

namespace zoo {

static constexpr int TOP_LEVELS = 12;              // nodes 1 .. 4095 staged in smem (16 KB)
static constexpr int TOP_N = 1 << TOP_LEVELS;      // index bound (exclusive)
static constexpr int UPD_CHUNK = 4096;             // batch entries per update launch
static constexpr int KBITS = 12;                   // log2(UPD_CHUNK)

template <typename T>
__global__ void __launch_bounds__(256)
sumtree_sample_kernel(const float *__restrict__ tree, long long tree_len, long long nparents, long long capacity,
                      long long first, const T *__restrict__ u, long long n, int mode, long long *__restrict__ idx,
                      float *__restrict__ prio) {
    __shared__ float top[TOP_N];
    const long long lim = tree_len + 1 < TOP_N ? tree_len + 1 : TOP_N;
    for (int i = threadIdx.x; i < lim; i += blockDim.x) top[i] = i == 0 ? 0.f : tree[i];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T total = (T)top[1];
    T v;
    if (mode == ZOO_SAMPLE_STRATIFIED) v = Rn<T>::mul(Rn<T>::div(Rn<T>::add((T)i, u[i]), (T)n), total);
    else v = Rn<T>::mul(u[i], total);
    long long node = 1;
    for (;;) {
        const long long l = 2 * node;
        if (l > tree_len) break;
        const float tl = l < TOP_N ? top[l] : __ldg(tree + l);
        if (v <= (T)tl) node = l;
        else { v = Rn<T>::sub(v, (T)tl); node = l + 1; }
        if (node > tree_len) { node = tree_len; break; }
    }
    prio[i] = tree[node];
    const long long k = node - nparents;
    idx[i] = k >= first ? k - first + 1 : k + capacity - first + 1;
}

__global__ void sumtree_get_kernel(const float *__restrict__ tree, long long nparents, long long capacity, long long first,
                                   const long long *__restrict__ idx, long long n, float *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long k = idx[i] + first - 1;
    if (k > capacity) k -= capacity;
    out[i] = tree[nparents + k];
}

// in-smem bitonic sort of n2 (power of two) 64-bit keys
__device__ void bitonic_sort(unsigned long long *keys, int n2) {
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < n2 / 2; t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool up = (lo & size) == 0;
                unsigned long long a = keys[lo], b = keys[hi];
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
        }
    }
    __syncthreads();
}

// Phase 1 (one CTA): leaf level.  change[k] = p[k] - (value of the leaf just before entry k), leaf := last p.
__global__ void __launch_bounds__(1024)
sumtree_update_leaves_kernel(float *__restrict__ tree, long long nparents, long long capacity, long long first,
                             int physical, const long long *__restrict__ idx, const float *__restrict__ p, int n, int n2,
                             long long *__restrict__ heap_idx, float *__restrict__ change) {
    extern __shared__ unsigned long long keys[];
    for (int k = threadIdx.x; k < n2; k += blockDim.x) {
        unsigned long long key = ~0ull;
        if (k < n) {
            long long s = idx[k];
            if (!physical) { s = s + first - 1; if (s > capacity) s -= capacity; }
            long long h = nparents + s;
            heap_idx[k] = h;
            key = ((unsigned long long)h << KBITS) | (unsigned long long)k;
        }
        keys[k] = key;
    }
    bitonic_sort(keys, n2);
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const unsigned long long key = keys[j];
        const unsigned long long h = key >> KBITS;
        if (j > 0 && (keys[j - 1] >> KBITS) == h) continue;  // not the head of this leaf's run
        float prev = tree[h];
        int jj = j;
        do {
            const int k = (int)(keys[jj] & (UPD_CHUNK - 1));
            const float pk = p[k];
            change[k] = __fsub_rn(pk, prev);
            prev = pk;
            ++jj;
        } while (jj < n && (keys[jj] >> KBITS) == h);
        tree[h] = prev;
    }
}

// Phase 2 (one CTA per parent level, blockIdx.x = level 0..depth-1): tree[anc] += change[k] in batch order.
__global__ void __launch_bounds__(1024)
sumtree_update_parents_kernel(float *__restrict__ tree, int depth, const long long *__restrict__ heap_idx,
                              const float *__restrict__ change, int n, int n2) {
    extern __shared__ unsigned long long keys[];
    float *ch = (float *)(keys + n2);
    const int shift = depth - (int)blockIdx.x;  // leaf heap index >> shift = ancestor on this level
    for (int k = threadIdx.x; k < n2; k += blockDim.x) {
        unsigned long long key = ~0ull;
        if (k < n) {
            key = ((unsigned long long)(heap_idx[k] >> shift) << KBITS) | (unsigned long long)k;
            ch[k] = change[k];
        }
        keys[k] = key;
    }
    bitonic_sort(keys, n2);
    // Sorted copies (node id, change) so that the serial part below walks two dense arrays: the loop-carried dependency is
    // then only the fp32 add itself (the order of additions per node is the batch order -- that is what makes this bit-exact
    // with the reference's sequential `tree[node] += change`; near the root one thread adds thousands of terms).
    float *sch = ch + n2;
    unsigned *nd = (unsigned *)(sch + n2);
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const unsigned long long key = keys[j];
        nd[j] = (unsigned)(key >> KBITS);
        sch[j] = ch[key & (UPD_CHUNK - 1)];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const unsigned a = nd[j];
        if (j > 0 && nd[j - 1] == a) continue;
        float acc = tree[a];
        int jj = j;
        // runs of 8 with the membership tests hoisted: the adds of a full group issue back to back
        while (jj + 8 <= n && nd[jj + 7] == a) {
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = __fadd_rn(acc, sch[jj + u]);
            jj += 8;
        }
        while (jj < n && nd[jj] == a) { acc = __fadd_rn(acc, sch[jj]); ++jj; }
        tree[a] = acc;
    }
}

// Small batches (n <= SMALL_N, the Dopamine batch of 32): ONE launch of one CTA, no sort.  Thread (level, k) checks in shared
// memory whether an earlier entry shares its node on that level; the first entry of every node ("head") applies that node's
// changes serially in batch order -- the same order-preserving rule as the sorted path, hence still bit-exact -- and all nodes of
// all levels proceed in parallel: latency = one L2 round trip for the leaves + one for the ancestors.
static constexpr int SMALL_N = 256;
__global__ void __launch_bounds__(1024)
sumtree_update_small_kernel(float *__restrict__ tree, long long nparents, long long capacity, long long first, int depth, int physical,
                            const long long *__restrict__ idx, const float *__restrict__ p, int n, int dbg) {
    __shared__ long long heap[SMALL_N];
    __shared__ float pv[SMALL_N], change[SMALL_N];
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        long long s = idx[k];
        if (!physical) { s = s + first - 1; if (s > capacity) s -= capacity; }
        heap[k] = nparents + s;
        pv[k] = p[k];
    }
    __syncthreads();
    if (!(dbg & 1))
    for (int k = threadIdx.x; k < n; k += blockDim.x) {  // leaf level: change[k] = p[k] - (leaf value just before entry k)
        const long long h = heap[k];
        bool head = true;
        for (int j = 0; j < k; ++j) if (heap[j] == h) { head = false; break; }
        if (!head) continue;
        float prev = tree[h];
        for (int j = k; j < n; ++j)
            if (heap[j] == h) { change[j] = __fsub_rn(pv[j], prev); prev = pv[j]; }
        tree[h] = prev;
    }
    __syncthreads();
    if (!(dbg & 2))
    for (int t = threadIdx.x; t < depth * n; t += blockDim.x) {  // ancestors: level = t / n (shift 1 .. depth), entry k = t % n
        const int shift = t / n + 1, k = t - (shift - 1) * n;
        const long long a = heap[k] >> shift;
        bool head = true;
        for (int j = 0; j < k; ++j) if ((heap[j] >> shift) == a) { head = false; break; }
        if (!head) continue;
        float acc = tree[a];
        for (int j = k; j < n; ++j)
            if ((heap[j] >> shift) == a) acc = __fadd_rn(acc, change[j]);
        tree[a] = acc;
    }
}

// Dopamine-sized batches (n <= 32): ONE launch, one warp per tree level.  Lane k owns batch entry k.  Warp 0 settles the leaf
// level: __match_any groups the lanes that hit the same leaf; within a group an entry's change is p[k] minus the p of the group's
// previous lane (or the stored leaf value for the first), the last lane stores the leaf.  After one barrier warp l (1 .. depth)
// owns the ancestors l levels up: the lowest lane of every group of equal ancestors adds the group's changes IN LANE ORDER =
// batch order (a uniform 32-step shuffle loop, no shared-memory traffic), which is the reference's sequential
// `tree[node] += change` -- bit-exact, duplicates included -- with every level and every node in parallel.
__global__ void __launch_bounds__(1024)
sumtree_update_warp_kernel(float *__restrict__ tree, long long nparents, long long capacity, long long first, int depth, int physical,
                           const long long *__restrict__ idx, const float *__restrict__ p, int n) {
    __shared__ float change_s[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool on = lane < n;
    long long h = -1 - lane;  // inactive lanes: distinct dummies, never grouped
    float pk = 0.f;
    if (on) {
        long long s = idx[lane];
        if (!physical) { s = s + first - 1; if (s > capacity) s -= capacity; }
        h = nparents + s;
        pk = p[lane];
    }
    // every warp prefetches the node it will update (independent of the leaf phase)
    const long long a = (on && warp >= 1 && warp <= depth) ? (h >> warp) : -1 - lane;
    float acc = (on && warp >= 1 && warp <= depth) ? tree[a] : 0.f;
    if (warp == 0) {
        const float leaf = on ? tree[h] : 0.f;
        const unsigned grp = __match_any_sync(0xffffffffu, h);
        const unsigned below = grp & ((1u << lane) - 1u);
        const int prev = below ? 31 - __clz(below) : -1;
        const float pprev = __shfl_sync(0xffffffffu, pk, prev < 0 ? 0 : prev);
        change_s[lane] = on ? __fsub_rn(pk, prev < 0 ? leaf : pprev) : 0.f;
        if (on && (grp >> lane) == 1u) tree[h] = pk;  // highest lane of the group: the leaf ends up holding the last p
    }
    __syncthreads();
    if (warp >= 1 && warp <= depth) {
        const float ch = change_s[lane];
        const unsigned grp = __match_any_sync(0xffffffffu, a);
        const bool leader = on && (grp & ((1u << lane) - 1u)) == 0u;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float cj = __shfl_sync(0xffffffffu, ch, j);
            if ((grp >> j) & 1u) acc = __fadd_rn(acc, cj);  // grp = the lanes sharing MY node; lane order = batch order
        }
        if (leader) tree[a] = acc;
    }
}

// push! of ONE priority (common.jl:36): slot and value travel as kernel arguments (no staging copies, no host sync); lane l owns
// the ancestor l levels above the leaf -- all distinct nodes, so the single `+= change` per node needs no ordering.
__global__ void sumtree_push1_kernel(float *__restrict__ tree, long long leaf, float p, int depth) {
    const int l = threadIdx.x;
    const float change = __fsub_rn(p, tree[leaf]);  // every lane reads the old leaf value before lane 0 overwrites it
    __syncwarp();
    if (l == 0) tree[leaf] = p;
    else if (l <= depth) { const long long a = leaf >> l; tree[a] = __fadd_rn(tree[a], change); }
}

}  // namespace zoo

using namespace zoo;

struct zoo_sumtree {
    long long capacity = 0, nparents = 0, tree_len = 0, first = 1, length = 0;
    int depth = 0;
    float *tree = nullptr;  // device, tree_len + 1 floats (index 0 unused)
    cudaStream_t st = nullptr;
    // scratch (device)
    void *d_u = nullptr; long long *d_idx = nullptr; float *d_p = nullptr; long long cap_scratch = 0;
    long long *d_heap = nullptr; float *d_change = nullptr;
    // stream contract: the tree is also read / written on caller-chosen streams (sumtree_sample_on / sumtree_update_on: the
    // learner's stream).  ev_push marks the last push on `st` (those streams wait for it before touching the tree); ev_use marks
    // the last foreign-stream use (everything on `st` waits for it first) -- so the two streams never interleave on the tree or
    // on the d_heap / d_change scratch.
    cudaEvent_t ev_push = nullptr, ev_use = nullptr;
    bool use_pending = false, push_pending = false;
};

// everything that touches the tree on t->st first waits for the last use on a foreign stream
static zoo_status st_begin(zoo_sumtree *t) {
    if (t->use_pending) {
        ZOO_CUDA(cudaStreamWaitEvent(t->st, t->ev_use, 0));
        t->use_pending = false;
    }
    return ZOO_OK;
}

static zoo_status st_reserve(zoo_sumtree *t, long long n) {
    if (n <= t->cap_scratch) return ZOO_OK;
    long long c = std::max<long long>(n, 4096);
    if (t->d_u) { cudaFree(t->d_u); cudaFree(t->d_idx); cudaFree(t->d_p); }
    ZOO_CUDA(cudaMalloc(&t->d_u, c * sizeof(double)));
    ZOO_CUDA(cudaMalloc(&t->d_idx, c * sizeof(long long)));
    ZOO_CUDA(cudaMalloc(&t->d_p, c * sizeof(float)));
    t->cap_scratch = c;
    return ZOO_OK;
}

extern "C" {

zoo_status zoo_sumtree_create(int64_t capacity, zoo_sumtree **out) {
    ZOO_REQUIRE(out && capacity >= 1, "zoo_sumtree_create: capacity must be >= 1");
    ZOO_TRY(require_sm100());
    zoo_sumtree *t = new zoo_sumtree();
    long long p = 1; int d = 0;
    while (p < capacity) { p <<= 1; ++d; }
    t->capacity = capacity; t->nparents = p - 1; t->tree_len = t->nparents + capacity; t->depth = d;
    ZOO_CUDA(cudaStreamCreateWithFlags(&t->st, cudaStreamNonBlocking));
    ZOO_CUDA(cudaEventCreateWithFlags(&t->ev_push, cudaEventDisableTiming));
    ZOO_CUDA(cudaEventCreateWithFlags(&t->ev_use, cudaEventDisableTiming));
    ZOO_CUDA(cudaMalloc(&t->tree, (t->tree_len + 1) * sizeof(float)));
    ZOO_CUDA(cudaMemsetAsync(t->tree, 0, (t->tree_len + 1) * sizeof(float), t->st));
    ZOO_CUDA(cudaMalloc(&t->d_heap, UPD_CHUNK * sizeof(long long)));
    ZOO_CUDA(cudaMalloc(&t->d_change, UPD_CHUNK * sizeof(float)));
    ZOO_CUDA(cudaFuncSetAttribute(sumtree_update_parents_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UPD_CHUNK * 20));
    ZOO_CUDA(cudaStreamSynchronize(t->st));
    *out = t;
    return ZOO_OK;
}

zoo_status zoo_sumtree_destroy(zoo_sumtree *t) {
    if (!t) return ZOO_OK;
    cudaDeviceSynchronize();
    if (t->ev_push) cudaEventDestroy(t->ev_push);
    if (t->ev_use) cudaEventDestroy(t->ev_use);
    cudaFree(t->tree); cudaFree(t->d_u); cudaFree(t->d_idx); cudaFree(t->d_p); cudaFree(t->d_heap); cudaFree(t->d_change);
    cudaStreamDestroy(t->st);
    delete t;
    return ZOO_OK;
}

static zoo_status update_dev_impl(zoo_sumtree *t, const long long *idx_dev, const float *p_dev, long long n, int physical, cudaStream_t st = nullptr) {
    if (!st) st = t->st;
    // measured on B200 (scripts/probe_sumtree_small.py): this single-launch variant is SLOWER than the two sorted launches
    // (B = 32: 59 vs 10 us), so it is opt-in (ZOO_SUMTREE_SMALL=1) until that is understood
    static const bool warp_on = [] { const char *e = getenv("ZOO_SUMTREE_WARP"); return !(e && e[0] == '0'); }();
    if (warp_on && n > 0 && n <= 32 && t->depth <= 31) {
        sumtree_update_warp_kernel<<<1, 32 * (t->depth + 1), 0, st>>>(t->tree, t->nparents, t->capacity, t->first, t->depth, physical, idx_dev, p_dev, (int)n);
        ZOO_LAUNCH_CHECK();
        return ZOO_OK;
    }
    static const bool small_on = [] { const char *e = getenv("ZOO_SUMTREE_SMALL"); return e && e[0] == '1'; }();
    if (small_on && n > 0 && n <= SMALL_N) {
        const int threads = (int)std::min<long long>(1024, std::max<long long>(32, ((long long)(t->depth + 1) * n + 31) / 32 * 32));
        static const int dbg = [] { const char *e = getenv("ZOO_ST_DBG"); return e ? atoi(e) : 0; }();
        sumtree_update_small_kernel<<<1, (dbg & 4) ? 1024 : threads, 0, st>>>(t->tree, t->nparents, t->capacity, t->first, t->depth, physical, idx_dev, p_dev, (int)n, dbg);
        ZOO_LAUNCH_CHECK();
        return ZOO_OK;
    }
    for (long long off = 0; off < n; off += UPD_CHUNK) {
        int m = (int)std::min<long long>(UPD_CHUNK, n - off);
        int n2 = 32;
        while (n2 < m) n2 <<= 1;
        int threads = std::min(1024, std::max(32, n2 / 2));
        sumtree_update_leaves_kernel<<<1, threads, n2 * 8, st>>>(t->tree, t->nparents, t->capacity, t->first, physical,
                                                                    idx_dev + off, p_dev + off, m, n2, t->d_heap, t->d_change);
        ZOO_LAUNCH_CHECK();
        if (t->depth > 0) {
            sumtree_update_parents_kernel<<<t->depth, threads, n2 * 20, st>>>(t->tree, t->depth, t->d_heap, t->d_change, m, n2);
            ZOO_LAUNCH_CHECK();
        }
    }
    return ZOO_OK;
}

zoo_status zoo_sumtree_update_dev(zoo_sumtree *t, const int64_t *idx_dev, const float *p_dev, int64_t n) {
    ZOO_REQUIRE(t && idx_dev && p_dev && n >= 0, "zoo_sumtree_update_dev: bad argument");
    ZOO_TRY(st_begin(t));
    return update_dev_impl(t, (const long long *)idx_dev, p_dev, n, 0);
}

zoo_status zoo_sumtree_update(zoo_sumtree *t, const int64_t *idx, const float *p, int64_t n) {
    ZOO_REQUIRE(t && idx && p && n >= 0, "zoo_sumtree_update: bad argument");
    if (n == 0) return ZOO_OK;
    for (int64_t i = 0; i < n; ++i)
        ZOO_REQUIRE(idx[i] >= 1 && idx[i] <= t->capacity, "zoo_sumtree_update: index %lld out of 1..%lld", (long long)idx[i], t->capacity);
    ZOO_TRY(st_begin(t));
    ZOO_TRY(st_reserve(t, n));
    ZOO_CUDA(cudaMemcpyAsync(t->d_idx, idx, n * sizeof(long long), cudaMemcpyHostToDevice, t->st));
    ZOO_CUDA(cudaMemcpyAsync(t->d_p, p, n * sizeof(float), cudaMemcpyHostToDevice, t->st));
    ZOO_TRY(update_dev_impl(t, t->d_idx, t->d_p, n, 0));
    ZOO_CUDA(cudaStreamSynchronize(t->st));
    return ZOO_OK;
}

zoo_status zoo_sumtree_push_many(zoo_sumtree *t, const float *p, int64_t n) {
    ZOO_REQUIRE(t && p && n >= 0, "zoo_sumtree_push_many: bad argument");
    // push! = (advance first | grow length) then setindex!(p, length): the physical slot written advances by
    // one per push, so a run of pushes is one batched update on consecutive physical slots.
    int64_t done = 0;
    std::vector<long long> slots;
    ZOO_TRY(st_begin(t));
    while (done < n) {
        int64_t m = std::min<int64_t>(std::min<int64_t>(UPD_CHUNK, t->capacity), n - done);
        slots.resize(m);
        for (int64_t i = 0; i < m; ++i) {
            if (t->length == t->capacity) t->first = (t->first == t->capacity) ? 1 : t->first + 1;
            else t->length += 1;
            long long k = t->length + t->first - 1;
            if (k > t->capacity) k -= t->capacity;
            slots[i] = k;
        }
        ZOO_TRY(st_reserve(t, m));
        ZOO_CUDA(cudaMemcpyAsync(t->d_idx, slots.data(), m * sizeof(long long), cudaMemcpyHostToDevice, t->st));
        ZOO_CUDA(cudaMemcpyAsync(t->d_p, p + done, m * sizeof(float), cudaMemcpyHostToDevice, t->st));
        ZOO_TRY(update_dev_impl(t, t->d_idx, t->d_p, m, 1));
        ZOO_CUDA(cudaStreamSynchronize(t->st));
        done += m;
    }
    return ZOO_OK;
}

// one push: asynchronous (one 32-thread launch on the tree's stream, no copies, no host sync); consumers on other streams wait
// for ev_push, host-facing reads synchronise the tree's stream as before
zoo_status zoo_sumtree_push(zoo_sumtree *t, float p) {
    ZOO_REQUIRE(t, "zoo_sumtree_push: null tree");
    ZOO_REQUIRE(t->depth < 32, "zoo_sumtree_push: tree too deep");
    ZOO_TRY(st_begin(t));
    if (t->length == t->capacity) t->first = (t->first == t->capacity) ? 1 : t->first + 1;
    else t->length += 1;
    long long k = t->length + t->first - 1;
    if (k > t->capacity) k -= t->capacity;
    sumtree_push1_kernel<<<1, 32, 0, t->st>>>(t->tree, t->nparents + k, p, t->depth);
    ZOO_LAUNCH_CHECK();
    ZOO_CUDA(cudaEventRecord(t->ev_push, t->st));
    t->push_pending = true;
    return ZOO_OK;
}

static zoo_status sample_dev_impl(zoo_sumtree *t, const void *u_dev, int u_dtype, long long n, int mode, long long *idx_dev,
                                  float *prio_dev, cudaStream_t st = nullptr) {
    if (!st) st = t->st;
    int blocks = cdiv(n, 256);
    if (u_dtype == ZOO_DRAW_F64)
        sumtree_sample_kernel<double><<<blocks, 256, 0, st>>>(t->tree, t->tree_len, t->nparents, t->capacity, t->first,
                                                                 (const double *)u_dev, n, mode, idx_dev, prio_dev);
    else
        sumtree_sample_kernel<float><<<blocks, 256, 0, st>>>(t->tree, t->tree_len, t->nparents, t->capacity, t->first,
                                                                (const float *)u_dev, n, mode, idx_dev, prio_dev);
    ZOO_LAUNCH_CHECK();
    return ZOO_OK;
}

zoo_status zoo_sumtree_sample_dev(zoo_sumtree *t, const void *u_dev, int32_t u_dtype, int64_t n, int32_t mode,
                                  int64_t *idx_dev, float *prio_dev) {
    ZOO_REQUIRE(t && u_dev && idx_dev && prio_dev && n >= 0, "zoo_sumtree_sample_dev: bad argument");
    ZOO_REQUIRE(u_dtype == ZOO_DRAW_F32 || u_dtype == ZOO_DRAW_F64, "zoo_sumtree_sample_dev: bad draw dtype");
    if (n == 0) return ZOO_OK;
    ZOO_TRY(st_begin(t));
    return sample_dev_impl(t, u_dev, u_dtype, n, mode, (long long *)idx_dev, prio_dev);
}

zoo_status zoo_sumtree_sample(zoo_sumtree *t, const void *u, int32_t u_dtype, int64_t n, int32_t mode, int64_t *idx_out,
                              float *prio_out) {
    ZOO_REQUIRE(t && u && idx_out && prio_out && n >= 0, "zoo_sumtree_sample: bad argument");
    ZOO_REQUIRE(u_dtype == ZOO_DRAW_F32 || u_dtype == ZOO_DRAW_F64, "zoo_sumtree_sample: bad draw dtype");
    ZOO_REQUIRE(mode == ZOO_SAMPLE_INDEPENDENT || mode == ZOO_SAMPLE_STRATIFIED, "zoo_sumtree_sample: bad mode");
    if (n == 0) return ZOO_OK;
    ZOO_TRY(st_begin(t));
    ZOO_TRY(st_reserve(t, n));
    size_t esz = u_dtype == ZOO_DRAW_F64 ? 8 : 4;
    ZOO_CUDA(cudaMemcpyAsync(t->d_u, u, n * esz, cudaMemcpyHostToDevice, t->st));
    ZOO_TRY(sample_dev_impl(t, t->d_u, u_dtype, n, mode, t->d_idx, t->d_p));
    ZOO_CUDA(cudaMemcpyAsync(idx_out, t->d_idx, n * sizeof(long long), cudaMemcpyDeviceToHost, t->st));
    ZOO_CUDA(cudaMemcpyAsync(prio_out, t->d_p, n * sizeof(float), cudaMemcpyDeviceToHost, t->st));
    ZOO_CUDA(cudaStreamSynchronize(t->st));
    return ZOO_OK;
}

zoo_status zoo_sumtree_total(zoo_sumtree *t, float *total) {
    ZOO_REQUIRE(t && total, "zoo_sumtree_total: bad argument");
    ZOO_TRY(st_begin(t));
    ZOO_CUDA(cudaMemcpyAsync(total, t->tree + 1, sizeof(float), cudaMemcpyDeviceToHost, t->st));
    ZOO_CUDA(cudaStreamSynchronize(t->st));
    return ZOO_OK;
}

zoo_status zoo_sumtree_get(zoo_sumtree *t, const int64_t *idx, int64_t n, float *prio_out) {
    ZOO_REQUIRE(t && idx && prio_out && n >= 0, "zoo_sumtree_get: bad argument");
    if (n == 0) return ZOO_OK;
    ZOO_TRY(st_begin(t));
    ZOO_TRY(st_reserve(t, n));
    ZOO_CUDA(cudaMemcpyAsync(t->d_idx, idx, n * sizeof(long long), cudaMemcpyHostToDevice, t->st));
    sumtree_get_kernel<<<cdiv(n, 256), 256, 0, t->st>>>(t->tree, t->nparents, t->capacity, t->first, t->d_idx, n, t->d_p);
    ZOO_LAUNCH_CHECK();
    ZOO_CUDA(cudaMemcpyAsync(prio_out, t->d_p, n * sizeof(float), cudaMemcpyDeviceToHost, t->st));
    ZOO_CUDA(cudaStreamSynchronize(t->st));
    return ZOO_OK;
}

zoo_status zoo_sumtree_length(const zoo_sumtree *t, int64_t *length, int64_t *first) {
    ZOO_REQUIRE(t, "zoo_sumtree_length: null tree");
    if (length) *length = t->length;
    if (first) *first = t->first;
    return ZOO_OK;
}

zoo_status zoo_sumtree_tree_len(const zoo_sumtree *t, int64_t *n) {
    ZOO_REQUIRE(t && n, "zoo_sumtree_tree_len: bad argument");
    *n = t->tree_len;
    return ZOO_OK;
}

zoo_status zoo_sumtree_export(zoo_sumtree *t, float *host_tree, int64_t n) {
    ZOO_REQUIRE(t && host_tree && n == t->tree_len, "zoo_sumtree_export: expected %lld floats", t ? t->tree_len : 0);
    ZOO_TRY(st_begin(t));
    ZOO_CUDA(cudaMemcpyAsync(host_tree, t->tree + 1, n * sizeof(float), cudaMemcpyDeviceToHost, t->st));
    ZOO_CUDA(cudaStreamSynchronize(t->st));
    return ZOO_OK;
}

zoo_status zoo_sumtree_import(zoo_sumtree *t, const float *host_tree, int64_t n, int64_t length, int64_t first) {
    ZOO_REQUIRE(t && host_tree && n == t->tree_len, "zoo_sumtree_import: expected %lld floats", t ? t->tree_len : 0);
    ZOO_REQUIRE(length >= 0 && length <= t->capacity && first >= 1 && first <= t->capacity, "zoo_sumtree_import: bad length/first");
    ZOO_TRY(st_begin(t));
    ZOO_CUDA(cudaMemcpyAsync(t->tree + 1, host_tree, n * sizeof(float), cudaMemcpyHostToDevice, t->st));
    ZOO_CUDA(cudaStreamSynchronize(t->st));
    t->length = length; t->first = first;
    return ZOO_OK;
}

zoo_status zoo_sumtree_stream(zoo_sumtree *t, void **stream) {
    ZOO_REQUIRE(t && stream, "zoo_sumtree_stream: bad argument");
    *stream = (void *)t->st;
    return ZOO_OK;
}

zoo_status zoo_sumtree_sync(zoo_sumtree *t) {
    ZOO_REQUIRE(t, "zoo_sumtree_sync: null tree");
    ZOO_CUDA(cudaStreamSynchronize(t->st));
    return ZOO_OK;
}

}  // extern "C"

// internal: the same device-side sample / update enqueued on a caller-chosen stream (replay.cu)
namespace zoo {
// a foreign stream first waits for the pushes enqueued on the tree's own stream, and leaves a marker the next push waits for
static zoo_status foreign_begin(zoo_sumtree *t, cudaStream_t st) {
    if (st != t->st && t->push_pending) ZOO_CUDA(cudaStreamWaitEvent(st, t->ev_push, 0));
    return ZOO_OK;
}
static zoo_status foreign_end(zoo_sumtree *t, cudaStream_t st) {
    if (st != t->st) {
        ZOO_CUDA(cudaEventRecord(t->ev_use, st));
        t->use_pending = true;
    }
    return ZOO_OK;
}
zoo_status sumtree_sample_on(zoo_sumtree *t, cudaStream_t st, const void *u_dev, int u_dtype, long long n, int mode, long long *idx_dev, float *prio_dev) {
    ZOO_TRY(foreign_begin(t, st));
    ZOO_TRY(sample_dev_impl(t, u_dev, u_dtype, n, mode, idx_dev, prio_dev, st));
    return foreign_end(t, st);
}
zoo_status sumtree_update_on(zoo_sumtree *t, cudaStream_t st, const long long *idx_dev, const float *p_dev, long long n) {
    ZOO_TRY(foreign_begin(t, st));
    ZOO_TRY(update_dev_impl(t, idx_dev, p_dev, n, 0, st));
    return foreign_end(t, st);
}
void sumtree_state(const zoo_sumtree *t, long long *length, long long *first) { *length = t->length; *first = t->first; }
void sumtree_view(const zoo_sumtree *t, SumTreeView *v) {
    v->tree = t->tree; v->tree_len = t->tree_len; v->nparents = t->nparents; v->capacity = t->capacity; v->first = t->first;
}
}  // namespace zoo
