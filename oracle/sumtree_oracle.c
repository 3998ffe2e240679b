/* CPU oracle for the prioritized-replay SumTree of ReinforcementLearningCore 0.7.2
 * (src/utils/sum_tree.jl upstream; NOT vendored under /root/reference -- pinned only by the
 * compat range in /root/reference/Project.toml:43).  Call sites in the reference:
 *   sample      src/algorithms/dqns/common.jl:18
 *   update      src/algorithms/dqns/common.jl:22   (t[:priority][inds] .= priorities)
 *   push        src/algorithms/dqns/common.jl:28-37
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded only by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  PARITY UNPINNED at this boundary: the
 * structure is restated from the published algorithm and pinned by
 * the RLCore docstring example (SumTree(8), push 1..16 -> leaves 9..16, root 100) and the
 * hand-computed cases of SURVEY.md A.8.
 *
 * Semantics restated (1-based, as Julia):
 *   tree :: Float32[nparents + capacity], nparents = 2^ceil(log2(capacity)) - 1, root = 1,
 *   children 2i / 2i+1, leaf slot k at nparents + k, circular `first` maps logical<->physical.
 *   setindex!: change = p - tree[leaf]; tree[leaf] = p; while node != 1: node /= 2; tree[node] += change
 *              (fp32, sequential in batch order, duplicates see earlier writes)
 *   get(v):    from the root: v <= tree[left] ? go left : (v -= tree[left]; go right), until the
 *              node has no children (left > length(tree)).
 *   sample:    independent v = u * tree[1];  stratified v = (i - 1 + u_i) / n * tree[1]
 *   The dtype of the uniform draw (Float64 default `rand(rng)` vs Float32) is not verifiable
 *   here, so both are implemented; the descent arithmetic runs in the draw's dtype.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t capacity, nparents, first, length, tree_len;
    float *tree; /* 1-based: tree[0] unused */
} st_oracle;

st_oracle *st_create(int64_t capacity) {
    st_oracle *t = (st_oracle *)calloc(1, sizeof(st_oracle));
    int64_t p = 1;
    while (p < capacity) p <<= 1;
    t->capacity = capacity;
    t->nparents = p - 1;
    t->tree_len = t->nparents + capacity;
    t->first = 1;
    t->length = 0;
    t->tree = (float *)calloc((size_t)t->tree_len + 1, sizeof(float));
    return t;
}
void st_destroy(st_oracle *t) { if (t) { free(t->tree); free(t); } }
int64_t st_length(const st_oracle *t) { return t->length; }
int64_t st_first(const st_oracle *t) { return t->first; }
int64_t st_tree_len(const st_oracle *t) { return t->tree_len; }
float st_total(const st_oracle *t) { return t->tree[1]; }
void st_export(const st_oracle *t, float *out) { memcpy(out, t->tree + 1, (size_t)t->tree_len * sizeof(float)); }

static int64_t phys(const st_oracle *t, int64_t i) {
    int64_t k = i + t->first - 1;
    return k > t->capacity ? k - t->capacity : k;
}
float st_get_leaf(const st_oracle *t, int64_t i) { return t->tree[t->nparents + phys(t, i)]; }

void st_set(st_oracle *t, int64_t i, float p) {
    int64_t n = t->nparents + phys(t, i);
    volatile float change = p - t->tree[n]; /* volatile: forbid fp contraction / excess precision */
    t->tree[n] = p;
    while (n != 1) {
        n /= 2;
        volatile float s = t->tree[n] + change;
        t->tree[n] = s;
    }
}
void st_update(st_oracle *t, const int64_t *idx, const float *p, int64_t n) {
    for (int64_t k = 0; k < n; ++k) st_set(t, idx[k], p[k]);
}
void st_push(st_oracle *t, float p) {
    if (t->length == t->capacity) t->first = (t->first == t->capacity) ? 1 : t->first + 1;
    else t->length += 1;
    st_set(t, t->length, p);
}
static int64_t logical(const st_oracle *t, int64_t n) {
    int64_t k = n - t->nparents;
    return k >= t->first ? k - t->first + 1 : k + t->capacity - t->first + 1;
}
static void get_f64(const st_oracle *t, double v, int64_t *idx, float *prio) {
    int64_t n = 1;
    for (;;) {
        int64_t l = 2 * n;
        if (l > t->tree_len) break;
        double tl = (double)t->tree[l];
        if (v <= tl) n = l; else { v -= tl; n = l + 1; }
        if (n > t->tree_len) { n = t->tree_len; break; } /* right child past the end: cannot occur for v<=total */
    }
    *prio = t->tree[n];
    *idx = logical(t, n);
}
static void get_f32(const st_oracle *t, float v, int64_t *idx, float *prio) {
    int64_t n = 1;
    for (;;) {
        int64_t l = 2 * n;
        if (l > t->tree_len) break;
        float tl = t->tree[l];
        if (v <= tl) n = l; else { volatile float d = v - tl; v = d; n = l + 1; }
        if (n > t->tree_len) { n = t->tree_len; break; }
    }
    *prio = t->tree[n];
    *idx = logical(t, n);
}
/* u: n uniforms in [0,1); dtype 0 = f32, 1 = f64; mode 0 = independent, 1 = stratified */
void st_sample(const st_oracle *t, const void *u, int dtype, int64_t n, int mode, int64_t *idx, float *prio) {
    for (int64_t i = 0; i < n; ++i) {
        if (dtype == 1) {
            double ui = ((const double *)u)[i];
            double v = mode == 1 ? ((double)i + ui) / (double)n * (double)t->tree[1] : ui * (double)t->tree[1];
            get_f64(t, v, idx + i, prio + i);
        } else {
            float ui = ((const float *)u)[i];
            volatile float a, v;
            if (mode == 1) { a = (float)i + ui; a = a / (float)n; v = a * t->tree[1]; }
            else v = ui * t->tree[1];
            get_f32(t, v, idx + i, prio + i);
        }
    }
}
