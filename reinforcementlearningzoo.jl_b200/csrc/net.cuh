// Q-network engine: Flux Chain / ImplicitQuantileNet / DuelingNetwork forward + backward on the device.
// Reference: model definitions src/experiments/atari/Dopamine_{DQN,Rainbow,IQN}_Atari.jl:26-44,
//            ImplicitQuantileNet src/algorithms/dqns/iqn.jl:17-33, DuelingNetwork src/algorithms/dqns/common.jl:44-56.
#pragma once
#include "common.cuh"

namespace zoo {

// How a Flux parameter tensor (ABI layout = Julia bytes) maps to the internal, GEMM-friendly layout.
//   dense W : ABI [in][out]            -> internal [out][in]   (K-major B operand of the forward GEMM)
//   conv  W : ABI [O][C][kh][kw]       -> internal [O][kh][kw][C] for convs that read internal NHWC activations,
//                                         unchanged for a first conv that reads the NCHW observation
//   "flat"  : a feature axis that is the flatten of a conv output is stored (h,w,c) internally (NHWC) but is
//             (c,h,w) in the reference (reshape(x,:,B) of a (W,H,C,B) array) -> that axis is permuted.
struct TensorInfo {
    int kind;      // 0 dense W, 1 conv W, 2 bias
    int n_out, n_in, ksize;
    int conv_nhwc;           // conv W stored [O][kh][kw][C]
    int perm_in, perm_out;   // axis lives in the flatten space (fc, fh, fw)
    int fc, fh, fw;
    long long offset, count; // into the flat parameter / gradient / optimiser-state buffers
};

struct LayerExec {
    int kind;  // ZOO_LAYER_CONV or ZOO_LAYER_DENSE
    int cin, cout, ksize, stride, pad, relu;
    int ih, iw, oh, ow;  // conv geometry
    int first_raw;       // reads the raw observation (u8/f32 NCHW or f32 vector)
    int scale255;        // fold x ./ 255 into the first layer's load
    int out_fp32;        // network output layer: fp32 result, fp32 incoming gradient
    int wt, bt;          // tensor ids
    long long K;         // GEMM reduction length (conv: k*k*cin, dense: in)
    long long opr;       // outputs rows per sample (conv: oh*ow, dense: 1)
    void *cols = nullptr, *out = nullptr, *dout = nullptr, *dcols = nullptr;
    // fused IQN merge (set per forward call on phi's last layer): out = relu(.) .* rowmul[row / rowmul_div]  -- iqn.jl:28-30
    const void *rowmul = nullptr; int rowmul_div = 1; long long rowmul_ld = 0;
    int rowmul_nonneg = 0;  // psi ends in a relu (the Nature-CNN trunk does): the multiplier is >= 0
};

struct Chain {
    std::vector<LayerExec> layers;
    long long rows_cap = 0;        // sample rows this chain's buffers hold
    long long in_feat = 0, out_feat = 0;
    void *out() const { return layers.empty() ? nullptr : layers.back().out; }
};

}  // namespace zoo

struct zoo_net {
    zoo_net_desc desc;
    int prec;      // ZOO_PREC_*
    int esz;       // bytes per internal activation element (4 | 2)
    int tf32_passes = 0;  // fp32 operands on tcgen05: 3 (FP32), 1 (TF32), 0 (FP32_SIMT)
    std::vector<zoo::TensorInfo> tensors;
    long long n_flat = 0;      // padded element count of the flat buffers
    long long n_params = 0;    // true parameter count
    float *params = nullptr;   // fp32 master, internal layout
    zoo::bf16 *shadow = nullptr;  // bf16 copy (BF16 precision only), same layout
    float *grads = nullptr;    // allocated when a learner/optimiser attaches
    bool grads_nccl = false;   // grads came from ncclMemAlloc (data-parallel learner: registered for zero-copy all-reduce)
    bool grads_dirty = true;   // grads may be non-zero (the optimiser kernels re-zero them; see learner.cu run_update)
    zoo::Chain a, b, c;
    int fc = 0, fh = 0, fw = 0;  // flatten space
    int n_out = 0;
    long long max_batch = 0, max_q = 0;
    // fp32 network output / its gradient
    float *out_final = nullptr, *dout_final = nullptr;
    // IQN extras.  iqn_fused: phi's last GEMM multiplies by psi in its epilogue and writes `merged` directly (phi's own output
    // is never materialised: its last layer's `out` aliases `merged`); the backward recovers what it needs from merged and psi
    // fused small-batch tail (narrow.cu: dqn_tail): the learner runs the head layer itself -- forward stops before the last layer of
    // chain a, backward starts below it (its gradients and the previous layer's bias gradient are already in place)
    bool skip_head = false;
    bool iqn_fused = false;
    void *emb = nullptr, *merged = nullptr, *dmerged = nullptr;
    float *tau_dev = nullptr;
    // dueling extras
    void *dh1 = nullptr, *dh2 = nullptr;
    float *ws = nullptr;  // split-K workspace
    size_t ws_elems = 0;
    void *obs = nullptr;  // staged observation for zoo_net_forward
    bool obs_nhwc = false;        // first conv reads an fp32 NHWC copy of the observation (fp32-storage nets, 4 channels)
    void *obs_nhwc_buf = nullptr;
    // set by the learner's staging launch when it has already written the padded NHWC copy of `prestaged_rows` observation rows
    // starting at `prestaged_src` into obs_nhwc_buf (learner.cu stage_batch_kernel): the forward then skips its own repack launch
    const void *prestaged_src = nullptr;
    long long prestaged_rows = 0;
    float *stage = nullptr; size_t stage_elems = 0;  // param up/download staging
    cudaStream_t st = nullptr;
    // stream contract: learner updates run on the learner's stream and may return before they finish (loss_out == NULL); every
    // net-level entry point (zoo_net_copy / forward / get / set params, optimiser state io) runs on `st` and first waits, on the
    // device, for ev_busy = the end of the last update that used this net (net_wait_busy)
    cudaEvent_t ev_busy = nullptr;
    bool busy_pending = false;
    bool bwd_ready = false;
    // side lane: a second stream (joined back with events, so it also becomes a parallel branch of a captured CUDA graph)
    // for work that is off the critical path -- the target net's forward, and bias-grad + wgrad while dgrad proceeds
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    float *ws_side = nullptr;  // split-K workspace of the side lane
    // late-bucket hook (see learner.cu): called on the side lane once every tensor at offset >= late_off has its gradient
    long long late_off = -1;
    zoo_status (*late_cb)(void *ctx, cudaStream_t st) = nullptr;
    void *late_ctx = nullptr;
    cudaEvent_t ev_late = nullptr, ev_late2 = nullptr, ev_comm = nullptr;
    cudaStream_t comm = nullptr;  // third lane: the late bucket's all-reduce + optimiser
    bool comm_used = false;
    size_t ws_side_elems = 0;
};

namespace zoo {
zoo_status net_alloc_backward(zoo_net *net);
// obs: device pointer, raw observation rows [rows][C][H][W] (u8|f32).  tau: device [rows][n_tau] (IQN).
// Result: net->out_final fp32 [rows (* n_tau)][n_out].
// bwd_rows: leading rows that a later net_backward will differentiate (their im2col matrices are kept for wgrad); 0 = none
zoo_status net_forward(zoo_net *net, const void *obs, int obs_dtype, int rows, const float *tau, int n_tau, cudaStream_t st, int bwd_rows);
// consumes net->dout_final (fp32, same shape as out_final, for the first `rows` samples) and accumulates into net->grads
zoo_status net_backward(zoo_net *net, int obs_dtype, const void *obs, int rows, int n_tau, cudaStream_t st);
zoo_status net_refresh_shadow(zoo_net *net, cudaStream_t st);
zoo_status net_mark_busy(zoo_net *net, cudaStream_t user);  // an update using `net` has been enqueued on `user`
zoo_status net_wait_busy(zoo_net *net);                     // net->st waits for it
// side lane plumbing: fork = side waits for everything enqueued on `main` so far; join = main waits for the side lane
zoo_status net_side_fork(zoo_net *net, cudaStream_t main);
zoo_status net_side_join(zoo_net *net, cudaStream_t main);
long long obs_elems(const zoo_net *net);
}  // namespace zoo
