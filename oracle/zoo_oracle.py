"""CPU oracle for the value-based learner update of ReinforcementLearningZoo.jl v0.3.7.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the CPU arm.  The product path
(``reinforcementlearningzoo.jl_b200``) never imports this module and fails loudly when
``libzoo_sm100.so`` is missing.

PARITY UNPINNED.  The reference is pure Julia and cannot be executed in this image
(no ``julia``), its own tests (``test/runtests.jl:34-48``) only smoke-run the experiments
and pin no loss / gradient / priority, and half of the arithmetic lives in un-vendored
dependencies (Flux 0.11, Zygote 0.5/0.6, ReinforcementLearningCore 0.7.2 -- compat ranges
from ``Project.toml:33-50``, no Manifest).  This file therefore *restates* the algorithm:
every function cites the reference lines it follows; the Flux/Zygote/RLCore formulas are
restated from their published definitions.  The restatement is
defended by (1) an independent PyTorch-autograd twin (``oracle/torch_twin.py``) that must
agree to <=1e-5, (2) the hand-computable micro cases of SURVEY.md A.8, (3) fp64
finite-difference gradient checks (``tests/test_oracle_*.py``).

Conventions (all row-major numpy; identical bytes to the Julia column-major arrays):
  frames / conv activations  [B, C, H, W]      == Julia (W, H, C, B)
  CrossCor weight            [Cout, Cin, kh, kw] == Julia (kw, kh, Cin, Cout)
  Dense weight               [out, in] logical; the C ABI ships it as Julia's column-major
                             (out, in), i.e. row-major [in][out] -- see ``to_abi_layout``.
  actions                    0-based int (Julia glue subtracts 1)
All learner arithmetic is float32 unless ``dtype=np.float64`` is requested (used only to
calibrate tolerances and for finite differences).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# --------------------------------------------------------------------------------------
# network description
# --------------------------------------------------------------------------------------
# A "chain" is a list of layer specs (Flux ``Chain``):
#   ("scale255",)                                 x ./ 255           Dopamine_DQN_Atari.jl:28
#   ("conv", cin, cout, k, stride, pad, relu)     CrossCor           Dopamine_DQN_Atari.jl:29-31
#   ("flatten",)                                  reshape(x,:,B)     Dopamine_DQN_Atari.jl:32
#   ("dense", nin, nout, relu)                    Dense              Dopamine_DQN_Atari.jl:33-34
# A net is a dict {"kind": "chain"|"iqn"|"dueling", ...}.


def nature_cnn(n_out: int, in_ch: int = 4, head: bool = True):
    """Dopamine Nature-CNN with the reference's padding (q9): 84 -> 21 -> 11 -> 11, fc1 fan-in 7744.
    src/experiments/atari/Dopamine_DQN_Atari.jl:26-35."""
    layers = [
        ("scale255",),
        ("conv", in_ch, 32, 8, 4, 2, True),
        ("conv", 32, 64, 4, 2, 2, True),
        ("conv", 64, 64, 3, 1, 1, True),
        ("flatten",),
    ]
    if head:
        layers += [("dense", 11 * 11 * 64, 512, True), ("dense", 512, n_out, False)]
    return layers


def mlp(sizes, relu_last=False):
    """Chain(Dense(ns,128,relu), Dense(128,128,relu), Dense(128,na)) --
    src/experiments/rl_envs/JuliaRL_BasicDQN_CartPole.jl:23-27."""
    layers = []
    for i in range(len(sizes) - 1):
        last = i == len(sizes) - 2
        layers.append(("dense", sizes[i], sizes[i + 1], (not last) or relu_last))
    return layers


def chain_net(layers):
    return {"kind": "chain", "layers": list(layers)}


def iqn_net(psi, phi, header):
    """ImplicitQuantileNet(ψ, ϕ, header) -- src/algorithms/dqns/iqn.jl:17-23."""
    return {"kind": "iqn", "psi": list(psi), "phi": list(phi), "header": list(header)}


def dueling_net(base, val, adv):
    """DuelingNetwork(base, val, adv) -- src/algorithms/dqns/common.jl:44-56."""
    return {"kind": "dueling", "base": list(base), "val": list(val), "adv": list(adv)}


def atari_iqn_net(n_actions, n_em=64, in_ch=4):
    """src/experiments/atari/Dopamine_IQN_Atari.jl:30-44."""
    return iqn_net(
        nature_cnn(0, in_ch, head=False),
        [("dense", n_em, 11 * 11 * 64, True)],
        [("dense", 11 * 11 * 64, 512, True), ("dense", 512, n_actions, False)],
    )


def cartpole_iqn_net(ns=4, na=2, n_em=16, n_hidden=64):
    """src/experiments/rl_envs/JuliaRL_IQN_CartPole.jl:25-34."""
    return iqn_net(
        [("dense", ns, n_hidden, True)],
        [("dense", n_em, n_hidden, True)],
        [("dense", n_hidden, na, False)],
    )


def chain_param_shapes(layers):
    shapes = []
    for l in layers:
        if l[0] == "conv":
            _, cin, cout, k, _, _, _ = l
            shapes += [(cout, cin, k, k), (cout,)]
        elif l[0] == "dense":
            _, nin, nout, _ = l
            shapes += [(nout, nin), (nout,)]
    return shapes


def net_chains(net):
    """Sub-chains in Flux ``params`` order (``Flux.@functor`` field order)."""
    if net["kind"] == "chain":
        return [net["layers"]]
    if net["kind"] == "iqn":
        return [net["psi"], net["phi"], net["header"]]
    if net["kind"] == "dueling":
        return [net["base"], net["val"], net["adv"]]
    raise ValueError(net["kind"])


def param_shapes(net):
    out = []
    for c in net_chains(net):
        out += chain_param_shapes(c)
    return out


def glorot_uniform(rng, shape):
    """Flux 0.11 ``glorot_uniform``: (rand(Float32, dims) .- 0.5f0) .* sqrt(24f0/(fan_in+fan_out))
    (SURVEY.md A.3).  Dense (out,in): fans (in,out); conv: fan = k*k*c."""
    if len(shape) == 2:
        fan_in, fan_out = shape[1], shape[0]
    else:
        rf = int(np.prod(shape[2:]))
        fan_in, fan_out = shape[1] * rf, shape[0] * rf
    u = rng.random(shape, dtype=np.float32)
    return ((u - F32(0.5)) * np.sqrt(F32(24.0) / F32(fan_in + fan_out))).astype(F32)


def init_params(net, seed, bias_scale=0.0):
    """Synthetic weights for tests/bench: glorot_uniform weights; biases 0 as in
    Flux, or small random when ``bias_scale`` > 0 so that bias paths are exercised."""
    rng = np.random.default_rng(seed)
    ps = []
    for shp in param_shapes(net):
        if len(shp) == 1:
            b = np.zeros(shp, F32)
            if bias_scale:
                b = ((rng.random(shp, dtype=np.float32) - F32(0.5)) * F32(bias_scale)).astype(F32)
            ps.append(b)
        else:
            ps.append(glorot_uniform(rng, shp))
    return ps


def to_abi_layout(p):
    """logical numpy param -> bytes the C ABI expects (= the Julia array's memory).
    Dense (out,in) column-major == row-major [in][out]; conv OIHW and biases unchanged."""
    if p.ndim == 2:
        return np.ascontiguousarray(p.T)
    return np.ascontiguousarray(p)


def from_abi_layout(buf, shape):
    if len(shape) == 2:
        return np.ascontiguousarray(np.asarray(buf).reshape(shape[1], shape[0]).T)
    return np.asarray(buf).reshape(shape).copy()


# --------------------------------------------------------------------------------------
# layers: forward / backward (Flux 0.11 + NNlib + Zygote adjoints)
# --------------------------------------------------------------------------------------
def _im2col(x, k, s, p):
    B, C, H, W = x.shape
    xp = np.pad(x, ((0, 0), (0, 0), (p, p), (p, p)))
    OH = (H + 2 * p - k) // s + 1
    OW = (W + 2 * p - k) // s + 1
    win = np.lib.stride_tricks.sliding_window_view(xp, (k, k), axis=(2, 3))  # B,C,H',W',k,k
    win = win[:, :, ::s, ::s][:, :, :OH, :OW]
    cols = win.transpose(0, 2, 3, 1, 4, 5).reshape(B, OH * OW, C * k * k)
    return np.ascontiguousarray(cols), OH, OW


def _col2im(dcols, xshape, k, s, p, OH, OW):
    B, C, H, W = xshape
    dxp = np.zeros((B, C, H + 2 * p, W + 2 * p), dcols.dtype)
    d = dcols.reshape(B, OH, OW, C, k, k)
    for i in range(k):
        for j in range(k):
            dxp[:, :, i : i + s * OH : s, j : j + s * OW : s] += d[:, :, :, :, i, j].transpose(0, 3, 1, 2)
    return dxp[:, :, p : p + H, p : p + W]


def chain_forward(layers, params, x, dtype=F32):
    """Returns (y, cache).  CrossCor = true cross-correlation, symmetric zero pad, out size
    floor((i+2p-k)/s)+1; Dense y = σ(W x + b); ``x ./ 255`` is an IEEE divide."""
    cache = []
    pi = 0
    x = x.astype(dtype)
    for l in layers:
        if l[0] == "scale255":
            cache.append(None)
            x = x / dtype(255)
        elif l[0] == "conv":
            _, cin, cout, k, s, p, relu = l
            w, b = params[pi].astype(dtype), params[pi + 1].astype(dtype)
            pi += 2
            cols, OH, OW = _im2col(x, k, s, p)
            y = cols @ w.reshape(cout, -1).T + b
            if relu:
                y = np.maximum(y, 0)
            cache.append((x.shape, cols, y, OH, OW))
            x = np.ascontiguousarray(y.transpose(0, 2, 1)).reshape(x.shape[0], cout, OH, OW)
        elif l[0] == "flatten":
            cache.append(x.shape)
            x = x.reshape(x.shape[0], -1)
        elif l[0] == "dense":
            _, nin, nout, relu = l
            w, b = params[pi].astype(dtype), params[pi + 1].astype(dtype)
            pi += 2
            y = x @ w.T + b
            if relu:
                y = np.maximum(y, 0)
            cache.append((x, y))
            x = y
    return x, cache


def chain_backward(layers, params, cache, dy, need_dx=False):
    """Reverse-mode through a chain.  relu' = 1{y>0}.  Returns (grads list, dx or None)."""
    dtype = dy.dtype
    n_p = len(chain_param_shapes(layers))
    grads = [None] * n_p
    pi = n_p
    first_param_layer = next((i for i, l in enumerate(layers) if l[0] in ("conv", "dense")), len(layers))
    for li in range(len(layers) - 1, -1, -1):
        l = layers[li]
        c = cache[li]
        if l[0] == "scale255":
            if need_dx:
                dy = dy / dtype.type(255)
        elif l[0] == "flatten":
            dy = dy.reshape(c)
        elif l[0] == "dense":
            _, nin, nout, relu = l
            pi -= 2
            x, y = c
            w = params[pi].astype(dtype)
            if relu:
                dy = dy * (y > 0)
            grads[pi] = dy.T @ x
            grads[pi + 1] = dy.sum(0)
            if need_dx or li > first_param_layer:
                dy = dy @ w
        elif l[0] == "conv":
            _, cin, cout, k, s, p, relu = l
            pi -= 2
            xshape, cols, y, OH, OW = c
            w = params[pi].astype(dtype)
            B = xshape[0]
            dym = np.ascontiguousarray(dy.reshape(B, cout, OH * OW).transpose(0, 2, 1))
            if relu:
                dym = dym * (y > 0)
            grads[pi] = np.einsum("bpo,bpk->ok", dym, cols, optimize=True).reshape(w.shape)
            grads[pi + 1] = dym.sum((0, 1))
            if need_dx or li > first_param_layer:
                dcols = dym @ w.reshape(cout, -1)
                dy = _col2im(dcols, xshape, k, s, p, OH, OW)
    return grads, (dy if need_dx else None)


def embed(tau, n_em, dtype=F32):
    """``embed(x, Nₑₘ) = cos.(Float32(π) .* (1:Nₑₘ) .* reshape(x, 1, :))`` -- iqn.jl:157.
    tau [B, N] -> [B*N, n_em] (quantile-fastest rows == Julia columns n + N*b)."""
    t = tau.reshape(-1, 1).astype(dtype)
    i = np.arange(1, n_em + 1, dtype=dtype).reshape(1, -1)
    return np.cos((dtype(np.float32(np.pi)) * i) * t).astype(dtype)


def net_forward(net, params, x, tau=None, dtype=F32):
    """Forward of any net kind.  Returns (out, cache).
    chain   -> out [B, n_out]
    dueling -> val .+ adv .- mean(adv, dims=1)                    common.jl:51-56
    iqn     -> out [B, N, A]  (Julia (A, N, B))                   iqn.jl:25-33
    """
    kind = net["kind"]
    if kind == "chain":
        return chain_forward(net["layers"], params, x, dtype)
    if kind == "dueling":
        nb = len(chain_param_shapes(net["base"]))
        nv = len(chain_param_shapes(net["val"]))
        h, cb = chain_forward(net["base"], params[:nb], x, dtype)
        v, cv = chain_forward(net["val"], params[nb : nb + nv], h, dtype)
        a, ca = chain_forward(net["adv"], params[nb + nv :], h, dtype)
        out = v + a - a.mean(axis=1, keepdims=True, dtype=dtype)
        return out, (cb, cv, ca, nb, nv)
    if kind == "iqn":
        npsi = len(chain_param_shapes(net["psi"]))
        nphi = len(chain_param_shapes(net["phi"]))
        n_em = net["phi"][0][1]
        B, N = tau.shape
        feat, cpsi = chain_forward(net["psi"], params[:npsi], x, dtype)  # [B, F]
        emb = embed(tau, n_em, dtype)  # [B*N, n_em]
        ea, cphi = chain_forward(net["phi"], params[npsi : npsi + nphi], emb, dtype)  # [B*N, F]
        merged = feat[:, None, :] * ea.reshape(B, N, -1)  # iqn.jl:28-30
        z, chd = chain_forward(net["header"], params[npsi + nphi :], merged.reshape(B * N, -1), dtype)
        return z.reshape(B, N, -1), (cpsi, cphi, chd, feat, ea, npsi, nphi, B, N)
    raise ValueError(kind)


def net_backward(net, params, cache, dout):
    """Gradient w.r.t. all params given d(out)."""
    kind = net["kind"]
    if kind == "chain":
        return chain_backward(net["layers"], params, cache, dout)[0]
    if kind == "dueling":
        cb, cv, ca, nb, nv = cache
        dv = dout.sum(axis=1, keepdims=True)
        da = dout - dout.mean(axis=1, keepdims=True)
        gv, dh1 = chain_backward(net["val"], params[nb : nb + nv], cv, dv, need_dx=True)
        ga, dh2 = chain_backward(net["adv"], params[nb + nv :], ca, da, need_dx=True)
        gb, _ = chain_backward(net["base"], params[:nb], cb, dh1 + dh2)
        return gb + gv + ga
    if kind == "iqn":
        cpsi, cphi, chd, feat, ea, npsi, nphi, B, N = cache
        gh, dmerged = chain_backward(net["header"], params[npsi + nphi :], chd, dout.reshape(B * N, -1), need_dx=True)
        dm = dmerged.reshape(B, N, -1)
        dfeat = (dm * ea.reshape(B, N, -1)).sum(1)
        dea = (dm * feat[:, None, :]).reshape(B * N, -1)
        gphi, _ = chain_backward(net["phi"], params[npsi : npsi + nphi], cphi, dea)
        gpsi, _ = chain_backward(net["psi"], params[:npsi], cpsi, dfeat)
        return gpsi + gphi + gh
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# losses (Flux 0.11)
# --------------------------------------------------------------------------------------
def huber_per_sample(G, q, delta=1.0):
    """Flux ``huber_loss(ŷ=G, y=q; agg=identity, δ=1)``: strict ``<`` mask (q4).
    Returns (per-sample loss, d loss_b / d q_b, d loss_b / d G_b)."""
    dt = G.dtype.type
    e = G - q
    ae = np.abs(e)
    mask = ae < dt(delta)
    per = np.where(mask, dt(0.5) * ae * ae, dt(delta) * (ae - dt(0.5) * dt(delta)))
    dG = np.where(mask, e, dt(delta) * np.sign(e))
    return per.astype(G.dtype), (-dG).astype(G.dtype), dG.astype(G.dtype)


def per_weights(priority, beta):
    """``w = 1f0 ./ ((p .+ 1f-10) .^ β); w ./= maximum(w)`` -- prioritized_dqn.jl:125-126,
    rainbow.jl:171-172, iqn.jl:197-198 (host, fp32 pow, q10)."""
    p = priority.astype(F32)
    w = F32(1.0) / np.power(p + F32(1e-10), F32(beta), dtype=F32)
    return (w / w.max()).astype(F32)


def new_priorities(per_sample_loss, beta):
    """``(batch_losses .+ 1f-10) .^ β`` -- prioritized_dqn.jl:143, rainbow.jl:186, iqn.jl:228."""
    return np.power(per_sample_loss.astype(F32) + F32(1e-10), F32(beta), dtype=F32)


def gamma_n(gamma, n, dtype=F32):
    """``γ^n`` with γ::Float32, n::Int (dqn.jl:133).  Evaluated in fp64 and rounded once."""
    return dtype(np.float64(np.float32(gamma)) ** int(n))


def _first_argmax(q):
    """Julia ``argmax`` / Zygote ``maximum`` adjoint: first maximum (q8)."""
    return np.argmax(q, axis=1)


def _mask_add(q, mask):
    """``q .+= ifelse.(mask, 0f0, typemin(Float32))`` -- dqn.jl:121-124."""
    if mask is None:
        return q
    return q + np.where(mask.astype(bool), q.dtype.type(0), q.dtype.type(-np.inf))


# --------------------------------------------------------------------------------------
# learner updates: return dict(loss, grads, priorities?, aux)
# --------------------------------------------------------------------------------------
def basic_dqn_grads(net, params, batch, gamma=0.99, dtype=F32):
    """``update!(::BasicDQNLearner, batch)`` -- basic_dqn.jl:69-90.  The TD target is built
    *inside* the gradient closure, so the gradient flows through Q(s) and Q(s′) (q3)."""
    s, a, r, t, s2 = batch["state"], batch["action"], batch["reward"], batch["terminal"], batch["next_state"]
    B = len(a)
    q_all, c1 = net_forward(net, params, s, dtype=dtype)
    q2_all, c2 = net_forward(net, params, s2, dtype=dtype)
    q = q_all[np.arange(B), a]
    amax = _first_argmax(q2_all)
    q2 = q2_all[np.arange(B), amax]
    disc = dtype(np.float32(gamma)) * (1 - t.astype(dtype))
    G = r.astype(dtype) + disc * q2
    per, dq, dG = huber_per_sample(G, q)
    loss = per.mean(dtype=dtype)
    d1 = np.zeros_like(q_all)
    d1[np.arange(B), a] = dq / dtype(B)
    d2 = np.zeros_like(q2_all)
    d2[np.arange(B), amax] = dG * disc / dtype(B)
    g1 = net_backward(net, params, c1, d1)
    g2 = net_backward(net, params, c2, d2)
    return {"loss": dtype(loss), "grads": [x + y for x, y in zip(g1, g2)], "G": G, "q": q}


def dqn_target(net, params, tparams, batch, gamma, n, double_dqn, dtype=F32):
    """dqn.jl:115-133 (double, default on -- q6) / prioritized_dqn.jl:129-135 (max of target)."""
    s2 = batch["next_state"]
    B = s2.shape[0]
    mask = batch.get("next_legal_actions_mask")
    if double_dqn:
        qv, _ = net_forward(net, params, s2, dtype=dtype)
        qv = _mask_add(qv, mask)
        sel = _first_argmax(qv)
        qt, _ = net_forward(net, tparams, s2, dtype=dtype)
        q2 = qt[np.arange(B), sel]
    else:
        qt, _ = net_forward(net, tparams, s2, dtype=dtype)
        qt = _mask_add(qt, mask)
        q2 = qt.max(axis=1)
    G = batch["reward"].astype(dtype) + gamma_n(gamma, n, dtype) * (1 - batch["terminal"].astype(dtype)) * q2
    return G.astype(dtype)


def dqn_grads(net, params, tparams, batch, gamma=0.99, n=1, double_dqn=True, dtype=F32):
    """``update!(::DQNLearner, batch)`` -- dqn.jl:102-145; loss = huber_loss(G, q) (mean)."""
    a = batch["action"]
    B = len(a)
    G = dqn_target(net, params, tparams, batch, gamma, n, double_dqn, dtype)
    q_all, c = net_forward(net, params, batch["state"], dtype=dtype)
    q = q_all[np.arange(B), a]
    per, dq, _ = huber_per_sample(G, q)
    loss = per.mean(dtype=dtype)
    d = np.zeros_like(q_all)
    d[np.arange(B), a] = dq / dtype(B)
    return {"loss": dtype(loss), "grads": net_backward(net, params, c, d), "G": G, "q": q}


def prioritized_dqn_grads(net, params, tparams, batch, gamma=0.99, n=1, beta=0.5, dtype=F32):
    """``update!(::PrioritizedDQNLearner, batch)`` -- prioritized_dqn.jl:111-151.
    No double-DQN; loss = dot(w, ℓ) * 1//B (q7); returns (ℓ + 1f-10)^β."""
    a = batch["action"]
    B = len(a)
    w = per_weights(batch["priority"], beta).astype(dtype)
    G = dqn_target(net, params, tparams, batch, gamma, n, False, dtype)
    q_all, c = net_forward(net, params, batch["state"], dtype=dtype)
    q = q_all[np.arange(B), a]
    per, dq, _ = huber_per_sample(G, q)
    loss = np.dot(w, per) / dtype(B)
    d = np.zeros_like(q_all)
    d[np.arange(B), a] = dq * w / dtype(B)
    return {
        "loss": dtype(loss),
        "grads": net_backward(net, params, c, d),
        "priorities": new_priorities(per, beta),
        "per_sample": per,
        "weights": w,
        "G": G,
        "q": q,
    }


def softmax(x, axis=-1):
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)


def logsoftmax(x, axis=-1):
    m = x.max(axis=axis, keepdims=True)
    z = x - m
    return z - np.log(np.exp(z).sum(axis=axis, keepdims=True))


def project_distribution(supports, weights, target_support, delta_z, vmin, vmax):
    """rainbow.jl:204-221, dense triangular-kernel form (NOT an index-arithmetic scatter):
    out[b, j] = Σ_i w[b,i] · clamp(1 − |clamp(Tz[b,i], vmin, vmax) − z[j]| / Δz, 0, 1).
    supports, weights: [B, n_atoms]; target_support: [n_atoms]."""
    dt = supports.dtype.type
    c = np.clip(supports, dt(vmin), dt(vmax))  # [B, i]
    k = np.clip(dt(1) - np.abs(c[:, :, None] - target_support[None, None, :]) / dt(delta_z), dt(0), dt(1))
    return (k * weights[:, :, None]).sum(axis=1).astype(supports.dtype)


def rainbow_grads(net, params, tparams, batch, n_actions, n_atoms=51, vmax=10.0, vmin=-10.0, gamma=0.99,
                  n=1, beta=0.5, support=None, dtype=F32):
    """``update!(::RainbowLearner, batch)`` -- rainbow.jl:126-196.  Net output per sample is
    n_atoms*n_actions logits, atom index fastest (reshape(…, n_atoms, n_actions, :), rainbow.jl:177).
    Default support = range(-Vmax, Vmax) ignores Vmin (q5, rainbow.jl:72).  No double-DQN selection.
    PER iff the batch has a ``priority`` entry (rainbow.jl:168)."""
    a = batch["action"]
    B = len(a)
    if support is None:
        support = np.linspace(np.float32(-vmax), np.float32(vmax), n_atoms).astype(F32)  # range(Float32...)
    support = support.astype(dtype)
    delta_z = dtype(np.float32(support[1] - support[0])) if dtype is F32 else dtype(support[1] - support[0])
    disc = gamma_n(gamma, n, dtype) * (1 - batch["terminal"].astype(dtype))  # [B]
    tsupport = batch["reward"].astype(dtype)[:, None] + support[None, :] * disc[:, None]  # rainbow.jl:146-148
    next_logits, _ = net_forward(net, tparams, batch["next_state"], dtype=dtype)
    next_probs = softmax(next_logits.reshape(B, n_actions, n_atoms), axis=2)  # rainbow.jl:150-151
    next_q = (support[None, None, :] * next_probs).sum(axis=2)  # rainbow.jl:152
    next_q = _mask_add(next_q, batch.get("next_legal_actions_mask"))
    sel = _first_argmax(next_q)  # select_best_probs, rainbow.jl:198-202
    next_prob_select = next_probs[np.arange(B), sel]  # [B, n_atoms]
    target = project_distribution(tsupport, next_prob_select, support, delta_z, vmin, vmax)
    use_per = "priority" in batch and batch["priority"] is not None
    logits_all, c = net_forward(net, params, batch["state"], dtype=dtype)
    logits = logits_all.reshape(B, n_actions, n_atoms)
    sl = logits[np.arange(B), a]  # [B, n_atoms]
    lsm = logsoftmax(sl, axis=1)
    per = -(target * lsm).sum(axis=1)  # logitcrossentropy(...; agg=identity), Dopamine_Rainbow_Atari.jl:54
    if use_per:
        w = per_weights(batch["priority"], beta).astype(dtype)
        loss = np.dot(w, per) / dtype(B)
        scale = w / dtype(B)
    else:
        loss = per.mean(dtype=dtype)
        scale = np.full(B, dtype(1) / dtype(B), dtype)
    # d/dŷ = softmax(ŷ)·Σy − y   (Σy is 1 only up to rounding -> general form)
    dsl = (np.exp(lsm) * target.sum(axis=1, keepdims=True) - target) * scale[:, None]
    d = np.zeros_like(logits)
    d[np.arange(B), a] = dsl
    out = {
        "loss": dtype(loss),
        "grads": net_backward(net, params, c, d.reshape(B, -1)),
        "per_sample": per,
        "target_distribution": target,
        "selected": sel,
    }
    if use_per:
        out["priorities"] = new_priorities(per, beta)
    return out


def quantile_huber(target, q, tau, kappa=1.0):
    """iqn.jl:207-219.  target [B, N′], q [B, N], tau [B, N].
    TD[b,i,j] = target[b,i] − q[b,j]; raw = |τ[b,j] − 1{TD<0}| · huber_κ(TD) / κ;
    per-sample = mean_j Σ_i raw  (sum over N′, mean over N: q2).
    Returns (per_sample [B], d per_sample_b / d q[b,j])."""
    dt = target.dtype.type
    k = dt(kappa)
    TD = target[:, :, None] - q[:, None, :]
    ae = np.abs(TD)
    quad = np.minimum(ae, k)
    lin = ae - quad
    hub = dt(0.5) * quad * quad + k * lin
    wgt = np.abs(tau[:, None, :] - (TD < 0).astype(target.dtype))
    raw = wgt * hub / k
    N = q.shape[1]
    per = raw.sum(axis=1).mean(axis=1, dtype=target.dtype)
    dq = -(wgt * np.clip(TD, -k, k) / k).sum(axis=1) / dt(N)
    return per.astype(target.dtype), dq.astype(target.dtype)


def iqn_grads(net, params, tparams, batch, tau, tau_prime, gamma=0.99, kappa=1.0, beta=0.5, dtype=F32):
    """``update!(::IQNLearner, batch)`` -- iqn.jl:159-237.  τ, τ′ are inputs ([B, N], [B, N′]) because
    the CURAND stream (iqn.jl:173,191) is not reproducible.  Bootstraps with γ, not γⁿ (q1, iqn.jl:189)."""
    a = batch["action"]
    B = len(a)
    zt, _ = net_forward(net, tparams, batch["next_state"], tau=tau_prime, dtype=dtype)  # [B, N′, A]
    avg = zt.mean(axis=1, dtype=dtype)  # iqn.jl:176
    avg = _mask_add(avg, batch.get("next_legal_actions_mask"))
    at = _first_argmax(avg)  # iqn.jl:184
    qt = zt[np.arange(B), :, at]  # [B, N′]
    target = batch["reward"].astype(dtype)[:, None] + dtype(np.float32(gamma)) * (1 - batch["terminal"].astype(dtype))[:, None] * qt
    z, c = net_forward(net, params, batch["state"], tau=tau, dtype=dtype)  # [B, N, A]
    q = z[np.arange(B), :, a]  # [B, N]
    per, dq = quantile_huber(target.astype(dtype), q, tau.astype(dtype), kappa)
    use_per = "priority" in batch and batch["priority"] is not None
    if use_per:
        w = per_weights(batch["priority"], beta).astype(dtype)
        loss = np.dot(w, per) / dtype(B)
        scale = w / dtype(B)
    else:
        loss = per.mean(dtype=dtype)
        scale = np.full(B, dtype(1) / dtype(B), dtype)
    dz = np.zeros_like(z)
    dz[np.arange(B), :, a] = dq * scale[:, None]
    out = {"loss": dtype(loss), "grads": net_backward(net, params, c, dz), "per_sample": per, "target": target, "selected": at}
    if use_per:
        out["priorities"] = new_priorities(per, beta)
    return out


def qr_cum_prob(N, dtype=F32):
    """``range(0.5f0 / N; length = N, step = 1.0f0 / N)`` -- qr_dqn.jl:11 (Float32 range: reference value and step are
    Float32, the elements are evaluated in higher precision and rounded once)."""
    first, step = np.float64(F32(0.5) / F32(N)), np.float64(F32(1.0) / F32(N))
    return (first + np.arange(N, dtype=np.float64) * step).astype(dtype)


def quantile_huber_loss_qr(yhat, y, kappa=1.0):
    """``quantile_huber_loss(ŷ, y; κ)`` -- qr_dqn.jl:3-14.  ŷ, y: [B, N].  Δ[b,i,j] = y[b,i] − ŷ[b,j]; the cumulative
    probabilities are broadcast along the FIRST Julia axis, i.e. the target index i (as written in the reference);
    loss = mean over (j, b) of Σ_i.  Returns (per-sample loss [B] = mean_j Σ_i, d per_sample_b / d ŷ[b,j])."""
    dt = y.dtype.type
    k = dt(kappa)
    N = y.shape[1]
    D = y[:, :, None] - yhat[:, None, :]
    ae = np.abs(D)
    quad = np.minimum(ae, k)
    lin = ae - quad
    hub = dt(0.5) * quad * quad + k * lin
    cp = qr_cum_prob(N, y.dtype)
    w = np.abs(cp[None, :, None] - (D < 0).astype(y.dtype))
    per = (w * hub).sum(axis=1).mean(axis=1, dtype=y.dtype)
    dyhat = -(w * np.clip(D, -k, k)).sum(axis=1) / dt(N)
    return per.astype(y.dtype), dyhat.astype(y.dtype)


def qr_dqn_grads(net, params, tparams, batch, n_actions, N, gamma=0.99, kappa=1.0, dtype=F32):
    """``update!(::QRDQNLearner, batch)`` -- qr_dqn.jl:106-138.  Net output per sample is N*n_actions values, quantile index
    fastest (reshape(Q(s), N, :, B)).  Target: mean over quantiles -> first-max action -> y = r + γ(1−t)·z_t[:, a*]
    (γ, not γⁿ, as written: qr_dqn.jl:123).  No legal-action mask, no PER."""
    a = batch["action"]
    B = len(a)
    zt, _ = net_forward(net, tparams, batch["next_state"], dtype=dtype)
    zt = zt.reshape(B, n_actions, N)
    at = _first_argmax(zt.mean(axis=2, dtype=dtype))
    y = batch["reward"].astype(dtype)[:, None] + dtype(np.float32(gamma)) * (1 - batch["terminal"].astype(dtype))[:, None] * zt[np.arange(B), at]
    z_all, c = net_forward(net, params, batch["state"], dtype=dtype)
    z = z_all.reshape(B, n_actions, N)
    per, dq = quantile_huber_loss_qr(z[np.arange(B), a], y.astype(dtype), kappa)
    d = np.zeros_like(z)
    d[np.arange(B), a] = dq / dtype(B)
    return {"loss": dtype(per.mean(dtype=dtype)), "grads": net_backward(net, params, c, d.reshape(B, -1)), "per_sample": per, "y": y}


def rem_dqn_grads(net, params, tparams, batch, n_actions, E, w, gamma=0.99, n=1, dtype=F32):
    """``update!(::REMDQNLearner, batch)`` -- rem_dqn.jl:100-145.  Net output per sample is n_actions*E values, action index
    fastest (reshape(q, :, E, B)); w [E] is the convex combination (``rand`` normalised, or ones/E), an input here because
    the reference draws it from a host RNG.  loss_func = huber_loss (JuliaRL_REMDQN_CartPole.jl:40)."""
    a = batch["action"]
    B = len(a)
    w = np.asarray(w, dtype)
    qt, _ = net_forward(net, tparams, batch["next_state"], dtype=dtype)
    tq = (w[None, :, None] * qt.reshape(B, E, n_actions)).sum(axis=1, dtype=dtype)
    tq = _mask_add(tq, batch.get("next_legal_actions_mask"))
    G = batch["reward"].astype(dtype) + gamma_n(gamma, n, dtype) * (1 - batch["terminal"].astype(dtype)) * tq.max(axis=1)
    q_all, c = net_forward(net, params, batch["state"], dtype=dtype)
    qe = q_all.reshape(B, E, n_actions)
    q = (w[None, :, None] * qe).sum(axis=1, dtype=dtype)[np.arange(B), a]
    per, dq, _ = huber_per_sample(G.astype(dtype), q)
    d = np.zeros_like(qe)
    d[np.arange(B), :, a] = (dq / dtype(B))[:, None] * w[None, :]
    return {"loss": dtype(per.mean(dtype=dtype)), "grads": net_backward(net, params, c, d.reshape(B, -1)), "per_sample": per, "G": G, "q": q}


# --------------------------------------------------------------------------------------
# optimisers (Flux 0.11): Float64 scalars x Float32 arrays -> each fused
# broadcast line is evaluated per element in Float64 and rounded once on store.
# --------------------------------------------------------------------------------------
EPS = 1e-8


def rmsprop_init(params):
    return {"acc": [np.zeros_like(p) for p in params]}


def rmsprop_step(params, grads, state, eta=0.00025, rho=0.95):
    """``RMSProp(η, ρ)``: acc = ρ·acc + (1−ρ)·Δ²;  Δ *= η/(√acc + ε);  x .-= Δ.
    No centering, no momentum (Dopamine_DQN_Atari.jl:42-43).  Δ^2 and √acc are Float32 ops
    (literal_pow / sqrt of a Float32 element) before promotion."""
    new = []
    for i, (p, g) in enumerate(zip(params, grads)):
        g = g.astype(F32)
        g2 = (g * g).astype(np.float64)
        acc = (rho * state["acc"][i].astype(np.float64) + (1 - rho) * g2).astype(F32)
        state["acc"][i] = acc
        step = (g.astype(np.float64) * (eta / (np.sqrt(acc).astype(np.float64) + EPS))).astype(F32)
        new.append((p - step).astype(F32))
    return new


def adam_init(params, beta=(0.9, 0.999)):
    return {"m": [np.zeros_like(p) for p in params], "v": [np.zeros_like(p) for p in params],
            "bp": [(beta[0], beta[1]) for _ in params]}


def adam_step(params, grads, state, eta=0.001, beta=(0.9, 0.999)):
    """``ADAM(η, β)``: m = β₁m + (1−β₁)Δ; v = β₂v + (1−β₂)Δ²;
    Δ = m/(1−β₁ᵗ)/(√(v/(1−β₂ᵗ)) + ε)·η; βᵗ kept per parameter tensor, starts at β."""
    new = []
    for i, (p, g) in enumerate(zip(params, grads)):
        g = g.astype(F32)
        b1p, b2p = state["bp"][i]
        m = (beta[0] * state["m"][i].astype(np.float64) + (1 - beta[0]) * g.astype(np.float64)).astype(F32)
        v = (beta[1] * state["v"][i].astype(np.float64) + (1 - beta[1]) * (g * g).astype(np.float64)).astype(F32)
        state["m"][i], state["v"][i] = m, v
        step = (m.astype(np.float64) / (1 - b1p) / (np.sqrt(v.astype(np.float64) / (1 - b2p)) + EPS) * eta).astype(F32)
        state["bp"][i] = (b1p * beta[0], b2p * beta[1])
        new.append((p - step).astype(F32))
    return new


# --------------------------------------------------------------------------------------
# synthetic batches
# --------------------------------------------------------------------------------------
def synth_batch(seed, B, obs_shape, n_actions, atari=True, priority=False, mask=False):
    rng = np.random.default_rng(seed)
    if atari:
        s = rng.integers(0, 256, (B,) + tuple(obs_shape), dtype=np.uint8)
        s2 = rng.integers(0, 256, (B,) + tuple(obs_shape), dtype=np.uint8)
        r = rng.integers(-1, 2, B).astype(F32)
    else:
        s = rng.standard_normal((B,) + tuple(obs_shape)).astype(F32)
        s2 = rng.standard_normal((B,) + tuple(obs_shape)).astype(F32)
        r = np.ones(B, F32)
    batch = {
        "state": s,
        "next_state": s2,
        "action": rng.integers(0, n_actions, B).astype(np.int32),
        "reward": r,
        "terminal": (rng.random(B) < 0.1).astype(np.uint8),
    }
    if priority:
        p = np.power(rng.random(B, dtype=np.float32) + F32(1e-10), F32(0.5)).astype(F32)
        p[rng.random(B) < 0.1] = F32(100.0)
        batch["priority"] = p
    if mask:
        m = (rng.random((B, n_actions)) < 0.7).astype(np.uint8)
        m[np.arange(B), rng.integers(0, n_actions, B)] = 1
        batch["next_legal_actions_mask"] = m
    return batch
