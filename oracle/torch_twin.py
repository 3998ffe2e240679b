"""Independent PyTorch-CPU autograd twin of ``oracle/zoo_oracle.py``.

TEST INFRASTRUCTURE ONLY (see the header of ``zoo_oracle.py``): never imported by the product.
It restates the same reference lines (cited per function) but shares NO code with the numpy
oracle: layers come from ``torch.nn.functional`` and every gradient comes from autograd
instead of hand-written adjoints.  Agreement of the two (tests/test_oracle_twin.py) is what
stands in for the golden vectors the reference does not have (PARITY UNPINNED).

It is also the CPU arm that ``bench.py`` times on the GPU box's host cores
(``cpu_baseline.kind = "port"`` and ``--impl reference``): a multi-threaded MKL/oneDNN fp32
execution of the same update, the closest stand-in available for "the reference's CPU Flux
path" when ``julia`` cannot run.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as Fn


def _act(pre, masks, rec):
    """relu -- or, for the full-size parity tests, `pre * mask` with the relu mask the DEVICE used for this layer (so that a
    unit whose pre-activation is within rounding of zero takes the same branch in both); `rec` collects the pre-activations."""
    if rec is not None:
        rec.append(pre.detach())
    if masks is not None:
        return pre * masks.pop(0)
    return torch.relu(pre)


def _chain(layers, params, x, masks=None, rec=None):
    pi = 0
    for l in layers:
        if l[0] == "scale255":  # Dopamine_DQN_Atari.jl:28
            x = x / 255.0
        elif l[0] == "conv":  # CrossCor == torch conv2d (no flip), Dopamine_DQN_Atari.jl:29-31
            _, cin, cout, k, s, p, relu = l
            x = Fn.conv2d(x, params[pi], params[pi + 1], stride=s, padding=p)
            pi += 2
            if relu:
                x = _act(x, masks, rec)
        elif l[0] == "flatten":
            x = x.flatten(1)
        elif l[0] == "dense":
            _, nin, nout, relu = l
            x = Fn.linear(x, params[pi], params[pi + 1])
            pi += 2
            if relu:
                x = _act(x, masks, rec)
    return x, pi


def _nparams(layers):
    return 2 * sum(1 for l in layers if l[0] in ("conv", "dense"))


def forward(net, params, x, tau=None, masks=None, rec=None):
    """masks: optional list of 0/1 tensors, one per relu layer in evaluation order (consumed); rec: optional list that
    receives every relu layer's pre-activation (see _act)."""
    kind = net["kind"]
    if kind == "chain":
        return _chain(net["layers"], params, x, masks, rec)[0]
    if kind == "dueling":  # common.jl:51-56
        nb, nv = _nparams(net["base"]), _nparams(net["val"])
        h = _chain(net["base"], params[:nb], x, masks, rec)[0]
        v = _chain(net["val"], params[nb : nb + nv], h, masks, rec)[0]
        a = _chain(net["adv"], params[nb + nv :], h, masks, rec)[0]
        return v + a - a.mean(dim=1, keepdim=True)
    if kind == "iqn":  # iqn.jl:25-33,157
        npsi, nphi = _nparams(net["psi"]), _nparams(net["phi"])
        n_em = net["phi"][0][1]
        B, N = tau.shape
        feat = _chain(net["psi"], params[:npsi], x, masks, rec)[0]
        i = torch.arange(1, n_em + 1, dtype=tau.dtype)
        emb = torch.cos((torch.tensor(np.float32(np.pi), dtype=tau.dtype) * i)[None, :] * tau.reshape(-1, 1))
        ea = _chain(net["phi"], params[npsi : npsi + nphi], emb, masks, rec)[0]
        merged = feat[:, None, :] * ea.reshape(B, N, -1)
        z = _chain(net["header"], params[npsi + nphi :], merged.reshape(B * N, -1), masks, rec)[0]
        return z.reshape(B, N, -1)
    raise ValueError(kind)


def _t(x, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.from_numpy(np.ascontiguousarray(x)).to(dtype)


def _leaf(params, dtype):
    # already-leaf tensors (the CPU timing arm keeps its parameters as torch leaves) are used in place
    if all(isinstance(p, torch.Tensor) and p.requires_grad for p in params):
        return list(params)
    return [_t(p, dtype).clone().requires_grad_(True) for p in params]


def _huber(G, q):
    # Flux huber_loss(ŷ=G, y=q; δ=1): strict < mask, mask not differentiated (q4)
    ae = (G - q).abs()
    mask = (ae < 1.0).to(G.dtype).detach()
    return 0.5 * ae * ae * mask + (ae - 0.5) * (1 - mask)


def _mask_add(q, batch):
    m = batch.get("next_legal_actions_mask")
    if m is None:
        return q
    m = torch.from_numpy(np.asarray(m).astype(bool))
    return q + torch.where(m, torch.zeros((), dtype=q.dtype), torch.full((), -math.inf, dtype=q.dtype))


def _per_w(batch, beta, dtype):
    p = torch.from_numpy(batch["priority"].astype(np.float32))
    w = 1.0 / (p + 1e-10) ** np.float32(beta)
    return (w / w.max()).to(dtype)


KEEP_TORCH_GRADS = False  # the CPU timing arm sets this: leave gradients in p.grad, skip the numpy copies


def _finish(loss, ps, extra):
    loss.backward()
    if KEEP_TORCH_GRADS:
        return {"loss": loss.detach()}
    out = {"loss": loss.detach().numpy().copy(), "grads": [p.grad.numpy().copy() if p.grad is not None else np.zeros(p.shape, np.float32) for p in ps]}
    out.update(extra)
    return out


def _gn(gamma, n):
    return float(np.float64(np.float32(gamma)) ** int(n))


def basic_dqn_grads(net, params, batch, gamma=0.99, dtype=torch.float32):
    """basic_dqn.jl:69-90 (gradient through the target, q3)."""
    ps = _leaf(params, dtype)
    a = torch.from_numpy(batch["action"].astype(np.int64))
    r, t = _t(batch["reward"], dtype), _t(batch["terminal"], dtype)
    q = forward(net, ps, _t(batch["state"], dtype)).gather(1, a[:, None])[:, 0]
    q2 = forward(net, ps, _t(batch["next_state"], dtype)).max(dim=1).values
    G = r + float(np.float32(gamma)) * (1 - t) * q2
    return _finish(_huber(G, q).mean(), ps, {})


def _dqn_target(net, ps, tps, batch, gamma, n, double, dtype):
    s2 = _t(batch["next_state"], dtype)
    with torch.no_grad():
        if double:  # dqn.jl:115-128
            sel = _mask_add(forward(net, ps, s2), batch).argmax(dim=1)
            q2 = forward(net, tps, s2).gather(1, sel[:, None])[:, 0]
        else:  # dqn.jl:130 / prioritized_dqn.jl:129-135
            q2 = _mask_add(forward(net, tps, s2), batch).max(dim=1).values
        return _t(batch["reward"], dtype) + _gn(gamma, n) * (1 - _t(batch["terminal"], dtype)) * q2


def dqn_grads(net, params, tparams, batch, gamma=0.99, n=1, double_dqn=True, dtype=torch.float32, masks=None, rec=None):
    """dqn.jl:102-145."""
    ps, tps = _leaf(params, dtype), [_t(p, dtype) for p in tparams]
    a = torch.from_numpy(batch["action"].astype(np.int64))
    G = _dqn_target(net, ps, tps, batch, gamma, n, double_dqn, dtype)
    q = forward(net, ps, _t(batch["state"], dtype), masks=masks, rec=rec).gather(1, a[:, None])[:, 0]
    return _finish(_huber(G, q).mean(), ps, {})


def prioritized_dqn_grads(net, params, tparams, batch, gamma=0.99, n=1, beta=0.5, dtype=torch.float32, masks=None, rec=None):
    """prioritized_dqn.jl:111-151."""
    ps, tps = _leaf(params, dtype), [_t(p, dtype) for p in tparams]
    a = torch.from_numpy(batch["action"].astype(np.int64))
    B = len(a)
    w = _per_w(batch, beta, dtype)
    G = _dqn_target(net, ps, tps, batch, gamma, n, False, dtype)
    q = forward(net, ps, _t(batch["state"], dtype), masks=masks, rec=rec).gather(1, a[:, None])[:, 0]
    per = _huber(G, q)
    pr = ((per.detach().float() + 1e-10) ** np.float32(beta)).numpy()
    return _finish(torch.dot(w, per) / B, ps, {"priorities": pr, "per_sample": per.detach().numpy()})


def rainbow_grads(net, params, tparams, batch, n_actions, n_atoms=51, vmax=10.0, vmin=-10.0, gamma=0.99, n=1,
                  beta=0.5, support=None, dtype=torch.float32, masks=None, rec=None):
    """rainbow.jl:126-221 -- projection written exactly as the reference's tiled broadcast."""
    ps, tps = _leaf(params, dtype), [_t(p, dtype) for p in tparams]
    a = torch.from_numpy(batch["action"].astype(np.int64))
    B = len(a)
    if support is None:
        support = np.linspace(np.float32(-vmax), np.float32(vmax), n_atoms).astype(np.float32)
    z = _t(support, dtype)
    dz = z[1] - z[0]
    with torch.no_grad():
        disc = _gn(gamma, n) * (1 - _t(batch["terminal"], dtype))
        Tz = _t(batch["reward"], dtype)[:, None] + z[None, :] * disc[:, None]
        nl = forward(net, tps, _t(batch["next_state"], dtype)).reshape(B, n_actions, n_atoms)
        npb = torch.softmax(nl, dim=2)
        nq = _mask_add((z * npb).sum(2), batch)
        sel = nq.argmax(dim=1)
        wsel = npb[torch.arange(B), sel]
        tiled = Tz.clamp(vmin, vmax)[:, :, None].expand(B, n_atoms, n_atoms)  # [b, i, j]
        proj = (1 - (tiled - z[None, None, :]).abs() / dz).clamp(0, 1) * wsel[:, :, None]
        target = proj.sum(1)
    logits = forward(net, ps, _t(batch["state"], dtype), masks=masks, rec=rec).reshape(B, n_actions, n_atoms)
    sl = logits[torch.arange(B), a]
    per = -(target * torch.log_softmax(sl, dim=1)).sum(1)
    extra = {"per_sample": per.detach().numpy(), "target_distribution": target.numpy()}
    if batch.get("priority") is not None:
        loss = torch.dot(_per_w(batch, beta, dtype), per) / B
        extra["priorities"] = ((per.detach().float() + 1e-10) ** np.float32(beta)).numpy()
    else:
        loss = per.mean()
    return _finish(loss, ps, extra)


def iqn_grads(net, params, tparams, batch, tau, tau_prime, gamma=0.99, kappa=1.0, beta=0.5, dtype=torch.float32, masks=None, rec=None):
    """iqn.jl:159-237."""
    ps, tps = _leaf(params, dtype), [_t(p, dtype) for p in tparams]
    a = torch.from_numpy(batch["action"].astype(np.int64))
    B = len(a)
    tau_t, taup_t = _t(tau, dtype), _t(tau_prime, dtype)
    with torch.no_grad():
        zt = forward(net, tps, _t(batch["next_state"], dtype), taup_t)  # [B, N', A]
        at = _mask_add(zt.mean(1), batch).argmax(1)
        qt = zt[torch.arange(B), :, at]
        target = _t(batch["reward"], dtype)[:, None] + float(np.float32(gamma)) * (1 - _t(batch["terminal"], dtype))[:, None] * qt
    z = forward(net, ps, _t(batch["state"], dtype), tau_t, masks=masks, rec=rec)
    q = z[torch.arange(B), :, a]  # [B, N]
    TD = target[:, :, None] - q[:, None, :]
    ae = TD.abs()
    quad = torch.clamp(ae, max=kappa)
    hub = 0.5 * quad * quad + kappa * (ae - quad)
    raw = (tau_t[:, None, :] - (TD < 0).to(dtype).detach()).abs() * hub / kappa
    per = raw.sum(1).mean(1)
    extra = {"per_sample": per.detach().numpy(), "target": target.numpy()}
    if batch.get("priority") is not None:
        loss = torch.dot(_per_w(batch, beta, dtype), per) / B
        extra["priorities"] = ((per.detach().float() + 1e-10) ** np.float32(beta)).numpy()
    else:
        loss = per.mean()
    return _finish(loss, ps, extra)


def qr_dqn_grads(net, params, tparams, batch, n_actions, N, gamma=0.99, kappa=1.0, dtype=torch.float32):
    """qr_dqn.jl:3-14,106-138 with autograd."""
    ps, tps = _leaf(params, dtype), [_t(p, dtype) for p in tparams]
    a = torch.as_tensor(np.asarray(batch["action"], np.int64))
    B = len(a)
    ar = torch.arange(B)
    with torch.no_grad():
        zt = forward(net, tps, _t(batch["next_state"], dtype)).reshape(B, n_actions, N)
        at = zt.mean(dim=2).argmax(dim=1)
        y = _t(batch["reward"], dtype)[:, None] + float(np.float32(gamma)) * (1 - _t(batch["terminal"], dtype))[:, None] * zt[ar, at]
    yhat = forward(net, ps, _t(batch["state"], dtype)).reshape(B, n_actions, N)[ar, a]
    D = y[:, :, None] - yhat[:, None, :]
    ae = D.abs()
    quad = torch.clamp(ae, max=kappa)
    hub = 0.5 * quad * quad + kappa * (ae - quad)
    cp = torch.as_tensor((np.float64(np.float32(0.5) / np.float32(N)) + np.arange(N) * np.float64(np.float32(1.0) / np.float32(N))), dtype=dtype)
    w = (cp[None, :, None] - (D < 0).to(dtype)).abs().detach()
    per = (w * hub).sum(dim=1).mean(dim=1)
    return _finish(per.mean(), ps, {"per_sample": per.detach().numpy().copy()})


def rem_dqn_grads(net, params, tparams, batch, n_actions, E, w, gamma=0.99, n=1, dtype=torch.float32):
    """rem_dqn.jl:100-145 with autograd."""
    ps, tps = _leaf(params, dtype), [_t(p, dtype) for p in tparams]
    a = torch.as_tensor(np.asarray(batch["action"], np.int64))
    B = len(a)
    wt = _t(np.asarray(w), dtype)
    with torch.no_grad():
        tq = (wt[None, :, None] * forward(net, tps, _t(batch["next_state"], dtype)).reshape(B, E, n_actions)).sum(dim=1)
        tq = _mask_add(tq, batch)
        G = _t(batch["reward"], dtype) + _gn(gamma, n) * (1 - _t(batch["terminal"], dtype)) * tq.max(dim=1).values
    q = (wt[None, :, None] * forward(net, ps, _t(batch["state"], dtype)).reshape(B, E, n_actions)).sum(dim=1)[torch.arange(B), a]
    per = _huber(G, q)
    return _finish(per.mean(), ps, {"per_sample": per.detach().numpy().copy()})


# ---- optimisers, Flux 0.11 formulas, plain fp32 torch ops (for the CPU arm) ----
class RMSProp:
    def __init__(self, params, eta=0.00025, rho=0.95):
        self.eta, self.rho = eta, rho
        self.acc = [torch.zeros_like(p) for p in params]

    @torch.no_grad()
    def step(self, params, grads):
        for p, g, acc in zip(params, grads, self.acc):
            acc.mul_(self.rho).addcmul_(g, g, value=1 - self.rho)
            p.sub_(g * (self.eta / (acc.sqrt() + 1e-8)))


class Adam:
    def __init__(self, params, eta=0.001, beta=(0.9, 0.999)):
        self.eta, self.beta = eta, beta
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.bp = [beta[0], beta[1]]

    @torch.no_grad()
    def step(self, params, grads):
        b1, b2 = self.beta
        for p, g, m, v in zip(params, grads, self.m, self.v):
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            p.sub_(m / (1 - self.bp[0]) / ((v / (1 - self.bp[1])).sqrt() + 1e-8) * self.eta)
        self.bp = [self.bp[0] * b1, self.bp[1] * b2]


class CpuLearner:
    """The reference's update!(learner, batch) executed on the host cores with torch (MKL/oneDNN, fp32): the CPU arm
    bench.py times (cpu_baseline / --impl reference).  kind in {"basic", "dqn", "rainbow", "iqn"}."""

    def __init__(self, kind, net, params, tparams, opt, **cfg):
        self.kind, self.net, self.cfg = kind, net, cfg
        self.ps = [torch.from_numpy(np.ascontiguousarray(p)).clone().requires_grad_(True) for p in params]
        self.tps = [torch.from_numpy(np.ascontiguousarray(p)).clone() for p in tparams]
        self.opt = RMSProp(self.ps, *opt[1:]) if opt[0] == "rmsprop" else Adam(self.ps, *opt[1:])

    def update(self, batch, tau=None, tau_prime=None):
        global KEEP_TORCH_GRADS
        KEEP_TORCH_GRADS = True
        try:
            for p in self.ps:
                p.grad = None
            if self.kind == "basic":
                r = basic_dqn_grads(self.net, self.ps, batch, **self.cfg)
            elif self.kind == "dqn":
                r = dqn_grads(self.net, self.ps, self.tps, batch, **self.cfg)
            elif self.kind == "rainbow":
                r = rainbow_grads(self.net, self.ps, self.tps, batch, **self.cfg)
            else:
                r = iqn_grads(self.net, self.ps, self.tps, batch, tau, tau_prime, **self.cfg)
            self.opt.step(self.ps, [p.grad for p in self.ps])
            return float(r["loss"])
        finally:
            KEEP_TORCH_GRADS = False
