"""Host-side mirror of the model / approximator surface the reference learners are built from:
Flux ``Chain/Dense/CrossCor``, ``ImplicitQuantileNet`` (src/algorithms/dqns/iqn.jl:17-33), ``DuelingNetwork``
(src/algorithms/dqns/common.jl:44-56), Flux ``RMSProp/ADAM`` and RLCore's ``NeuralNetworkApproximator``
(used at src/algorithms/dqns/dqn.jl:44-80).  These are descriptions only -- all arithmetic happens in libzoo_sm100."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


# ---- Flux layer descriptions -----------------------------------------------------------------------
class Dense:
    """``Dense(in, out, σ)`` (σ = relu | identity).  W is (out, in), b is (out,)."""

    def __init__(self, n_in, n_out, relu=False):
        self.n_in, self.n_out, self.relu = int(n_in), int(n_out), bool(relu)

    def spec(self):
        return [(L.LAYER_DENSE, self.n_in, self.n_out, 0, 0, 0, int(self.relu))]

    def shapes(self):
        return [(self.n_out, self.n_in), (self.n_out,)]


class CrossCor:
    """``CrossCor((k,k), cin=>cout, σ; stride, pad)``.  Weight is OIHW (== Julia (kw,kh,cin,cout) bytes)."""

    def __init__(self, ksize, n_in, n_out, relu=False, stride=1, pad=0):
        self.k, self.n_in, self.n_out, self.relu, self.stride, self.pad = int(ksize), int(n_in), int(n_out), bool(relu), int(stride), int(pad)

    def spec(self):
        return [(L.LAYER_CONV, self.n_in, self.n_out, self.k, self.stride, self.pad, int(self.relu))]

    def shapes(self):
        return [(self.n_out, self.n_in, self.k, self.k), (self.n_out,)]


class Scale255:
    """``x -> x ./ 255`` (Dopamine_DQN_Atari.jl:28)."""

    def spec(self):
        return [(L.LAYER_SCALE255, 0, 0, 0, 0, 0, 0)]

    def shapes(self):
        return []


class Flatten:
    """``x -> reshape(x, :, size(x)[end])`` (Dopamine_DQN_Atari.jl:32)."""

    def spec(self):
        return [(L.LAYER_FLATTEN, 0, 0, 0, 0, 0, 0)]

    def shapes(self):
        return []


class Chain:
    def __init__(self, *layers):
        self.layers = list(layers)

    def spec(self):
        return [s for l in self.layers for s in l.spec()]

    def shapes(self):
        return [s for l in self.layers for s in l.shapes()]

    kind = L.NET_CHAIN

    def chains(self):
        return [self, None, None]


class ImplicitQuantileNet:
    """``ImplicitQuantileNet(ψ, ϕ, header)`` -- iqn.jl:17-33."""
    kind = L.NET_IQN

    def __init__(self, psi, phi, header):
        self.psi, self.phi, self.header = (x if isinstance(x, Chain) else Chain(x) for x in (psi, phi, header))

    def chains(self):
        return [self.psi, self.phi, self.header]

    def shapes(self):
        return self.psi.shapes() + self.phi.shapes() + self.header.shapes()


class DuelingNetwork:
    """``DuelingNetwork(base, val, adv)`` -- common.jl:44-56."""
    kind = L.NET_DUELING

    def __init__(self, base, val, adv):
        self.base, self.val, self.adv = (x if isinstance(x, Chain) else Chain(x) for x in (base, val, adv))

    def chains(self):
        return [self.base, self.val, self.adv]

    def shapes(self):
        return self.base.shapes() + self.val.shapes() + self.adv.shapes()


def NatureCNN(n_out, in_ch=4, head=True):
    """Dopamine Nature-CNN with the reference's padding (84 -> 21 -> 11 -> 11; fc1 fan-in 7744),
    src/experiments/atari/Dopamine_DQN_Atari.jl:26-35."""
    layers = [Scale255(), CrossCor(8, in_ch, 32, True, 4, 2), CrossCor(4, 32, 64, True, 2, 2), CrossCor(3, 64, 64, True, 1, 1), Flatten()]
    if head:
        layers += [Dense(11 * 11 * 64, 512, True), Dense(512, n_out)]
    return Chain(*layers)


def from_oracle_net(net):
    """oracle-style net dict (tests) -> model description."""
    def chain(layers):
        out = []
        for l in layers:
            if l[0] == "scale255":
                out.append(Scale255())
            elif l[0] == "conv":
                out.append(CrossCor(l[3], l[1], l[2], l[6], l[4], l[5]))
            elif l[0] == "flatten":
                out.append(Flatten())
            else:
                out.append(Dense(l[1], l[2], l[3]))
        return Chain(*out)
    if net["kind"] == "chain":
        return chain(net["layers"])
    if net["kind"] == "iqn":
        return ImplicitQuantileNet(chain(net["psi"]), chain(net["phi"]), chain(net["header"]))
    return DuelingNetwork(chain(net["base"]), chain(net["val"]), chain(net["adv"]))


# ---- Flux optimiser descriptions --------------------------------------------------------------------
class RMSProp:
    """Flux ``RMSProp(η = 0.001, ρ = 0.9)``; ε = 1e-8 (Dopamine_DQN_Atari.jl:42)."""

    def __init__(self, eta=0.001, rho=0.9):
        self.kind, self.eta, self.a, self.b = L.OPT_RMSPROP, float(eta), float(rho), 0.0


class ADAM:
    """Flux ``ADAM(η = 0.001, β = (0.9, 0.999))``; ε = 1e-8 (Dopamine_Rainbow_Atari.jl:41)."""

    def __init__(self, eta=0.001, beta=(0.9, 0.999)):
        self.kind, self.eta, self.a, self.b = L.OPT_ADAM, float(eta), float(beta[0]), float(beta[1])


def _abi(p):
    """logical numpy parameter -> the Julia array's bytes: Dense (out,in) column-major == row-major [in][out]."""
    p = np.asarray(p, np.float32)
    return np.ascontiguousarray(p.T) if p.ndim == 2 else np.ascontiguousarray(p)


def _from_abi(buf, shape):
    if len(shape) == 2:
        return np.ascontiguousarray(buf.reshape(shape[1], shape[0]).T)
    return buf.reshape(shape).copy()


class NeuralNetworkApproximator:
    """RLCore ``NeuralNetworkApproximator(; model, optimizer = nothing)``.

    obs_shape: (C, H, W) for frames or (n,) for vector observations.  ``max_batch`` sizes the device
    workspaces (the online net evaluates [s; s′] in one pass, so it holds 2*max_batch rows).
    precision: "fp32" (tcgen05 3xTF32, fp32-grade, parity 1e-4), "fp32_simt" (CUDA cores only), "tf32", "bf16"."""

    def __init__(self, model, optimizer=None, obs_shape=(4, 84, 84), max_batch=32, max_quantiles=0, precision="fp32"):
        self.model, self.optimizer = model, optimizer
        lib = L.lib()
        d = L.NetDesc()
        d.kind = model.kind
        d.precision = {"fp32": L.PREC_FP32, "bf16": L.PREC_BF16, "tf32": L.PREC_TF32, "fp32_simt": L.PREC_FP32_SIMT}[precision]
        if len(obs_shape) == 1:
            d.in_c, d.in_h, d.in_w = int(obs_shape[0]), 1, 1
        else:
            d.in_c, d.in_h, d.in_w = (int(x) for x in obs_shape)
        d.max_batch, d.max_quantiles = int(max_batch), int(max_quantiles)
        for name, ch in zip("abc", model.chains()):
            spec = ch.spec() if ch is not None else []
            if len(spec) > L.ZOO_MAX_LAYERS:
                raise ValueError("chain too long")
            setattr(d, "n_" + name, len(spec))
            arr = getattr(d, name)
            for i, s in enumerate(spec):
                arr[i] = L.Layer(*s)
        self.precision = precision
        self.obs_shape = tuple(obs_shape)
        self.max_batch = int(max_batch)
        self.shapes = model.shapes()
        h = C.c_void_p()
        L.check(lib.zoo_net_create(C.byref(d), C.byref(h)))
        self.h = h
        n = C.c_int32()
        L.check(lib.zoo_net_out_dim(self.h, C.byref(n)))
        self.n_out = n.value
        self.opt_h = None
        if optimizer is not None:
            oh = C.c_void_p()
            L.check(lib.zoo_opt_create(self.h, optimizer.kind, optimizer.eta, optimizer.a, optimizer.b, 1e-8, C.byref(oh)))
            self.opt_h = oh

    def __del__(self):
        try:
            lib = L.lib()
            if getattr(self, "opt_h", None):
                lib.zoo_opt_destroy(self.opt_h)
                self.opt_h = None
            if getattr(self, "h", None):
                lib.zoo_net_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # params(app) / Flux.loadparams!
    def set_params(self, params):
        assert len(params) == len(self.shapes)
        for i, (p, shp) in enumerate(zip(params, self.shapes)):
            assert tuple(np.shape(p)) == tuple(shp), (i, np.shape(p), shp)
            a = _abi(p)
            L.check(L.lib().zoo_net_set_params(self.h, i, a.ctypes.data, a.size))

    def params(self):
        out = []
        for i, shp in enumerate(self.shapes):
            buf = np.empty(int(np.prod(shp)), np.float32)
            L.check(L.lib().zoo_net_get_params(self.h, i, buf.ctypes.data, buf.size))
            out.append(_from_abi(buf, shp))
        return out

    def opt_state(self, slot):
        out = []
        for i, shp in enumerate(self.shapes):
            buf = np.empty(int(np.prod(shp)), np.float32)
            L.check(L.lib().zoo_opt_get_state(self.opt_h, slot, i, buf.ctypes.data, buf.size))
            out.append(_from_abi(buf, shp))
        return out

    def copyto(self, src):
        """``copyto!(self, src)`` -- common.jl:12-14."""
        L.check(L.lib().zoo_net_copy(self.h, src.h))

    def __call__(self, obs, tau=None):
        """``app(x)``: batch forward (acting path, dqn.jl:90-100).  obs [B, ...] uint8 | float32; IQN: tau [B, n_tau]."""
        obs = np.ascontiguousarray(obs)
        if obs.dtype != np.uint8:
            obs = obs.astype(np.float32)
        B = obs.shape[0]
        n_tau = 0
        if tau is not None:
            tau = np.ascontiguousarray(tau, np.float32)
            n_tau = tau.shape[1]
            out = np.empty((B, n_tau, self.n_out), np.float32)
        else:
            out = np.empty((B, self.n_out), np.float32)
        L.check(L.lib().zoo_net_forward(self.h, obs.ctypes.data, L.OBS_U8 if obs.dtype == np.uint8 else L.OBS_F32, B,
                                        L.ptr(tau), n_tau, out.ctypes.data))
        return out
