"""Host-side mirror of the reference learners (same names, keywords and return values):
``BasicDQNLearner`` (basic_dqn.jl:22-90), ``DQNLearner`` (dqn.jl:3-145), ``PrioritizedDQNLearner``
(prioritized_dqn.jl:24-151), ``RainbowLearner`` (rainbow.jl:28-196), ``IQNLearner`` (iqn.jl:61-237) and the
trajectory-level ``update!`` (common.jl:7-26).  ``update(learner, batch)`` is ``RLBase.update!(learner, batch)``:
it returns the new priorities (``Vector{Float32}``) for PER batches, else ``None``, and sets ``learner.loss``.
Every learner only marshals arguments into ``zoo_learner_update``; there is no Python arithmetic on this path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

SARTS = ("state", "action", "reward", "terminal", "next_state")


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dt)


class _Learner:
    kind = None

    def _create(self, cfg, dp=None):
        lib = L.lib()
        h = C.c_void_p()
        tgt = self.target_approximator.h if self.target_approximator is not None else None
        L.check(lib.zoo_learner_create(C.byref(cfg), self.approximator.h, tgt, self.approximator.opt_h,
                                       dp.h if dp is not None else None, C.byref(h)))
        self.h = h
        self.loss = np.float32(0)
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            L.lib().zoo_learner_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _batch(self, batch, tau=None, tau_prime=None):
        """dict of host numpy arrays, or of device tensors (anything with ``data_ptr()``, e.g. torch CUDA tensors on the
        library's device with the ABI dtypes: u8|f32 frames, int32 actions, f32 rewards, u8 terminals) -> zoo_batch."""
        b = L.Batch()
        if hasattr(batch["state"], "data_ptr"):
            s = batch["state"]
            keep = [s, batch["next_state"], batch["action"], batch["reward"], batch["terminal"], batch.get("priority"),
                    batch.get("next_legal_actions_mask"), tau, tau_prime]
            want = [None, None, 4, 4, 1, 4, 1, 4, 4]
            for x, w in zip(keep, want):
                if x is not None and not (x.is_cuda and x.is_contiguous() and (w is None or x.element_size() == w)):
                    raise ValueError("device batches must be contiguous CUDA tensors with the ABI dtypes")
            b.B = int(s.shape[0])
            b.obs_dtype = L.OBS_U8 if s.element_size() == 1 else L.OBS_F32
            (b.state, b.next_state, b.action, b.reward, b.terminal, b.priority, b.next_mask, b.tau, b.tau_prime) = (
                None if x is None else x.data_ptr() for x in keep)
            b.on_device = 1
            return b, keep
        s = np.ascontiguousarray(batch["state"])
        s2 = np.ascontiguousarray(batch["next_state"])
        if s.dtype != np.uint8:
            s, s2 = s.astype(np.float32, copy=False), s2.astype(np.float32, copy=False)
        keep = [s, s2, _c(batch["action"], np.int32), _c(batch["reward"], np.float32), _c(batch["terminal"], np.uint8),
                _c(batch.get("priority"), np.float32), _c(batch.get("next_legal_actions_mask"), np.uint8),
                _c(tau, np.float32), _c(tau_prime, np.float32)]
        b.B = s.shape[0]
        b.obs_dtype = L.OBS_U8 if s.dtype == np.uint8 else L.OBS_F32
        (b.state, b.next_state, b.action, b.reward, b.terminal, b.priority, b.next_mask, b.tau, b.tau_prime) = (L.ptr(x) for x in keep)
        b.on_device = 0
        return b, keep

    def update(self, batch, tau=None, tau_prime=None, want_loss=True, apply=True, want_priorities=True):
        """``RLBase.update!(learner, batch::NamedTuple)``.  0-based actions.  IQN: tau/tau_prime [B,N]/[B,N′]
        override the on-device RNG (used for parity)."""
        lib = L.lib()
        b, keep = self._batch(batch, tau, tau_prime)
        if b.on_device:
            # the update may return before the device has read the batch (want_loss=False): keep the tensors alive until the next
            # update or sync, and make the learner's stream wait for the stream that produced them (no host block)
            self._keep = keep
            try:
                import torch
                cur = torch.cuda.current_stream(keep[0].device)
                if cur.cuda_stream != self.stream():
                    ev = torch.cuda.Event()
                    ev.record(cur)
                    torch.cuda.ExternalStream(self.stream(), device=keep[0].device).wait_event(ev)
            except ImportError:
                pass
        per = batch.get("priority") is not None
        loss = C.c_float(0)
        prio = np.empty(b.B, np.float32) if (per and want_priorities) else None
        fn = lib.zoo_learner_update if apply else lib.zoo_learner_compute_grads
        L.check(fn(self.h, C.byref(b), C.byref(loss) if want_loss else None, L.ptr(prio)))
        if want_loss:
            self.loss = np.float32(loss.value)
        return prio

    def submit(self, batch, tau=None, tau_prime=None, want_priorities=False):
        """``update!`` without waiting for its loss: the result is fetched by a later :meth:`collect` (first in, first out; at most
        four updates in flight).  Host batches travel on a copy stream into one of two staging slots, so the H2D copy of the next
        batch overlaps this update's kernels (``zoo_learner_submit``).  The batch's arrays are kept alive until the collect."""
        b, keep = self._batch(batch, tau, tau_prime)
        if not hasattr(self, "_inflight"):
            self._inflight = []
        L.check(L.lib().zoo_learner_submit(self.h, C.byref(b), int(bool(want_priorities))))
        self._inflight.append((keep, bool(want_priorities), b.B))

    def collect(self):
        """Result of the oldest submitted update: sets ``self.loss``; returns the new priorities if they were asked for."""
        keep, want_prio, B = self._inflight.pop(0)
        loss = C.c_float(0)
        prio = np.empty(B, np.float32) if want_prio else None
        L.check(L.lib().zoo_learner_collect(self.h, C.byref(loss), L.ptr(prio)))
        self.loss = np.float32(loss.value)
        return prio

    def compute_grads(self, batch, tau=None, tau_prime=None):
        """The gradient half of update! (parity hook): returns (loss, grads list, priorities|None)."""
        prio = self.update(batch, tau, tau_prime, apply=False)
        grads = []
        for i, shp in enumerate(self.approximator.shapes):
            buf = np.empty(int(np.prod(shp)), np.float32)
            L.check(L.lib().zoo_learner_get_grad(self.h, i, buf.ctypes.data, buf.size))
            grads.append(buf.reshape(shp[1], shp[0]).T.copy() if len(shp) == 2 else buf.reshape(shp).copy())
        return self.loss, grads, prio

    def apply(self):
        L.check(L.lib().zoo_learner_apply(self.h))

    def per_sample_loss(self):
        out = np.empty(self.batch_size, np.float32)
        L.check(L.lib().zoo_learner_get_per_sample_loss(self.h, out.ctypes.data, out.size))
        return out

    def use_graph(self, enable=True):
        L.check(L.lib().zoo_learner_use_graph(self.h, int(enable)))

    def sync(self):
        L.check(L.lib().zoo_learner_sync(self.h))
        self._keep = None

    def stream(self):
        s = C.c_void_p()
        L.check(L.lib().zoo_learner_stream(self.h, C.byref(s)))
        return s.value

    def profile(self, batch, delay_us=3000, **kw):
        """One eager (non-graph) update under the built-in launch profiler: [(label, ms), ...] per kernel launch."""
        lib = L.lib()
        b, keep = self._batch(batch, kw.get("tau"), kw.get("tau_prime"))
        L.check(lib.zoo_learner_use_graph(self.h, 0))
        L.check(lib.zoo_prof_begin(self.stream(), int(delay_us)))
        try:
            L.check(lib.zoo_learner_update(self.h, C.byref(b), None, None))
        finally:
            n = C.c_int32()
            L.check(lib.zoo_prof_end(C.byref(n)))
        out = []
        name, ms = C.create_string_buffer(256), C.c_float()
        for i in range(n.value):
            L.check(lib.zoo_prof_get(i, name, 256, C.byref(ms)))
            out.append((name.value.decode(), ms.value))
        return out

    def last_launches(self):
        n = C.c_int32()
        L.check(L.lib().zoo_learner_last_launches(self.h, C.byref(n)))
        return n.value

    # ``(learner)(env)``: Q-values for acting -- dqn.jl:90-100
    def __call__(self, obs):
        return self.approximator(np.asarray(obs)[None])[0]


def _cfg(kind, batch_size, n_actions, gamma=0.99, n=1, double=1, beta=0.5):
    c = L.LearnerCfg()
    c.kind, c.batch_size, c.n_actions = kind, int(batch_size), int(n_actions)
    c.gamma, c.update_horizon, c.double_dqn, c.beta_priority = float(gamma), int(n), int(double), float(beta)
    c.kappa = 1.0
    return c


class BasicDQNLearner(_Learner):
    """basic_dqn.jl:43-60.  loss_func is Flux ``huber_loss`` (the only one the reference experiments use)."""

    def __init__(self, approximator, gamma=0.99, batch_size=32, min_replay_history=32, dp=None):
        self.approximator, self.target_approximator = approximator, None
        self.gamma, self.batch_size, self.min_replay_history = gamma, batch_size, min_replay_history
        self._create(_cfg(L.LEARNER_BASIC_DQN, batch_size, approximator.n_out, gamma), dp)


class DQNLearner(_Learner):
    """dqn.jl:44-80 (``copyto!(approximator, target_approximator)`` force-sync at dqn.jl:60 included)."""

    def __init__(self, approximator, target_approximator, gamma=0.99, batch_size=32, update_horizon=1, min_replay_history=32,
                 update_freq=1, target_update_freq=100, update_step=0, is_enable_double_DQN=True, stack_size=None, dp=None):
        approximator.copyto(target_approximator)
        self.approximator, self.target_approximator = approximator, target_approximator
        self.gamma, self.batch_size, self.update_horizon = gamma, batch_size, update_horizon
        self.min_replay_history, self.update_freq, self.target_update_freq, self.update_step = min_replay_history, update_freq, target_update_freq, update_step
        self.is_enable_double_DQN = is_enable_double_DQN
        self._create(_cfg(L.LEARNER_DQN, batch_size, approximator.n_out, gamma, update_horizon, int(is_enable_double_DQN)), dp)


class PrioritizedDQNLearner(_Learner):
    """prioritized_dqn.jl:45-83.  loss_func is ``huber_loss(ŷ, y; agg = identity)`` (JuliaRL_PrioritizedDQN_CartPole.jl:39)."""

    def __init__(self, approximator, target_approximator, gamma=0.99, batch_size=32, update_horizon=1, min_replay_history=32,
                 update_freq=1, target_update_freq=100, update_step=0, default_priority=100.0, beta_priority=0.5, stack_size=None, dp=None):
        approximator.copyto(target_approximator)
        self.approximator, self.target_approximator = approximator, target_approximator
        self.gamma, self.batch_size, self.update_horizon = gamma, batch_size, update_horizon
        self.min_replay_history, self.update_freq, self.target_update_freq, self.update_step = min_replay_history, update_freq, target_update_freq, update_step
        self.default_priority, self.beta_priority = np.float32(default_priority), np.float32(beta_priority)
        self._create(_cfg(L.LEARNER_PRIORITIZED_DQN, batch_size, approximator.n_out, gamma, update_horizon, 0, beta_priority), dp)


class RainbowLearner(_Learner):
    """rainbow.jl:64-116.  Default ``support = range(-Vmax, Vmax, length = n_atoms)`` ignores Vmin (q5, rainbow.jl:72);
    loss_func is ``logitcrossentropy(ŷ, y; agg = identity)`` (Dopamine_Rainbow_Atari.jl:54)."""

    def __init__(self, approximator, target_approximator, n_actions, Vmax, Vmin, n_atoms=51, support=None, delta_z=None, gamma=0.99,
                 batch_size=32, update_horizon=1, min_replay_history=32, update_freq=1, target_update_freq=500, update_step=0,
                 default_priority=100.0, beta_priority=0.5, stack_size=4, dp=None):
        if default_priority < 1.0:
            raise ValueError("default value must be >= 1.0f0")  # rainbow.jl:87
        approximator.copyto(target_approximator)
        self.approximator, self.target_approximator = approximator, target_approximator
        if support is None:
            support = np.linspace(np.float32(-Vmax), np.float32(Vmax), n_atoms).astype(np.float32)
        self.support = np.ascontiguousarray(support, np.float32)
        self.delta_z = np.float32(self.support[1] - self.support[0]) if delta_z is None else np.float32(delta_z)
        self.n_actions, self.n_atoms, self.Vmax, self.Vmin = n_actions, n_atoms, Vmax, Vmin
        self.gamma, self.batch_size, self.update_horizon = gamma, batch_size, update_horizon
        self.min_replay_history, self.update_freq, self.target_update_freq, self.update_step = min_replay_history, update_freq, target_update_freq, update_step
        self.default_priority, self.beta_priority = np.float32(default_priority), np.float32(beta_priority)
        c = _cfg(L.LEARNER_RAINBOW, batch_size, n_actions, gamma, update_horizon, 0, beta_priority)
        c.n_atoms, c.v_min, c.v_max, c.delta_z = int(n_atoms), float(Vmin), float(Vmax), float(self.delta_z)
        c.support = self.support.ctypes.data
        self._create(c, dp)

    def __call__(self, obs):  # rainbow.jl:118-124: expected value under softmax(logits)
        logits = self.approximator(np.asarray(obs)[None])[0].reshape(self.n_actions, self.n_atoms)
        e = np.exp(logits - logits.max(axis=1, keepdims=True))
        return (self.support * (e / e.sum(axis=1, keepdims=True))).sum(axis=1)


class IQNLearner(_Learner):
    """iqn.jl:90-146.  τ, τ′ come from the on-device Philox stream (``device_rng``) unless passed to ``update``."""

    def __init__(self, approximator, target_approximator, n_actions, kappa=1.0, N=32, N_prime=32, N_em=64, K=32, gamma=0.99,
                 batch_size=32, update_horizon=1, min_replay_history=20000, update_freq=4, target_update_freq=8000, update_step=0,
                 default_priority=100.0, beta_priority=0.5, device_seed=0, stack_size=4, dp=None):
        approximator.copyto(target_approximator)
        self.approximator, self.target_approximator = approximator, target_approximator
        self.n_actions, self.kappa, self.N, self.N_prime, self.N_em, self.K = n_actions, kappa, N, N_prime, N_em, K
        self.gamma, self.batch_size, self.update_horizon = gamma, batch_size, update_horizon
        self.min_replay_history, self.update_freq, self.target_update_freq, self.update_step = min_replay_history, update_freq, target_update_freq, update_step
        self.default_priority, self.beta_priority = np.float32(default_priority), np.float32(beta_priority)
        c = _cfg(L.LEARNER_IQN, batch_size, n_actions, gamma, update_horizon, 0, beta_priority)
        c.kappa, c.N, c.N_prime, c.N_em, c.seed = float(kappa), int(N), int(N_prime), int(N_em), int(device_seed)
        self._create(c, dp)
        self._rng = np.random.default_rng(device_seed)

    def __call__(self, obs):  # iqn.jl:148-155: mean over K quantiles
        tau = self._rng.random((1, self.K), dtype=np.float32)
        return self.approximator(np.asarray(obs)[None], tau)[0].mean(axis=0)


class QRDQNLearner(_Learner):
    """qr_dqn.jl:53-93.  The approximator outputs ``n_quantile * n_actions`` values per state (quantile index fastest);
    loss_func is ``quantile_huber_loss`` (qr_dqn.jl:3-14, κ = 1)."""

    def __init__(self, approximator, target_approximator, n_actions, gamma=0.99, batch_size=32, update_horizon=1, min_replay_history=32,
                 update_freq=1, n_quantile=1, target_update_freq=100, update_step=0, kappa=1.0, stack_size=None, dp=None):
        approximator.copyto(target_approximator)
        self.approximator, self.target_approximator = approximator, target_approximator
        self.n_actions, self.n_quantile = n_actions, n_quantile
        self.gamma, self.batch_size, self.update_horizon = gamma, batch_size, update_horizon
        self.min_replay_history, self.update_freq, self.target_update_freq, self.update_step = min_replay_history, update_freq, target_update_freq, update_step
        c = _cfg(L.LEARNER_QR_DQN, batch_size, n_actions, gamma, update_horizon, 0)
        c.kappa, c.N = float(kappa), int(n_quantile)
        self._create(c, dp)

    def __call__(self, obs):  # qr_dqn.jl:98-103: mean over the quantiles
        return self.approximator(np.asarray(obs)[None])[0].reshape(self.n_actions, self.n_quantile).mean(axis=1)


class REMDQNLearner(_Learner):
    """rem_dqn.jl:50-92.  The approximator outputs ``n_actions * ensemble_num`` values per state (action index fastest);
    loss_func is Flux ``huber_loss`` (JuliaRL_REMDQN_CartPole.jl:40).  The convex combination is drawn on the host, as the
    reference does (rem_dqn.jl:110-116), from ``rng``; pass ``convex_polygon`` to ``update`` to fix it (parity)."""

    def __init__(self, approximator, target_approximator, n_actions, gamma=0.99, batch_size=32, update_horizon=1, min_replay_history=32,
                 update_freq=1, ensemble_num=1, ensemble_method="rand", target_update_freq=100, update_step=0, rng=None, stack_size=None,
                 dp=None):
        approximator.copyto(target_approximator)
        self.approximator, self.target_approximator = approximator, target_approximator
        self.n_actions, self.ensemble_num, self.ensemble_method = n_actions, ensemble_num, ensemble_method
        self.gamma, self.batch_size, self.update_horizon = gamma, batch_size, update_horizon
        self.min_replay_history, self.update_freq, self.target_update_freq, self.update_step = min_replay_history, update_freq, target_update_freq, update_step
        self.rng = rng if rng is not None else np.random.default_rng()
        c = _cfg(L.LEARNER_REM_DQN, batch_size, n_actions, gamma, update_horizon, 0)
        c.N = int(ensemble_num)
        self._create(c, dp)

    def _polygon(self):
        w = self.rng.random(self.ensemble_num, dtype=np.float32) if self.ensemble_method == "rand" else np.ones(self.ensemble_num, np.float32)
        return (w / w.sum(dtype=np.float32)).astype(np.float32)

    def update(self, batch, convex_polygon=None, **kw):
        return super().update(batch, tau=self._polygon() if convex_polygon is None else np.asarray(convex_polygon, np.float32), **kw)

    def compute_grads(self, batch, convex_polygon=None):
        prio = self.update(batch, convex_polygon=convex_polygon, apply=False)
        grads = []
        for i, shp in enumerate(self.approximator.shapes):
            buf = np.empty(int(np.prod(shp)), np.float32)
            L.check(L.lib().zoo_learner_get_grad(self.h, i, buf.ctypes.data, buf.size))
            grads.append(buf.reshape(shp[1], shp[0]).T.copy() if len(shp) == 2 else buf.reshape(shp).copy())
        return self.loss, grads, prio

    def __call__(self, obs):  # rem_dqn.jl:94-98: mean over the ensemble
        return self.approximator(np.asarray(obs)[None])[0].reshape(self.ensemble_num, self.n_actions).mean(axis=0)


def update(learner, x, **kw):
    """``RLBase.update!``: dispatches on the second argument like the reference's methods --
    a batch (dict with the SARTS keys) or a trajectory (anything with ``sample`` / ``set_priority``)."""
    if isinstance(x, dict):
        return learner.update(x, **kw)
    return update_trajectory(learner, x)


def update_trajectory(learner, t):
    """``update!(learner, t::AbstractTrajectory)`` -- common.jl:7-26 (BasicDQN: basic_dqn.jl:62-67).
    ``t`` must provide ``n_terminal`` (length(t[:terminal])), ``sample(learner) -> (inds, batch)`` and, for a
    prioritized trajectory, ``is_prioritized`` and ``set_priority(inds, priorities)``."""
    if isinstance(learner, BasicDQNLearner):
        if len(t) >= learner.min_replay_history:
            _, batch = t.sample(learner)
            learner.update(batch)
        return None
    if t.n_terminal - learner.update_horizon <= learner.min_replay_history:  # common.jl:8
        return None
    learner.update_step += 1  # common.jl:10
    if learner.update_step % learner.target_update_freq == 0:  # common.jl:12-14
        learner.target_approximator.copyto(learner.approximator)
    if learner.update_step % learner.update_freq != 0:  # common.jl:16
        return None
    inds, batch = t.sample(learner)  # common.jl:18
    if getattr(t, "is_prioritized", False):
        priorities = learner.update(batch)  # common.jl:21
        t.set_priority(inds, priorities)  # common.jl:22
        return priorities
    learner.update(batch)  # common.jl:24
    return None
