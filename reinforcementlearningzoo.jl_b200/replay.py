"""Host-side mirror of the device-resident replay shard (csrc/replay.cu): RLCore's ``CircularArraySARTTrajectory`` /
``CircularArrayPSARTTrajectory`` + ``NStepBatchSampler`` as used at Dopamine_DQN_Atari.jl:55-66, with the sampling call of
``update!(learner, t::AbstractTrajectory)`` (common.jl:7-26) running entirely on the device."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class ReplayShard:
    def __init__(self, capacity, frame_shape=(84, 84), stack_size=4, n_step=1, gamma=0.99, max_batch=32, prioritized=False):
        h = C.c_void_p()
        L.check(L.lib().zoo_replay_create(int(capacity), int(frame_shape[0]), int(frame_shape[1]), int(stack_size), int(n_step), float(gamma),
                                          int(max_batch), int(bool(prioritized)), C.byref(h)))
        self.h, self.capacity, self.frame_shape, self.is_prioritized = h, int(capacity), tuple(frame_shape), bool(prioritized)
        self.stack_size, self.n_step = int(stack_size), int(n_step)

    def close(self):
        if getattr(self, "h", None):
            L.lib().zoo_replay_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        n = C.c_int64()
        L.check(L.lib().zoo_replay_length(self.h, C.byref(n), None))
        return n.value

    @property
    def n_terminal(self):  # length(t[:terminal]) as read by common.jl:8
        return len(self)

    def push(self, frame, action, reward, terminal, priority=100.0):
        """One transition; ``priority`` is the learner's default_priority (prioritized_dqn.jl:57) for a prioritized shard."""
        f = np.ascontiguousarray(frame, np.uint8)
        assert f.shape == self.frame_shape
        L.check(L.lib().zoo_replay_push(self.h, f.ctypes.data, int(action), float(reward), int(bool(terminal)), float(priority)))

    def push_many(self, frames, actions, rewards, terminals, priorities=None):
        f = np.ascontiguousarray(frames, np.uint8)
        a, r, t = np.ascontiguousarray(actions, np.int32), np.ascontiguousarray(rewards, np.float32), np.ascontiguousarray(terminals, np.uint8)
        p = None if priorities is None else np.ascontiguousarray(priorities, np.float32)
        assert f.shape[1:] == self.frame_shape and len(a) == len(r) == len(t) == f.shape[0]
        L.check(L.lib().zoo_replay_push_many(self.h, f.ctypes.data, a.ctypes.data, r.ctypes.data, t.ctypes.data, L.ptr(p), f.shape[0]))

    def set_spare_draws(self, learner, spare):
        """Uniforms for the redraws of out-of-range prioritized indices (RLCore redraws them): ``spare`` [B, R] in
        the dtype of the draws, sample b consumes row b; ``None`` switches redraws off (clamp + count)."""
        if spare is None:
            L.check(L.lib().zoo_replay_set_spare_draws(self.h, learner.stream(), None, L.DRAW_F64, 1, 0))
            return
        s = np.ascontiguousarray(spare)
        if s.dtype not in (np.float32, np.float64) or s.ndim != 2:
            raise TypeError("spare draws must be a float32 or float64 [B, R] array")
        L.check(L.lib().zoo_replay_set_spare_draws(self.h, learner.stream(), s.ctypes.data, L.DRAW_F64 if s.dtype == np.float64 else L.DRAW_F32,
                                                   s.shape[0], s.shape[1]))

    def sample_batch(self, learner, u, mode=None, want_inds=False, spare=None):
        """Assemble a minibatch on the device from the uniform draws ``u`` (float32 | float64, len = batch size).
        ``spare`` [B, R]: uniforms for prioritized redraws (see set_spare_draws).  Returns (zoo_batch with device pointers, indices | None)."""
        u = np.ascontiguousarray(u)
        if u.dtype not in (np.float32, np.float64):
            raise TypeError("draws must be float32 or float64")
        if spare is not None:
            self.set_spare_draws(learner, np.asarray(spare, u.dtype))
        if mode is None:
            mode = L.SAMPLE_INDEPENDENT if self.is_prioritized else L.REPLAY_UNIFORM
        b = L.Batch()
        inds = np.empty(u.size, np.int64) if want_inds else None
        L.check(L.lib().zoo_replay_sample(self.h, learner.stream(), u.ctypes.data, L.DRAW_F64 if u.dtype == np.float64 else L.DRAW_F32, u.size,
                                          int(mode), C.byref(b), L.ptr(inds)))
        return b, inds

    def update(self, learner, u, mode=None, want_loss=True, spare=None):
        """``update!(learner, t)`` (common.jl:18-24): sample -> ``update!(learner, batch)`` -> priority write-back, all enqueued on
        the learner's stream; the only host traffic is the draws (and the loss, if wanted)."""
        lib = L.lib()
        b, _ = self.sample_batch(learner, u, mode, spare=spare)
        loss = C.c_float(0)
        L.check(lib.zoo_learner_update(learner.h, C.byref(b), C.byref(loss) if want_loss else None, None))
        if self.is_prioritized:
            p = C.c_void_p()
            L.check(lib.zoo_learner_priorities_dev(learner.h, C.byref(p)))
            L.check(lib.zoo_replay_update_priorities(self.h, learner.stream(), p))
        if want_loss:
            learner.loss = np.float32(loss.value)
        return learner.loss

    def invalid_count(self):
        n = C.c_uint32()
        L.check(L.lib().zoo_replay_invalid_count(self.h, C.byref(n)))
        return n.value

    def priority_tree_export(self):
        from .sumtree import SumTree
        t = C.c_void_p()
        L.check(L.lib().zoo_replay_sumtree(self.h, C.byref(t)))
        view = SumTree.__new__(SumTree)
        view.h, view.capacity = None, self.capacity  # borrowed handle: never destroyed by the view
        n = C.c_int64()
        L.check(L.lib().zoo_sumtree_tree_len(t, C.byref(n)))
        out = np.empty(n.value, np.float32)
        L.check(L.lib().zoo_sumtree_export(t, out.ctypes.data, out.size))
        return out
