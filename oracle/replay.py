"""CPU restatement of the replay batch assembly (TEST INFRASTRUCTURE ONLY; PARITY UNPINNED).

Reference: the call ``inds, batch = sample(rng, t, learner.sampler)`` at src/algorithms/dqns/common.jl:18 with the trajectory /
sampler built at src/experiments/atari/Dopamine_DQN_Atari.jl:55-66 (``CircularArraySARTTrajectory(capacity, state = Matrix{Float32}
=> (84,84))``, ``NStepBatchSampler{SARTS}(γ, n, stack_size = 4, batch_size)``).  Their implementation is RLCore 0.7.2 /
CircularArrayBuffers 0.1, not vendored under /root/reference, so the semantics below are the ones restated in SURVEY.md
Appendix A.2 [RECALLED]:
    s  = frames[i−stack+1 : i]   s′ = frames[i+n−stack+1 : i+n]   a = action[i]
    m  = first k ∈ 1..n with terminal[i+k−1];  t = (m exists);  r = Σ_{k=1..(m or n)} γ^{k−1}·reward[i+k−1], fp32, back to front
    valid i: stack ≤ i ≤ length − n;   uniform draw u ↦ i = lo + ⌊u·(hi − lo + 1)⌋
One deliberate simplification, stated: every transition owns exactly one frame (its state); RLCore keeps one extra trailing state.
"""
from __future__ import annotations

import numpy as np

from .sumtree import CSumTree

F32 = np.float32


class ReplayOracle:
    def __init__(self, capacity, frame_shape=(84, 84), stack_size=4, n_step=1, gamma=0.99, prioritized=False):
        self.C, self.stack, self.n, self.gamma = int(capacity), int(stack_size), int(n_step), F32(gamma)
        self.frames = np.zeros((self.C,) + tuple(frame_shape), np.uint8)
        self.action = np.zeros(self.C, np.int32)
        self.reward = np.zeros(self.C, F32)
        self.terminal = np.zeros(self.C, np.uint8)
        self.first, self.length = 1, 0
        self.tree = CSumTree(self.C) if prioritized else None

    def _slot(self, i):
        return (self.first - 1 + i - 1) % self.C

    def push(self, frame, action, reward, terminal, priority=100.0):
        if self.length == self.C:
            self.first = 1 if self.first == self.C else self.first + 1
        else:
            self.length += 1
        s = self._slot(self.length)
        self.frames[s], self.action[s], self.reward[s], self.terminal[s] = frame, action, F32(reward), int(bool(terminal))
        if self.tree is not None:
            self.tree.push(F32(priority))

    def valid_range(self):
        return self.stack, self.length - self.n

    def indices(self, u, mode="uniform", spare=None):
        """uniform: i = lo + floor(u * n_valid) in the dtype of the draw; 'independent' / 'stratified': SumTree.
        A prioritized index outside the valid range is REDRAWN (RLCore's sampler loops until the index is valid):
        sample b consumes spare[b, 0], spare[b, 1], ... as independent single draws get(u * total); when the spares run out (or
        none are given) it is clamped to the range edge and takes that transition's priority.
        Returns (indices, priorities | None, number clamped)."""
        lo, hi = self.valid_range()
        u = np.asarray(u)
        if mode == "uniform":
            T = u.dtype.type
            i = lo + (u * T(hi - lo + 1)).astype(np.int64)
            return np.minimum(i, hi), None, 0
        idx, pr = self.tree.sample(u, stratified=(mode == "stratified"))
        idx, pr = np.array(idx, np.int64), np.array(pr, F32)
        nbad = 0
        R = 0 if spare is None else np.asarray(spare).shape[1]
        for b in range(len(idx)):
            r = 0
            while (idx[b] < lo or idx[b] > hi) and r < R:
                i1, p1 = self.tree.sample(np.asarray(spare)[b, r:r + 1].astype(u.dtype), stratified=False)
                idx[b], pr[b] = i1[0], p1[0]
                r += 1
            if idx[b] < lo or idx[b] > hi:
                nbad += 1
                idx[b] = lo if idx[b] < lo else hi
                pr[b] = self.tree[int(idx[b])]
        return idx, pr, nbad

    def fetch(self, inds):
        B = len(inds)
        s = np.zeros((B, self.stack) + self.frames.shape[1:], np.uint8)
        s2 = np.zeros_like(s)
        a, r, t = np.zeros(B, np.int32), np.zeros(B, F32), np.zeros(B, np.uint8)
        for b, i in enumerate(int(x) for x in inds):
            for c in range(self.stack):
                s[b, c] = self.frames[self._slot(i - self.stack + 1 + c)]
                s2[b, c] = self.frames[self._slot(i + self.n - self.stack + 1 + c)]
            a[b] = self.action[self._slot(i)]
            m = self.n
            for k in range(1, self.n + 1):
                if self.terminal[self._slot(i + k - 1)]:
                    m, t[b] = k, 1
                    break
            acc = F32(0)
            for k in range(m, 0, -1):
                acc = F32(self.reward[self._slot(i + k - 1)] + F32(self.gamma * acc))
            r[b] = acc
        return {"state": s, "next_state": s2, "action": a, "reward": r, "terminal": t}
