"""Host-side mirror of RLCore's ``SumTree`` (the object behind ``t[:priority]`` at common.jl:18,22,36) on top of the
device tree of libzoo_sm100.  Indices are 1-based logical indices exactly as the Julia code sees them."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class SumTree:
    def __init__(self, capacity):
        h = C.c_void_p()
        L.check(L.lib().zoo_sumtree_create(int(capacity), C.byref(h)))
        self.h, self.capacity = h, int(capacity)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                L.lib().zoo_sumtree_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def __len__(self):
        n, f = C.c_int64(), C.c_int64()
        L.check(L.lib().zoo_sumtree_length(self.h, C.byref(n), C.byref(f)))
        return n.value

    @property
    def first(self):
        n, f = C.c_int64(), C.c_int64()
        L.check(L.lib().zoo_sumtree_length(self.h, C.byref(n), C.byref(f)))
        return f.value

    def push(self, p):
        """``push!(t, p)`` -- common.jl:36."""
        L.check(L.lib().zoo_sumtree_push(self.h, float(np.float32(p))))

    def push_many(self, ps):
        ps = np.ascontiguousarray(ps, np.float32)
        L.check(L.lib().zoo_sumtree_push_many(self.h, ps.ctypes.data, ps.size))

    @property
    def total(self):
        v = C.c_float()
        L.check(L.lib().zoo_sumtree_total(self.h, C.byref(v)))
        return np.float32(v.value)

    def __getitem__(self, idx):
        scalar = np.isscalar(idx)
        idx = np.ascontiguousarray(np.atleast_1d(idx), np.int64)
        out = np.empty(idx.size, np.float32)
        L.check(L.lib().zoo_sumtree_get(self.h, idx.ctypes.data, idx.size, out.ctypes.data))
        return out[0] if scalar else out

    def update(self, idx, p):
        """``t[inds] .= p`` -- common.jl:22: sequential fp32 semantics, duplicates allowed, bit-exact."""
        idx = np.ascontiguousarray(idx, np.int64)
        p = np.ascontiguousarray(p, np.float32)
        assert idx.shape == p.shape
        L.check(L.lib().zoo_sumtree_update(self.h, idx.ctypes.data, p.ctypes.data, idx.size))

    __setitem__ = update

    def sample(self, u, stratified=True):
        """``sample(rng, t, …)`` -- common.jl:18 with the uniform draws ``u`` (float32 | float64, in [0,1)) supplied by
        the caller: v = (i-1+u_i)/n*total (stratified) or u_i*total (independent).  Returns (1-based indices, priorities)."""
        u = np.ascontiguousarray(u)
        if u.dtype not in (np.float32, np.float64):
            raise TypeError("draws must be float32 or float64")
        idx = np.empty(u.size, np.int64)
        pr = np.empty(u.size, np.float32)
        L.check(L.lib().zoo_sumtree_sample(self.h, u.ctypes.data, L.DRAW_F64 if u.dtype == np.float64 else L.DRAW_F32, u.size,
                                           L.SAMPLE_STRATIFIED if stratified else L.SAMPLE_INDEPENDENT, idx.ctypes.data, pr.ctypes.data))
        return idx, pr

    def export(self):
        n = C.c_int64()
        L.check(L.lib().zoo_sumtree_tree_len(self.h, C.byref(n)))
        out = np.empty(n.value, np.float32)
        L.check(L.lib().zoo_sumtree_export(self.h, out.ctypes.data, out.size))
        return out

    def load(self, tree, length, first=1):
        tree = np.ascontiguousarray(tree, np.float32)
        L.check(L.lib().zoo_sumtree_import(self.h, tree.ctypes.data, tree.size, int(length), int(first)))
