"""SumTree oracle front-ends (TEST INFRASTRUCTURE ONLY; PARITY UNPINNED -- see sumtree_oracle.c).

``CSumTree``  : ctypes binding of ``oracle/libsumtree_oracle.so`` (plain C restatement, fast enough
                for capacity 1M / batch 4096).
``PySumTree`` : an independent pure-Python/numpy-scalar restatement used to cross-check the C one on
                small cases (reference call sites: src/algorithms/dqns/common.jl:18,22,28-37; the
                structure itself is RLCore 0.7.2 ``SumTree``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsumtree_oracle.so")


def build():
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "sumtree_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "libsumtree_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.st_create.restype = ctypes.c_void_p
        lib.st_create.argtypes = [ctypes.c_int64]
        lib.st_destroy.argtypes = [ctypes.c_void_p]
        for f in ("st_length", "st_first", "st_tree_len"):
            getattr(lib, f).restype = ctypes.c_int64
            getattr(lib, f).argtypes = [ctypes.c_void_p]
        lib.st_total.restype = ctypes.c_float
        lib.st_total.argtypes = [ctypes.c_void_p]
        lib.st_get_leaf.restype = ctypes.c_float
        lib.st_get_leaf.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        lib.st_export.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.st_set.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_float]
        lib.st_push.argtypes = [ctypes.c_void_p, ctypes.c_float]
        lib.st_update.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        lib.st_sample.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                  ctypes.c_void_p, ctypes.c_void_p]
        _lib = lib
    return _lib


class CSumTree:
    def __init__(self, capacity):
        self.lib = _load()
        self.h = self.lib.st_create(int(capacity))
        self.capacity = int(capacity)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.st_destroy(self.h)
            self.h = None

    def __len__(self):
        return self.lib.st_length(self.h)

    @property
    def first(self):
        return self.lib.st_first(self.h)

    @property
    def total(self):
        return np.float32(self.lib.st_total(self.h))

    def push(self, p):
        self.lib.st_push(self.h, float(np.float32(p)))

    def push_many(self, ps):
        for p in np.asarray(ps, np.float32):
            self.lib.st_push(self.h, float(p))

    def __getitem__(self, i):
        return np.float32(self.lib.st_get_leaf(self.h, int(i)))

    def update(self, idx, p):
        idx = np.ascontiguousarray(idx, np.int64)
        p = np.ascontiguousarray(p, np.float32)
        self.lib.st_update(self.h, idx.ctypes.data, p.ctypes.data, len(idx))

    def sample(self, u, stratified=True):
        u = np.ascontiguousarray(u)
        assert u.dtype in (np.float32, np.float64)
        n = len(u)
        idx = np.empty(n, np.int64)
        pr = np.empty(n, np.float32)
        self.lib.st_sample(self.h, u.ctypes.data, int(u.dtype == np.float64), n, int(stratified), idx.ctypes.data,
                           pr.ctypes.data)
        return idx, pr

    def export(self):
        out = np.empty(self.lib.st_tree_len(self.h), np.float32)
        self.lib.st_export(self.h, out.ctypes.data)
        return out


class PySumTree:
    """Independent restatement with numpy float32 scalars (slow; small cases only)."""

    def __init__(self, capacity):
        self.capacity = capacity
        p = 1
        while p < capacity:
            p *= 2
        self.nparents = p - 1
        self.tree = np.zeros(self.nparents + capacity + 1, np.float32)  # 1-based
        self.first = 1
        self.length = 0

    def _phys(self, i):
        k = i + self.first - 1
        return k - self.capacity if k > self.capacity else k

    def __getitem__(self, i):
        return self.tree[self.nparents + self._phys(i)]

    def set(self, i, p):
        n = self.nparents + self._phys(i)
        p = np.float32(p)
        change = np.float32(p - self.tree[n])
        self.tree[n] = p
        while n != 1:
            n //= 2
            self.tree[n] = np.float32(self.tree[n] + change)

    def update(self, idx, ps):
        for i, p in zip(idx, ps):
            self.set(int(i), p)

    def push(self, p):
        if self.length == self.capacity:
            self.first = 1 if self.first == self.capacity else self.first + 1
        else:
            self.length += 1
        self.set(self.length, p)

    @property
    def total(self):
        return self.tree[1]

    def get(self, v):
        n = 1
        L = len(self.tree) - 1
        while True:
            l = 2 * n
            if l > L:
                break
            if v <= self.tree[l]:
                n = l
            else:
                v = type(v)(v - type(v)(self.tree[l]))
                n = l + 1
        k = n - self.nparents
        logical = k - self.first + 1 if k >= self.first else k + self.capacity - self.first + 1
        return logical, self.tree[n]

    def sample(self, u, stratified=True):
        u = np.asarray(u)
        T = u.dtype.type
        n = len(u)
        idx, pr = [], []
        for i in range(n):
            if stratified:
                v = T(T(T(i) + u[i]) / T(n)) * T(self.tree[1])
            else:
                v = u[i] * T(self.tree[1])
            a, b = self.get(T(v))
            idx.append(a)
            pr.append(b)
        return np.array(idx, np.int64), np.array(pr, np.float32)
