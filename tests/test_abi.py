"""CPU: the C-ABI shared library loads (no GPU needed for dlopen) and exports every symbol include/zoo_sm100.h declares,
the ctypes mirror covers all of them, the struct layouts match the header, and -- without a GPU -- compute entry
points refuse loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "zoo_sm100.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"ZOO_API\s+[\w\s\*]+?\b(zoo_\w+)\s*\(", src)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    assert len(syms) >= 50
    for must in ("zoo_learner_update", "zoo_sumtree_sample", "zoo_sumtree_update", "zoo_net_copy", "zoo_opt_create", "zoo_dp_init"):
        assert must in syms


def test_library_exports_every_declared_symbol(zoo):
    assert os.path.exists(zoo.LIB_PATH), "libzoo_sm100.so missing: run __graft_entry__.build()"
    out = subprocess.check_output(["nm", "-D", "--defined-only", zoo.LIB_PATH], text=True)
    exported = set(l.split()[-1] for l in out.splitlines() if " T " in l)
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing
    # nothing but the zoo_* C ABI leaks out (-fvisibility=hidden)
    assert all(s.startswith("zoo_") or s.startswith("_") for s in exported), sorted(exported)[:20]


def test_ctypes_mirror_covers_the_header(zoo):
    from importlib import import_module
    L = import_module("rlzoo_b200._lib")
    bound = set(L.SIGNATURES) | set(L._OTHER_RESTYPE)
    assert sorted(bound) == declared_symbols()
    lib = zoo.lib()
    assert lib.zoo_version() >= 1


def test_struct_layouts_match_the_header(zoo, tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirror."""
    from importlib import import_module
    L = import_module("rlzoo_b200._lib")
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "zoo_sm100.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(zoo_layer), sizeof(zoo_net_desc), sizeof(zoo_learner_cfg),"
        " sizeof(zoo_batch), offsetof(zoo_net_desc, b), offsetof(zoo_learner_cfg, support), offsetof(zoo_learner_cfg, seed),"
        " offsetof(zoo_batch, on_device));return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    want = [C.sizeof(L.Layer), C.sizeof(L.NetDesc), C.sizeof(L.LearnerCfg), C.sizeof(L.Batch), L.NetDesc.b.offset,
            L.LearnerCfg.support.offset, L.LearnerCfg.seed.offset, L.Batch.on_device.offset]
    assert got == want


def test_no_cpu_fallback(zoo):
    """Without an sm_100 device every compute entry point returns an error; nothing computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the refusal path is for CPU-only hosts")
    with pytest.raises(zoo.ZooError):
        zoo.SumTree(8)
    with pytest.raises(zoo.ZooError):
        zoo.NeuralNetworkApproximator(zoo.Chain(zoo.Dense(4, 2)), None, (4,), 4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "reinforcementlearningzoo.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("from_oracle_net", "").replace("oracle-style", ""), os.path.join(dp, f)
