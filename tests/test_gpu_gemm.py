"""GPU: the contraction engines on their own (zoo_test_gemm through the C ABI) against a float64 numpy product, over the
operand majors, split-K factors and the layer shapes of the Dopamine nets; plus the opt-in kernels (implicit-GEMM conv,
programmatic dependent launch) re-running the learner parity suite in a subprocess with their switch on.
Tolerances: CUDA-core fp32 and tcgen05 3xTF32 are fp32-grade (<= 5e-6 normwise here, 1e-4 end to end in the learner tests);
tcgen05 bf16 is compared with the product of the bf16-rounded inputs; single-pass tf32 is held to 2e-3."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bf16_round(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def run(zoo, engine, M, N, K, ak, bk, split):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    Am = np.ascontiguousarray(A if ak else A.T)
    Bm = np.ascontiguousarray(B if bk else B.T)
    D = np.zeros((M, N), np.float32)
    s = zoo.lib().zoo_test_gemm(engine, M, N, K, Am.ctypes.data, ak, Bm.ctypes.data, bk, D.ctypes.data, split)
    assert s == 0, zoo.lib().zoo_last_error()
    if engine in (1, 4):
        ref = bf16_round(A).astype(np.float64) @ bf16_round(B).astype(np.float64).T
    else:
        ref = A.astype(np.float64) @ B.astype(np.float64).T
    return float(np.abs(D - ref).max() / np.abs(ref).max())


SHAPES = [(128, 128, 64), (256, 32, 256), (300, 200, 136), (3872, 64, 512), (64, 512, 7744), (512, 7744, 32), (64, 576, 3872),
          (32, 256, 1000), (7744, 32, 512), (512, 64, 7744)]
TOL = {0: 5e-6, 1: 1e-4, 2: 2e-3, 3: 5e-6}  # engine 1 vs bf16-rounded inputs: what remains is the truncating TMEM accumulation


@pytest.mark.parametrize("engine", [0, 1, 2, 3])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_engines_all_majors(zoo, engine, M, N, K):
    for ak in (1, 0):
        for bk in (1, 0):
            if engine == 1 and (((M if not ak else K) % 8) or ((N if not bk else K) % 8)):
                continue  # bf16 TMA needs 16-byte rows
            if engine in (2, 3) and (((M if not ak else K) % 4) or ((N if not bk else K) % 4)):
                continue
            err = run(zoo, engine, M, N, K, ak, bk, 0)
            assert err <= TOL[engine], (engine, M, N, K, ak, bk, err)


@pytest.mark.parametrize("M,N,K", [(16384, 1024, 256), (16400, 1000, 320), (512, 7744, 4096), (20000, 264, 264)])
def test_persistent_engine_all_majors(zoo, M, N, K):
    """gemm_persist.cu (128 x 256 tiles, one CTA per SM, double-buffered TMEM accumulators, TMA-store epilogue) against the float64
    product of the bf16-rounded inputs: every operand major, ragged M / N / K tails, the long-K few-tiles case."""
    for ak in (1, 0):
        for bk in (1, 0):
            if ((M if not ak else K) % 8) or ((N if not bk else K) % 8):
                continue
            err = run(zoo, 4, M, N, K, ak, bk, 0)
            assert err <= TOL[1], (M, N, K, ak, bk, err)


@pytest.mark.parametrize("env,val", [("ZOO_PAIR_GEMM", "0"), ("ZOO_PAIR_MIN_K", "64")])
def test_persistent_engine_pair_modes(env, val):
    """The persistent engine picks CTA pairs (cta_group::2, 256 x 256 tiles, each CTA stages half of B) for K >= 1024 and one CTA per
    128 x 256 tile below that.  Both forced everywhere: the stand-alone GEMM cases above, and the IQN 512 x 64 bf16 update against the
    twin (rowmul / relu'-mask epilogues of the phi forward and the fc1 data gradient on the pair kernel)."""
    e = dict(os.environ)
    e[env] = val
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_gemm.py"), os.path.join(ROOT, "tests", "test_gpu_full_size.py"),
                        "-m", "gpu", "-q", "-x", "-k", "persistent_engine_all_majors or (pinned and iqn_atari_b512 and bf16)"],
                       env=e, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("M,N", [(1024, 6), (1025, 6), (2048, 6), (4096, 13), (2048, 8), (2048, 16)])
def test_narrow_heads_many_rows(zoo, M, N):
    # fc2 of an IQN batch (batch x quantiles rows, n_actions columns): above 1024 rows the CUDA-core engine switches to its
    # warp-per-row kernel, whose lanes N..NMAX-1 once stored into the next row's first columns
    assert run(zoo, 0, M, N, 512, 1, 1, 0) <= TOL[0]


@pytest.mark.parametrize("engine", [1, 3])
@pytest.mark.parametrize("split", [1, 2, 5, 16])
def test_split_k_is_consistent(zoo, engine, split):
    # in-kernel split-K finish (last CTA per tile applies the epilogue, workspace left zeroed): any split gives the same product
    for (M, N, K) in [(64, 512, 7744), (200, 96, 2048)]:
        assert run(zoo, engine, M, N, K, 1, 1, split) <= TOL[engine]
        assert run(zoo, engine, M, N, K, 1, 1, split) <= TOL[engine]  # second launch: the workspace really was left zeroed


@pytest.mark.parametrize("env,val", [("ZOO_IMPLICIT_CONV", "1"), ("ZOO_TMA_DGRAD", "0"), ("ZOO_TMA_WGRAD", "0"), ("ZOO_PDL", "0"), ("ZOO_TMA_CONV", "0"),
                                     ("ZOO_TMA_STORE", "0"), ("ZOO_PERSIST_GEMM", "0"), ("ZOO_DQN_TAIL", "1"), ("ZOO_DQN_TAIL", "2"), ("ZOO_NARROW", "0"), ("ZOO_IQN_FUSE", "0"), ("ZOO_TC_WIDE", "0"), ("ZOO_PRESTAGE", "0"), ("ZOO_HEAD_BF16", "0"), ("ZOO_STAGE_RAW", "1"), ("ZOO_MERGE_BWD_VAR", "0")])
def test_non_default_kernel_paths_keep_parity(env, val):
    """Every kernel path that has a switch: the gathered-A implicit conv (opt-in), and -- switched OFF here -- the TMA data / weight
    gradients (falls back to GEMM + col2im / im2col + GEMM), programmatic dependent launch, the TMA-im2col forward conv, the TMA-store
    epilogue, the narrow Q-head kernels and the fused IQN merge; the DQN tail runs as the single-CTA kernel ("1") and as the
    multi-CTA kernel ("2") instead of the default four launches.  In every setting the Atari learner parity tests
    (fp32 and bf16) must still pass."""
    e = dict(os.environ)
    e[env] = val
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_learners.py"), "-m", "gpu", "-q", "-x", "-k",
                        "atari and not simt"], env=e, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
