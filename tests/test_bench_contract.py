"""bench.py's driver-facing contract, as far as it can be checked without a GPU: the reference arm prints ONE JSON line with
the agreed keys (and only rank 0 prints under a multi-rank launch), and the roofline byte model parses every profiler label."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_bench(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "learner updates/s" and d["unit"] == "updates/s" and d["higher_is_better"] is True
    assert d["config"]["workload"] == "dqn_atari_b32" and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1.0  # one update per step


def test_reference_arm_other_ranks_stay_silent():
    r = run_bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_roofline_byte_model_parses_every_label_kind():
    import bench
    w = bench.WORKLOADS["dqn_atari_b32"]
    fl, by = bench.kernel_work("fwd:t6[64x512x7744]|launch_tc:956", w, 4, 0)
    assert fl == 2.0 * 64 * 512 * 7744 and by == (64 * 7744 + 512 * 7744) * 4 + 64 * 512 * 4
    fl, by = bench.kernel_work("wgrad:t6[512x7744x32]|launch_tc:956", w, 2, 0)
    assert by == (512 * 32 + 7744 * 32) * 2 + 512 * 7744 * 4  # gradients accumulate in fp32 whatever the operand type
    assert bench.kernel_work("im2col:t0[14112x256]|launch_im2col_nhwc:367", w, 4, 0) == (0.0, 14112 * 256 * 4 * 2.0)
    assert bench.kernel_work("dgrad:t4[3872x576x64]|launch_col2im_nhwc:383", w, 4, 0)[1] > 0
    assert bench.kernel_work("obs2nhwc[225792]|chain_forward:449", w, 4, 0) == (0.0, 225792 * 20.0)
    n = 4046502
    assert bench.kernel_work("opt|opt_range:181", w, 4, n) == (0.0, n * 4 * 6.0)  # RMSProp: p r/w, g r + re-zero, acc r/w
    assert bench.kernel_work("head|launch_dqn_head:335", w, 4, n) == (0.0, 0.0)


def test_whole_step_roofline_model_matches_the_survey_table():
    """DQN B=32 82.7 MB (HBM-bound, 12.6 us); Rainbow B=32 ~119 MB; Rainbow B=1024 and IQN tensor-bound."""
    import bench
    pk = dict(hbm=6548.5, tf_burst=1598.9, tf_sus=1375.3, src="measured")
    fl, by = bench.step_work(bench.WORKLOADS["dqn_atari_b32"], 32, 4046502)
    assert fl == 152_836_096 * 32 and abs(by / 1e6 - 82.7) < 0.1
    r = bench.step_roofline(bench.WORKLOADS["dqn_atari_b32"], 32, 4046502, 283.5e-6, pk, 3.0)
    assert r["bound"] == "hbm" and abs(r["bound_us"] - 12.6) < 0.1 and abs(r["frac"] - 0.0446) < 0.001
    _, by = bench.step_work(bench.WORKLOADS["rainbow_atari_b32"], 32, 4200402)
    assert abs(by / 1e6 - 119.4) < 0.5
    assert bench.step_roofline(bench.WORKLOADS["rainbow_atari_b1024"], 1024, 4200402, 1e-3, pk)["bound"] == "tensor"
    r = bench.step_roofline(bench.WORKLOADS["iqn_atari_b512"], 512, 4549862, 4.27e-3, pk)
    assert r["bound"] == "tensor" and abs(r["frac"] - 0.2015) < 0.002


def test_both_arms_print_the_same_config_object():
    import argparse

    import bench
    a = argparse.Namespace(workload="dqn_atari_b32", precision="fp32", gpus=4)
    c = bench.config_of(a, bench.WORKLOADS["dqn_atari_b32"], 4)
    assert c["global_batch"] == 128 and c["parallelism"] == "dp4" and "model" not in c
    r = run_bench(["--impl", "reference", "--gpus", "4", "--steps", "1", "--warmup", "0"], {"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "4", "OMP_NUM_THREADS": "1"})
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["config"] == c
    assert d["cpu_baseline"]["cores"] == bench.host_cores()  # torchrun's OMP_NUM_THREADS=1 does not starve the CPU arm
