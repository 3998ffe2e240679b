#!/usr/bin/env python
"""bench.py -- learner updates/s of the value-based learner update (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--precision fp32|bf16] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one `update!(learner, batch)` on one synthetic replay minibatch per GPU (weak scaling:
every rank holds its own replay shard and minibatch; gradients are all-reduced with NCCL inside the update).
Workloads (BASELINE.json configs): dqn_atari_b32 (default = configs[1], the one the metric is quoted on),
rainbow_atari_b32, rainbow_atari_b1024, iqn_atari_b512 (global 4096 at 8 GPUs), basic_dqn_cartpole_b32.

One JSON line on stdout (rank 0).  Legs:
  value      K updates, batch already resident in HBM, L2 flushed before every step, CUDA events on the learner's
             stream around each update, summed; max over ranks.
  e2e        K updates through the public API with pinned HOST buffers: H2D of the batch and D2H of the loss inside
             the timed region (wall clock bracketed by synchronisation; max over ranks).
  roofline   per-kernel device times from the library's in-situ launch profiler (CUDA events on the learner's stream
             behind a parked delay kernel), averaged over profiled updates; the dominant kernel's algorithmic
             bytes / flops against MEASURED_PEAKS.json.
  cpu_baseline  the oracle's torch-CPU restatement of the same update on the host cores (rank 0, N = 1 only).
--impl reference times only that CPU arm (the reference is pure Julia and `julia` is not in this image)."""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (learner, net, per-GPU batch, optimiser, extras)
    "dqn_atari_b32": dict(kind="dqn", B=32, A=6, opt=("rmsprop", 0.00025, 0.95), n=1, atari=True),
    "rainbow_atari_b32": dict(kind="rainbow", B=32, A=6, opt=("adam", 0.0000625), n=3, atari=True, priority=True),
    "rainbow_atari_b1024": dict(kind="rainbow", B=1024, A=6, opt=("adam", 0.0000625), n=3, atari=True, priority=True),
    "iqn_atari_b512": dict(kind="iqn", B=512, A=6, opt=("adam", 0.00005), n=3, atari=True, N=64),
    "iqn_atari_b32": dict(kind="iqn", B=32, A=6, opt=("adam", 0.00005), n=3, atari=True, N=64),
    "basic_dqn_cartpole_b32": dict(kind="basic", B=32, A=2, opt=("adam", 0.001), n=1, atari=False),
}
# FLOP per sample (2*MAC fwd; bwd = 2*fwd minus first-layer dgrad), SURVEY.md §8(d) / BASELINE.md §3
FLOP_PER_SAMPLE = {"dqn": 152_836_096, "rainbow": 122_052_608, "iqn": 2_311_012_352, "basic": 203_776}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]), tf_sus=float(p["bf16_tflops_sustained"]), src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


def oracle_net(w):
    from oracle import zoo_oracle as O
    if w["kind"] == "basic":
        return O.chain_net(O.mlp([4, 128, 128, 2]))
    if w["kind"] == "dqn":
        return O.chain_net(O.nature_cnn(w["A"]))
    if w["kind"] == "rainbow":
        return O.chain_net(O.nature_cnn(w["A"] * 51))
    return O.atari_iqn_net(w["A"])


def synth(w, seed, B=None):
    from oracle import zoo_oracle as O  # synthetic inputs only; not on the measured path
    B = B or w["B"]
    return O.synth_batch(seed, B, (4, 84, 84) if w["atari"] else (4,), w["A"], atari=w["atari"], priority=w.get("priority", False))


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_arm(w, steps, warmup, budget_s=None):
    """The reference's update on the host cores: oracle/torch_twin.CpuLearner (kind "port": a restatement; the Julia
    reference itself cannot run here).  Returns (updates/s, steps timed, threads, sample description)."""
    import torch
    from oracle import torch_twin as T
    from oracle import zoo_oracle as O
    net = oracle_net(w)
    p, tp = O.init_params(net, 1), O.init_params(net, 2)
    kind = w["kind"]
    cfg = dict(gamma=0.99)
    if w["kind"] == "dqn":
        cfg.update(n=w["n"], double_dqn=True)
    elif w["kind"] == "rainbow":
        cfg.update(n_actions=w["A"], n=w["n"])
    B = w["B"]
    if w["kind"] == "iqn":
        B = min(B, 8)  # bounded sample: 2.3 GFLOP/sample; the rate is scaled back to batches of w["B"] below
    if w["kind"] == "rainbow" and B > 64:
        B = 64
    lrn = T.CpuLearner(kind, net, p, tp, w["opt"], **cfg)
    pool = [synth(w, 100 + i, B) for i in range(4)]
    rng = np.random.default_rng(0)
    taus = [(rng.random((B, w.get("N", 1)), dtype=np.float32), rng.random((B, w.get("N", 1)), dtype=np.float32)) for _ in range(4)]

    def one(i):
        b = pool[i % 4]
        if w["kind"] == "iqn":
            lrn.update(b, *taus[i % 4])
        else:
            lrn.update(b)
    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        one(i)
        done += 1
        if budget_s and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    rate = done / dt * (B / w["B"])  # updates of the workload's batch size per second
    sample = "%d updates of batch %d (of %d) in %.1f s" % (done, B, w["B"], dt)
    return rate, done, torch.get_num_threads(), sample


def host_cores():
    """Physical cores this process may run on (what torch picks by itself when OMP_NUM_THREADS is not forced)."""
    cpus = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else set(range(os.cpu_count() or 1))
    try:
        cores, cur = set(), {}
        for line in open("/proc/cpuinfo"):
            if ":" in line:
                k, v = (x.strip() for x in line.split(":", 1))
                cur[k] = v
            elif cur:
                if int(cur.get("processor", -1)) in cpus:
                    cores.add((cur.get("physical id", "0"), cur.get("core id", cur.get("processor"))))
                cur = {}
        if cur and int(cur.get("processor", -1)) in cpus:
            cores.add((cur.get("physical id", "0"), cur.get("core id", cur.get("processor"))))
        return max(1, len(cores))
    except Exception:
        return max(1, len(cpus))


def config_of(args, w, world):
    """The `config` object of the JSON line -- the same keys and values for both arms (the driver compares them)."""
    prec = args.precision
    return {"workload": args.workload, "learner": w["kind"], "per_gpu_batch": w["B"], "global_batch": w["B"] * world,
            "precision_mode": {"fp32": "fp32 storage, contractions as 3xTF32 on tcgen05 (parity 1e-4)", "bf16": "bf16 operands on tcgen05, fp32 accumulate"}.get(prec, prec),
            "value_counts": "per-GPU minibatches of %d per second (one all-reduced optimiser step per %d of them)" % (w["B"], world),
            "l2": "flushed (256 MB fill) before every timed step", "cuda_graph": True, "parallelism": "dp%d" % world}


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    # torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the CPU arm is ONE replica that may use every host core
    # (the JULIA_NUM_THREADS analogue), so undo that here -- rank 0 is the only rank that runs
    torch.set_num_threads(host_cores())
    rate, done, threads, sample = cpu_arm(w, args.steps, args.warmup, budget_s=240)
    line = {
        "impl": "reference", "metric": "learner updates/s", "value": rate, "unit": "updates/s", "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_of(args, w, args.gpus),
        "note": "CPU arm: torch-CPU restatement of the reference update (julia absent); ONE replica on all host cores, whatever --gpus says "
                "(fp32 arithmetic; the precision_mode / l2 / cuda_graph keys describe the GPU arm it is compared with)",
        "cpu_baseline": {"value": rate, "unit": "updates/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- clocks
class Clocks:
    Q = "clocks.sm,clocks.max.sm,utilization.gpu,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.samples, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, windows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        inw = [s for t, s in self.samples if any(a <= t <= b for a, b in windows)] or [s for _, s in self.samples]
        busy = [s for s in inw if s[2].isdigit() and int(s[2]) > 0] or inw
        mhz = [int(s[0]) for s in busy if s[0].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in busy for n, v in zip(names, s[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(mhz) if mhz else None, "sm_max_mhz": int(busy[0][1]) if busy and busy[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(busy)}


# ----------------------------------------------------------------------------------------------- roofline
def kernel_work(label, w, esz, n_params):
    """Algorithmic (flops, bytes) of one launch from its profiler label (DESIGN.md "kernels")."""
    m = re.search(r"\[(\d+)x(\d+)x(\d+)\]", label)
    if m and (label.startswith("fwd") or label.startswith("wgrad") or label.startswith("dgrad")):
        M, N, K = (int(x) for x in m.groups())
        if "im2col" in label or "col2im" in label:
            return 0.0, float(M * K * esz * 2 if "im2col" in label else M * N * esz * 2)
        out_b = 4 if label.startswith("wgrad") else esz
        return 2.0 * M * N * K, float((M * K + N * K) * esz + M * N * out_b)
    m2 = re.search(r"\[(\d+)x(\d+)\]", label)
    if m2 and ("im2col" in label or "col2im" in label):  # explicit (un)folding: the cols matrix once + the activation once
        return 0.0, float(int(m2.group(1)) * int(m2.group(2)) * esz * 2)
    m1 = re.search(r"obs2nhwc\[(\d+)\]", label)
    if m1:  # u8 frame stack in, 4-channel NHWC rows out
        return 0.0, float(int(m1.group(1)) * (4 + 4 * esz))
    if label.startswith("opt"):
        per = 6 if w["opt"][0] == "rmsprop" else 8  # p r/w, g r + re-zero, state r/w
        return 0.0, float(n_params * 4 * per + (n_params * 2 if esz == 2 else 0))
    return 0.0, 0.0


def roofline(profiles, w, esz, n_params, pk, traffic):
    agg = {}
    for prof in profiles:
        for i, (label, ms) in enumerate(prof):
            a = agg.setdefault((i, label), [])
            a.append(ms)
    per = [(lab, statistics.mean(v)) for (i, lab), v in sorted(agg.items())]
    step_ms = sum(ms for _, ms in per)
    by = {}
    for lab, ms in per:
        key = lab.split("[")[0] + "|" + lab.split("|")[-1].split(":")[0]
        by[key] = by.get(key, 0.0) + ms
    # (the dominant kernel is one of OURS: at N > 1 the NCCL all-reduce of the late bucket can be the longest launch of the eager
    # profile, but it runs on the comm lane under the convolution backward and has no per-GPU roofline)
    lab, ms = max((x for x in per if not x[0].startswith("allreduce")), key=lambda x: x[1])
    fl, by_ = kernel_work(lab, w, esz, n_params)
    ridge = pk["tf_sus"] * 1e12 / (pk["hbm"] * 1e9)
    if fl > 0 and by_ > 0 and fl / by_ > ridge:
        bound, ach, peak, unit = "tensor", fl / (ms * 1e-3) / 1e12, pk["tf_sus"], "TFLOP/s"
    else:
        bound, ach, peak, unit = "hbm", by_ / (ms * 1e-3) / 1e9, pk["hbm"], "GB/s"
    top = sorted(by.items(), key=lambda x: -x[1])[:8]
    return {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak if peak else None,
            "traffic": traffic.get(lab.split("|")[0]), "kernel": lab, "kernel_us": ms * 1e3, "kernel_share_of_step": ms / step_ms if step_ms else None,
            "peak_source": pk["src"], "profiled_step_us": step_ms * 1e3, "launches_per_step": len(per),
            "top_phases_us": {k: round(v * 1e3, 2) for k, v in top}}


# ----------------------------------------------------------------------------------------------- our arm
def build_learner(zoo, w, prec, dp, rank):
    """Learner + online / target approximators of a workload with Flux-style glorot weights (online != target, as after the
    first target period)."""
    from oracle import zoo_oracle as O  # weight init only
    B, kind = w["B"], w["kind"]
    net = oracle_net(w)
    model = zoo.from_oracle_net(net)
    opt = zoo.RMSProp(*w["opt"][1:]) if w["opt"][0] == "rmsprop" else zoo.ADAM(w["opt"][1])
    obs_shape = (4, 84, 84) if w["atari"] else (4,)
    nq = w.get("N", 0)
    app = zoo.NeuralNetworkApproximator(model, opt, obs_shape, B, nq, prec)
    tapp = None if kind == "basic" else zoo.NeuralNetworkApproximator(model, None, obs_shape, B, nq, prec)
    if tapp is not None:
        tapp.set_params(O.init_params(net, 2))
    if kind == "basic":
        lrn = zoo.BasicDQNLearner(app, batch_size=B, dp=dp)
    elif kind == "dqn":
        lrn = zoo.DQNLearner(app, tapp, batch_size=B, update_horizon=w["n"], dp=dp)  # Double-DQN on (dqn.jl:58)
    elif kind == "rainbow":
        lrn = zoo.RainbowLearner(app, tapp, n_actions=w["A"], Vmax=10.0, Vmin=-10.0, update_horizon=w["n"], batch_size=B, dp=dp)
    else:
        lrn = zoo.IQNLearner(app, tapp, n_actions=w["A"], N=nq, N_prime=nq, N_em=64, batch_size=B, update_horizon=w["n"], device_seed=rank, dp=dp)
    app.set_params(O.init_params(net, 1))  # the constructors force-sync online := target (dqn.jl:60)
    return lrn, app, tapp


def make_pools(w, rank, dev, torch, n_pool=None):
    """synthetic replay minibatches: a pool of distinct batches per rank, pinned on the host and resident on the device"""
    n_pool = n_pool or (8 if w["B"] <= 64 else 2)
    host_pool, dev_pool = [], []
    for i in range(n_pool):
        b = synth(w, 1000 * rank + i)
        hb, db = {}, {}
        for k, v in b.items():
            t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
            hb[k] = t.numpy()
            db[k] = t.to(dev)
        host_pool.append((hb, [t for t in hb.values()]))
        dev_pool.append(db)
    return host_pool, dev_pool


def step_work(w, B, n_params):
    """Algorithmic work of ONE update of per-GPU batch B: flops (2*MAC, fwd + bwd) and the HBM bytes that must move
    at least once: the u8 frames of s and s', the online and target parameters read, the optimiser state read + written and the
    parameters written (RMSProp: 1 state tensor, Adam: 2)."""
    flops = float(FLOP_PER_SAMPLE[w["kind"]]) * B
    obs = (4 * 84 * 84) if w["atari"] else 4 * 4
    states = 1 if w["opt"][0] == "rmsprop" else 2
    targets = 0 if w["kind"] == "basic" else 1
    byts = 2.0 * B * obs + n_params * 4.0 * (1 + targets + 2 * states + 1)
    return flops, byts


def step_roofline(w, B, n_params, step_s, pk, tensor_div=1.0):
    """Whole-step roofline: the step's algorithmic flops and bytes against the measured peaks; the binding one is whichever
    takes longer at its peak.  tensor_div: 3 for the fp32-grade mode (three TF32 products per contraction)."""
    fl, by = step_work(w, B, n_params)
    t_tensor = fl / (pk["tf_sus"] * 1e12 / tensor_div)
    t_hbm = by / (pk["hbm"] * 1e9)
    if t_tensor >= t_hbm:
        ach = fl / step_s / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": pk["tf_sus"] / tensor_div, "unit": "TFLOP/s", "frac": ach / (pk["tf_sus"] / tensor_div),
                "flops": fl, "bytes": by, "bound_us": t_tensor * 1e6}
    ach = by / step_s / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "flops": fl, "bytes": by, "bound_us": t_hbm * 1e6}


def time_updates(lrn, dev_pool, K, stream, flush, torch):
    """K graph-replayed updates, L2 flushed before each, CUDA events on the learner's stream around each update; returns ms."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    with torch.cuda.stream(stream):
        for i in range(K):
            flush.fill_(i & 0xFF)
            ev[i][0].record(stream)
            lrn.update(dev_pool[i % len(dev_pool)], want_loss=False, want_priorities=False)
            ev[i][1].record(stream)
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev)


def trace(*a):
    if os.environ.get("ZOO_BENCH_TRACE"):
        print("[bench r%s %.2f]" % (os.environ.get("RANK", "0"), time.time() % 1000), *a, file=sys.stderr, flush=True)


def run_ours(args, w):
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("ZOO_BENCH_WATCHDOG_S", "900")), exit=True)  # never hang a GPU box
    import torch
    import torch.distributed as dist

    import zoo_loader
    from oracle import zoo_oracle as O  # weight init + synthetic batches only
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    zoo = zoo_loader.load()
    from rlzoo_b200 import _lib as ZL
    ZL.check(zoo.lib().zoo_set_device(local))
    dp = zoo.DataParallel(rank, world) if world > 1 else None

    B, kind, prec = w["B"], w["kind"], args.precision
    nq = w.get("N", 0)
    lrn, app, tapp = build_learner(zoo, w, prec, dp, rank)
    n_params = sum(int(np.prod(s)) for s in app.shapes)
    lrn.use_graph(True)

    host_pool, dev_pool = make_pools(w, rank, dev, torch)
    n_pool = len(dev_pool)
    h2d = sum(v.nbytes for v in host_pool[0][0].values())
    per = w.get("priority", False)
    d2h = 4 + (4 * B if per else 0)

    stream = torch.cuda.ExternalStream(lrn.stream(), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # 2x the 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    trace("learner built")
    clocks = Clocks(local) if rank == 0 else None
    windows = []
    # ---- warm-up: eager pass, graph capture, replays
    for i in range(max(3, args.warmup)):
        lrn.update(dev_pool[i % n_pool], want_loss=False, want_priorities=False)
    lrn.sync()

    trace("warm-up done")
    # ---- leg 1: value (HBM-resident inputs, L2 flushed before every step, events on the learner's stream)
    K = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    t_w0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for i in range(K):
            flush.fill_(i & 0xFF)
            ev[i][0].record(stream)
            lrn.update(dev_pool[i % n_pool], want_loss=False, want_priorities=False)
            ev[i][1].record(stream)
    barrier()
    windows.append((t_w0, time.perf_counter()))
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    dev_ms = maxr(dev_ms)
    launches = lrn.last_launches() * K

    trace("leg 1 done")
    # ---- leg 1b: same, back to back without the flush (weights L2-resident: the steady state of a real run)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(K):
            lrn.update(dev_pool[i % n_pool], want_loss=False, want_priorities=False)
        e1.record(stream)
    barrier()
    warm_ms = maxr(e0.elapsed_time(e1))

    trace("leg 1b done")
    # ---- leg 2: e2e through the public API with pinned host buffers; loss (and PER priorities) read back every step
    # (a) blocking: update!(learner, batch) returns once this update's loss is on the host
    for i in range(3):
        lrn.update(host_pool[i % n_pool][0])
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        lrn.update(host_pool[i % n_pool][0])  # blocks until the loss is on the host
    torch.cuda.synchronize()
    e2e_block_s = maxr(time.perf_counter() - t0)
    windows.append((t0, time.perf_counter()))
    barrier()
    # (b) pipelined (the headline e2e): submit(batch i + 1) before collect(loss i) -- every step still copies its own batch from
    # pinned host memory and its own loss back, but the H2D of the next batch overlaps this update's kernels (copy stream + two
    # staging slots, zoo_learner_submit / zoo_learner_collect).  The reference's update! is as asynchronous (dqn.jl:138-140).
    per = "priority" in host_pool[0][0]
    for i in range(3):
        lrn.submit(host_pool[i % n_pool][0], want_priorities=per)
        lrn.collect()
    barrier()
    t0 = time.perf_counter()
    lrn.submit(host_pool[0][0], want_priorities=per)
    for i in range(1, K):
        lrn.submit(host_pool[i % n_pool][0], want_priorities=per)
        lrn.collect()  # loss (and priorities) of update i - 1
    lrn.collect()
    torch.cuda.synchronize()
    e2e_s = maxr(time.perf_counter() - t0)
    windows.append((t0, time.perf_counter()))
    barrier()
    clk = clocks.stop(windows) if clocks else None

    trace("leg 2 done")
    # ---- leg 3 (DQN workload): the same update fed from a device-resident replay shard: per step the host
    # pushes 4 new frames (update_freq = 4, Dopamine_DQN_Atari.jl:45) and sends B uniform draws; sampling, frame stacking, the
    # n-step fold and the update stay on the device; the loss is read back every step
    e2e_replay = None
    if kind == "dqn" and w["atari"]:
        shard = zoo.ReplayShard(100_000, (84, 84), 4, w["n"], 0.99, max_batch=B)
        rng = np.random.default_rng(rank)
        nfill = 20_000
        shard.push_many(rng.integers(0, 256, (nfill, 84, 84), dtype=np.uint8), rng.integers(0, w["A"], nfill), rng.integers(-1, 2, nfill).astype(np.float32),
                        (rng.random(nfill) < 0.02))
        newf = rng.integers(0, 256, (64, 84, 84), dtype=np.uint8)
        draws = rng.random((K + 8, B))
        for i in range(3):
            shard.update(lrn, draws[i])
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            for j in range(4):
                shard.push(newf[(4 * i + j) % 64], 1, 0.0, False)
            shard.update(lrn, draws[i + 3])
        torch.cuda.synchronize()
        rp_s = maxr(time.perf_counter() - t0)
        e2e_replay = {"value": world * K / rp_s, "unit": "updates/s", "h2d_bytes_per_step": 4 * (84 * 84 + 9) + B * 8, "d2h_bytes_per_step": 4,
                      "note": "update!(learner, trajectory) on a device-resident replay shard: 4 frame pushes + B draws per step"}
        shard.close()
        barrier()
    trace("leg 3 done")

    # ---- roofline leg (rank 0 profiles; the other ranks must join the collectives, so they run the same eager updates)
    profs = []
    for i in range(4):
        pr = lrn.profile(dev_pool[i % n_pool], delay_us=3000 if B <= 64 else 20000)
        if i:
            profs.append(pr)
    trace("profiles done")
    esz = 2 if prec == "bf16" else 4
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload + ":" + prec, {})
    except Exception:
        traffic = {}
    pk = peaks()
    roof = roofline(profs, w, esz, n_params, pk, traffic)
    flop_step = FLOP_PER_SAMPLE[kind] * B
    step_s = dev_ms * 1e-3 / K
    roof["step"] = step_roofline(w, B, n_params, step_s, pk, 3.0 if prec == "fp32" else 1.0)

    # ---- extra.c5: the north_star data-parallel configuration (Dopamine IQN, N = N' = 64, 512 samples per GPU = global 4096
    # at 8 GPUs; Dopamine_IQN_Atari.jl:26-68), timed the same way (graph replay, L2 flushed, events, max over ranks) in the
    # fp32-grade mode and in bf16.  The headline `value` stays the dqn_atari_b32 line.
    c5 = None
    if args.c5 and args.workload == "dqn_atari_b32":
        lrn.close()
        del lrn, app, tapp, dev_pool, host_pool
        c5 = {"workload": "iqn_atari_b512", "per_gpu_batch": 512, "global_batch": 512 * world, "quantiles": "N = N' = 64",
              "steps": args.c5_steps, "gflop_per_step_per_gpu": FLOP_PER_SAMPLE["iqn"] * 512 / 1e9}
        w5 = WORKLOADS["iqn_atari_b512"]
        for p5 in ("fp32", "bf16"):
            l5, a5, t5 = build_learner(zoo, w5, p5, dp, rank)
            l5.use_graph(True)
            _, pool5 = make_pools(w5, rank, dev, torch, 2)
            st5 = torch.cuda.ExternalStream(l5.stream(), device=dev)
            for i in range(3):
                l5.update(pool5[i % 2], want_loss=False, want_priorities=False)
            l5.sync()
            barrier()
            ms5 = maxr(time_updates(l5, pool5, args.c5_steps, st5, flush, torch))
            barrier()
            s5 = ms5 * 1e-3 / args.c5_steps
            np5 = sum(int(np.prod(x)) for x in a5.shapes)
            c5[p5] = {"ms_per_step": ms5 / args.c5_steps, "updates_per_s": world / s5, "samples_per_s": world * 512 / s5,
                      "step_tflops_per_gpu": FLOP_PER_SAMPLE["iqn"] * 512 / s5 / 1e12,
                      "roofline_step": step_roofline(w5, 512, np5, s5, pk, 3.0 if p5 == "fp32" else 1.0), "launches_per_step": l5.last_launches()}
            l5.close()
            del l5, a5, t5, pool5
            torch.cuda.synchronize()
        trace("c5 done")

    if world > 1:
        dist.barrier()
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            rate, done, threads, sample = cpu_arm(w, 10_000, 2, budget_s=15)
            cpu = {"value": rate, "unit": "updates/s", "cores": threads, "kind": "port", "sample": sample}
        line = {
            "metric": "learner updates/s", "value": world * K / (dev_ms * 1e-3), "unit": "updates/s", "n_gpus": world, "steps": K,
            "warmup": max(3, args.warmup), "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if prec == "bf16" else "f32", "data": "synthetic",
            "config": config_of(args, w, world),
            "value_l2_warm": world * K / (warm_ms * 1e-3),
            "gflop_per_step_per_gpu": flop_step / 1e9,
            "step_tflops_per_gpu": flop_step / step_s / 1e12,
            "step_frac_of_tensor_peak": flop_step / step_s / 1e12 / pk["tf_sus"],
            "roofline": roof,
            "cpu_baseline": cpu,
            "e2e": {"value": world * K / e2e_s, "unit": "updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "mode": "submit(i+1) before collect(i): the next batch's H2D overlaps this update's kernels; every loss read back",
                    "blocking_value": world * K / e2e_block_s},
            "e2e_replay": e2e_replay,
            "extra": {"c5": c5} if c5 else None,
            "gpu_launches": launches,
            "clocks": clk,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # orderly collective teardown: learner (and its captured graph) first, then our NCCL communicator, then torch's
        if c5 is None:
            lrn.close()
        torch.cuda.synchronize()
        dist.barrier()
        dp.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dqn_atari_b32", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16", "tf32", "fp32_simt"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-c5", dest="c5", action="store_false", help="skip the extra.c5 block (IQN 512 per GPU, fp32 + bf16) of the default workload")
    ap.add_argument("--c5-steps", type=int, default=10)
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
