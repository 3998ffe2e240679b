"""GPU parity at BASELINE.json's FULL sizes -- the configurations bench.py times -- against the torch twin directly
(oracle/torch_twin.py, itself <= 1e-5 of the numpy oracle: tests/test_oracle_twin.py), through the C ABI.

A relu unit whose pre-activation is within rounding of zero can take either branch of relu' on two implementations that sum
in a different order, and one flipped unit moves a gradient tensor by far more than the 1e-4 gate.  So the comparison pins the
masks: the device's relu masks are read back (zoo_test_net_activation) and the twin is re-run with exactly those masks in place
of its own relus.  Asserted per configuration:
  (a) masks pinned: loss, per-sample loss, priorities and EVERY gradient tensor within the north_star tolerance
      (fp32 = 3xTF32 on tcgen05: 1e-4 normwise; bf16: 2e-2);
  (b) every unit whose branch differs from the twin's own has a twin pre-activation within FLIP_TOL x the layer's rms
      pre-activation of zero -- i.e. the flips really are rounding-level ties, not errors; the flipped fraction is printed.
References: dqn.jl:102-145, prioritized_dqn.jl:111-151, rainbow.jl:126-221, iqn.jl:159-237; configs Dopamine_{DQN,Rainbow,IQN}_Atari.jl."""
import pytest

from helpers import format_full_size_report, full_size_parity, make_approx, nrel
from oracle import zoo_oracle as O

pytestmark = pytest.mark.gpu

GRAD_TOL = {"fp32": 1e-4, "bf16": 2e-2}
FLIP_TOL = {"fp32": 1e-5, "bf16": 5e-2}   # |twin pre-activation| / layer rms of a unit allowed to flip (bf16 rounds every operand to 2^-9)

CASES = [("dqn_atari_b32", "fp32"), ("per_dqn_atari_b32", "fp32"), ("rainbow_atari_b32", "fp32"), ("rainbow_atari_b1024", "fp32"),
         ("iqn_atari_b32", "fp32"), ("iqn_atari_b512", "fp32"),
         ("dqn_atari_b32", "bf16"), ("rainbow_atari_b1024", "bf16"), ("iqn_atari_b512", "bf16")]


@pytest.mark.parametrize("case,prec", CASES)
def test_bench_config_matches_the_twin_with_pinned_relu_masks(zoo, case, prec):
    rep = full_size_parity(zoo, case, prec)
    print(format_full_size_report(rep))
    tol = GRAD_TOL[prec]
    assert rep["loss_rel"] <= tol, rep
    if "per_sample_err" in rep:
        assert rep["per_sample_err"] <= tol, rep
    if "priority_err" in rep:
        assert rep["priority_err"] <= tol, rep
    for i, e in enumerate(rep["grad_err_pinned"]):
        assert e <= tol, "%s [%s]: gradient tensor %d off by %.3e with the relu masks pinned (> %.0e)" % (case, prec, i, e, tol)
    for i, f in enumerate(rep["flips"]):
        assert f["worst_rel_preact"] <= FLIP_TOL[prec], "%s [%s]: relu layer %d flipped a unit with |pre-activation| = %.2e x rms" % (
            case, prec, i, f["worst_rel_preact"])
        # ties are rare.  Measured (profiles/r02_full_size_parity.txt): fp32 <= 19 of 254 M units; bf16 (every operand rounded to
        # 2^-9) 0.06-0.14 % of a layer
        assert f["flipped"] <= (max(32, f["units"] // 100000) if prec == "fp32" else max(32, f["units"] // 200)), (case, prec, i, f)


def part(b, lo, hi):
    return {k: (v[lo:hi] if v is not None else None) for k, v in b.items()}


def test_dqn_atari_b32_gradient_is_the_mean_of_its_halves(zoo):
    """Size-independent property kept beside the direct comparison: the loss is a mean over the batch, so the gradient of the
    bench batch equals the mean of the gradients of its parts (parts of 16 and 8: the golden-fixture size)."""
    B = 32
    net = O.chain_net(O.nature_cnn(6))
    p, tp = O.init_params(net, 1, 0.05), O.init_params(net, 2, 0.05)

    def build(bs):
        app = make_approx(zoo, net, p, zoo.RMSProp(0.00025, 0.95), max_batch=bs)
        tapp = make_approx(zoo, net, tp, max_batch=bs)
        lrn = zoo.DQNLearner(app, tapp, batch_size=bs)
        app.set_params(p)
        return lrn
    b = O.synth_batch(31, B, (4, 84, 84), 6, atari=True)
    loss, g, _ = build(B).compute_grads(b)
    for parts in (2, 4):
        step = B // parts
        small = build(step)
        acc, lacc = None, 0.0
        for i in range(parts):
            l_i, g_i, _ = small.compute_grads(part(b, i * step, (i + 1) * step))
            lacc += float(l_i) / parts
            acc = [x / parts for x in g_i] if acc is None else [a + x / parts for a, x in zip(acc, g_i)]
        assert abs(float(loss) - lacc) <= 1e-5 * max(1.0, abs(lacc)), (loss, lacc)
        assert max(nrel(x, y) for x, y in zip(g, acc)) <= 1e-4
