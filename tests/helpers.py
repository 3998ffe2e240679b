"""Shared helpers of the parity tests: build a device learner from an oracle-style net + numpy params, compare."""
import numpy as np

from oracle import zoo_oracle as O


def nrel(a, b):
    """normwise relative error max|a-b| / max|b| (used for distributions and gradients)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def obs_shape_of(net):
    first = O.net_chains(net)[0]
    for l in first:
        if l[0] == "conv":
            return (l[1], 84, 84)
        if l[0] == "dense":
            return (l[1],)
    raise ValueError


def make_approx(zoo, net, params, optimizer=None, max_batch=32, max_q=0, precision="fp32", obs_shape=None):
    app = zoo.NeuralNetworkApproximator(zoo.from_oracle_net(net), optimizer, obs_shape or obs_shape_of(net), max_batch, max_q, precision)
    app.set_params(params)
    return app


def assert_grads_close(got, ref, tol, what=""):
    assert len(got) == len(ref)
    worst = 0.0
    for i, (g, r) in enumerate(zip(got, ref)):
        assert g.shape == r.shape, (what, i, g.shape, r.shape)
        e = nrel(g, r)
        worst = max(worst, e)
        assert e <= tol, "%s: grad tensor %d %s normwise rel err %.3e > %.1e" % (what, i, r.shape, e, tol)
    return worst
