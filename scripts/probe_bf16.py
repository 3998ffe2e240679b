import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, zoo_loader
from oracle import zoo_oracle as O
from helpers import make_approx, nrel
zoo = zoo_loader.load()
for B in (8, 32):
    net = O.chain_net(O.nature_cnn(6))
    p, tp = O.init_params(net, 1, 0.05), O.init_params(net, 2, 0.05)
    b = O.synth_batch(11, B, (4, 84, 84), 6, atari=True)
    ref = O.dqn_grads(net, p, tp, b)
    ref64 = O.dqn_grads(net, p, tp, b, dtype=np.float64)
    print("   fp32 oracle vs fp64 oracle:", ["%.1e" % nrel(g, r) for g, r in zip(ref["grads"], ref64["grads"])])
    for prec in ("fp32", "bf16"):
        app = make_approx(zoo, net, p, zoo.RMSProp(0.00025, 0.95), max_batch=B, precision=prec)
        tapp = make_approx(zoo, net, tp, max_batch=B, precision=prec)
        lrn = zoo.DQNLearner(app, tapp, batch_size=B)
        app.set_params(p)
        loss, grads, _ = lrn.compute_grads(b)
        q = app(b["state"]); qr, _ = O.net_forward(net, p, b["state"])
        print("B", B, prec, "loss", loss, float(ref["loss"]), "fwd err", nrel(q, qr))
        print("   grad errs vs fp32 oracle:", ["%.1e" % nrel(g, r) for g, r in zip(grads, ref["grads"])])
        print("   cos:", ["%.5f" % (np.dot(g.ravel(), r.ravel()) / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30)) for g, r in zip(grads, ref["grads"])])
        print("   l2 rel:", ["%.1e" % (np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30)) for g, r in zip(grads, ref["grads"])])
