"""2-rank probe of the data-parallel path with progress prints (run under torchrun)."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(60, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def log(*a):
    print("[r%d %.1f]" % (rank, time.time() % 1000), *a, flush=True)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
log("pg up")
dist.barrier(); torch.cuda.synchronize()
log("barrier ok")
import zoo_loader
zoo = zoo_loader.load()
from rlzoo_b200 import _lib as ZL
ZL.check(zoo.lib().zoo_set_device(local))
dp = zoo.DataParallel(rank, world)
log("dp up")
from oracle import zoo_oracle as O
net = O.chain_net(O.mlp([4, 64, 2]))
p, tp = O.init_params(net, 1, 0.1), O.init_params(net, 2, 0.1)
full = O.synth_batch(5, 16, (4,), 2, atari=False)
mine = zoo.shard_batch(full, rank, world)
app = zoo.NeuralNetworkApproximator(zoo.from_oracle_net(net), zoo.ADAM(), (4,), 8, 0, "fp32")
tapp = zoo.NeuralNetworkApproximator(zoo.from_oracle_net(net), None, (4,), 8, 0, "fp32")
app.set_params(p); tapp.set_params(tp)
lrn = zoo.DQNLearner(app, tapp, batch_size=8, dp=dp)
app.set_params(p)
log("learner up")
loss, grads, _ = lrn.compute_grads(mine)
log("eager grads ok", loss)
ref = O.dqn_grads(net, p, tp, full)
err = max(np.abs(g - r).max() / (np.abs(r).max() + 1e-30) for g, r in zip(grads, ref["grads"]))
log("DP grads vs single-process oracle on the global batch: loss %.6f vs %.6f, worst grad err %.2e" % (loss, ref["loss"], err))
assert err < 1e-4 and abs(loss - ref["loss"]) < 1e-5
lrn.use_graph(True)
for i in range(3):
    lrn.compute_grads(mine)
    log("graph iter", i)
pr = lrn.profile(mine, delay_us=1000)
log("profile ok", len(pr))
dist.barrier(); torch.cuda.synchronize()
lrn.close()
dp.close()
log("done")
dist.destroy_process_group()
