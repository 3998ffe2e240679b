"""libzoo_sm100 host package: a B200-native (sm_100a) drop-in for ONE hot path of ReinforcementLearningZoo.jl --
``RLBase.update!`` of the value-based learners on a replay minibatch, plus the prioritized-replay SumTree.

The names mirror the reference (``BasicDQNLearner`` … ``IQNLearner``, ``NeuralNetworkApproximator``,
``ImplicitQuantileNet``, ``DuelingNetwork``, ``SumTree``, ``update``); all arithmetic is in
``libzoo_sm100.so`` (csrc/, C ABI in include/zoo_sm100.h).  There is no CPU fallback."""
from ._lib import LIB_PATH, ZooError, lib  # noqa: F401
from .checkpoint import load_policy, save_policy  # noqa: F401
from .dp import DataParallel, shard_batch, shard_bounds  # noqa: F401
from .learners import (SARTS, BasicDQNLearner, DQNLearner, IQNLearner, PrioritizedDQNLearner, QRDQNLearner, RainbowLearner,  # noqa: F401
                       REMDQNLearner, update, update_trajectory)
from .nn import (ADAM, Chain, CrossCor, Dense, DuelingNetwork, Flatten, ImplicitQuantileNet, NatureCNN,  # noqa: F401
                 NeuralNetworkApproximator, RMSProp, Scale255, from_oracle_net)
from .replay import ReplayShard  # noqa: F401
from .sumtree import SumTree  # noqa: F401
