"""ctypes binding of libzoo_sm100.so (include/zoo_sm100.h).  No fallback of any kind: if the CUDA
library is missing or the device is not sm_100, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZOO_SM100_LIB") or os.path.join(_HERE, "libzoo_sm100.so")  # override: A/B runs of two builds

ZOO_MAX_LAYERS = 12
LAYER_SCALE255, LAYER_CONV, LAYER_FLATTEN, LAYER_DENSE = 0, 1, 2, 3
NET_CHAIN, NET_IQN, NET_DUELING = 0, 1, 2
PREC_FP32, PREC_BF16, PREC_TF32, PREC_FP32_SIMT = 0, 1, 2, 3
OBS_U8, OBS_F32 = 0, 1
OPT_RMSPROP, OPT_ADAM = 0, 1
LEARNER_BASIC_DQN, LEARNER_DQN, LEARNER_PRIORITIZED_DQN, LEARNER_RAINBOW, LEARNER_IQN, LEARNER_QR_DQN, LEARNER_REM_DQN = range(7)
DRAW_F32, DRAW_F64 = 0, 1
SAMPLE_INDEPENDENT, SAMPLE_STRATIFIED, REPLAY_UNIFORM = 0, 1, 2
STATUS = {0: "OK", 1: "BAD_ARG", 2: "CUDA_ERROR", 3: "NCCL_ERROR", 4: "UNSUPPORTED_ARCH"}


class ZooError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("libzoo_sm100: %s: %s" % (STATUS.get(status, status), msg))
        self.status = status


class Layer(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_in", C.c_int32), ("n_out", C.c_int32), ("ksize", C.c_int32),
                ("stride", C.c_int32), ("pad", C.c_int32), ("relu", C.c_int32)]


class NetDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("precision", C.c_int32), ("in_c", C.c_int32), ("in_h", C.c_int32),
                ("in_w", C.c_int32), ("max_batch", C.c_int32), ("max_quantiles", C.c_int32),
                ("n_a", C.c_int32), ("n_b", C.c_int32), ("n_c", C.c_int32),
                ("a", Layer * ZOO_MAX_LAYERS), ("b", Layer * ZOO_MAX_LAYERS), ("c", Layer * ZOO_MAX_LAYERS)]


class LearnerCfg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("batch_size", C.c_int32), ("n_actions", C.c_int32), ("gamma", C.c_float),
                ("update_horizon", C.c_int32), ("double_dqn", C.c_int32), ("beta_priority", C.c_float),
                ("n_atoms", C.c_int32), ("v_min", C.c_float), ("v_max", C.c_float), ("support", C.c_void_p),
                ("delta_z", C.c_float), ("kappa", C.c_float), ("N", C.c_int32), ("N_prime", C.c_int32),
                ("N_em", C.c_int32), ("seed", C.c_uint64)]


class Batch(C.Structure):
    _fields_ = [("B", C.c_int32), ("obs_dtype", C.c_int32), ("state", C.c_void_p), ("next_state", C.c_void_p),
                ("action", C.c_void_p), ("reward", C.c_void_p), ("terminal", C.c_void_p), ("priority", C.c_void_p),
                ("next_mask", C.c_void_p), ("tau", C.c_void_p), ("tau_prime", C.c_void_p), ("on_device", C.c_int32)]


V, I32, I64, F32, F64, PV = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double, C.POINTER(C.c_void_p)

# name -> argtypes; every function returns zoo_status unless listed in _OTHER_RESTYPE
SIGNATURES = {
    "zoo_set_device": [I32],
    "zoo_net_create": [C.POINTER(NetDesc), PV],
    "zoo_net_destroy": [V],
    "zoo_net_num_tensors": [V, C.POINTER(I32)],
    "zoo_net_tensor_size": [V, I32, C.POINTER(I64)],
    "zoo_net_num_params": [V, C.POINTER(I64)],
    "zoo_net_set_params": [V, I32, V, I64],
    "zoo_net_get_params": [V, I32, V, I64],
    "zoo_net_copy": [V, V],
    "zoo_net_forward": [V, V, I32, I32, V, I32, V],
    "zoo_net_out_dim": [V, C.POINTER(I32)],
    "zoo_opt_create": [V, I32, F64, F64, F64, F64, PV],
    "zoo_opt_destroy": [V],
    "zoo_opt_get_state": [V, I32, I32, V, I64],
    "zoo_opt_set_state": [V, I32, I32, V, I64],
    "zoo_opt_get_step": [V, C.POINTER(I64)],
    "zoo_opt_set_step": [V, I64],
    "zoo_opt_get_rng": [V, C.POINTER(C.c_uint64)],
    "zoo_opt_set_rng": [V, C.c_uint64],
    "zoo_learner_create": [C.POINTER(LearnerCfg), V, V, V, V, PV],
    "zoo_learner_destroy": [V],
    "zoo_learner_update": [V, C.POINTER(Batch), V, V],
    "zoo_learner_compute_grads": [V, C.POINTER(Batch), V, V],
    "zoo_learner_submit": [V, C.POINTER(Batch), C.c_int32],
    "zoo_learner_collect": [V, V, V],
    "zoo_learner_apply": [V],
    "zoo_learner_get_grad": [V, I32, V, I64],
    "zoo_learner_grad_buffer": [V, PV, C.POINTER(I64)],
    "zoo_learner_get_per_sample_loss": [V, V, I32],
    "zoo_learner_stream": [V, PV],
    "zoo_learner_sync": [V],
    "zoo_learner_last_launches": [V, C.POINTER(I32)],
    "zoo_learner_use_graph": [V, I32],
    "zoo_sumtree_create": [I64, PV],
    "zoo_sumtree_destroy": [V],
    "zoo_sumtree_push": [V, F32],
    "zoo_sumtree_push_many": [V, V, I64],
    "zoo_sumtree_sample": [V, V, I32, I64, I32, V, V],
    "zoo_sumtree_update": [V, V, V, I64],
    "zoo_sumtree_total": [V, C.POINTER(F32)],
    "zoo_sumtree_get": [V, V, I64, V],
    "zoo_sumtree_length": [V, C.POINTER(I64), C.POINTER(I64)],
    "zoo_sumtree_tree_len": [V, C.POINTER(I64)],
    "zoo_sumtree_export": [V, V, I64],
    "zoo_sumtree_import": [V, V, I64, I64, I64],
    "zoo_sumtree_sample_dev": [V, V, I32, I64, I32, V, V],
    "zoo_sumtree_update_dev": [V, V, V, I64],
    "zoo_sumtree_stream": [V, PV],
    "zoo_sumtree_sync": [V],
    "zoo_replay_create": [I64, I32, I32, I32, I32, F32, I32, I32, PV],
    "zoo_replay_destroy": [V],
    "zoo_replay_push": [V, V, I32, F32, C.c_uint8, F32],
    "zoo_replay_push_many": [V, V, V, V, V, V, I64],
    "zoo_replay_length": [V, C.POINTER(I64), C.POINTER(I64)],
    "zoo_replay_sample": [V, V, V, I32, I32, I32, C.POINTER(Batch), V],
    "zoo_replay_update_priorities": [V, V, V],
    "zoo_replay_invalid_count": [V, C.POINTER(C.c_uint32)],
    "zoo_replay_set_spare_draws": [V, V, V, I32, I32, I32],
    "zoo_replay_sumtree": [V, PV],
    "zoo_learner_priorities_dev": [V, PV],
    "zoo_dp_unique_id": [V],
    "zoo_dp_init": [V, I32, I32, PV],
    "zoo_dp_destroy": [V],
    "zoo_prof_begin": [V, I32],
    "zoo_prof_end": [C.POINTER(I32)],
    "zoo_prof_timeline": [I32, C.POINTER(I32)],
    "zoo_prof_timeline_get": [I32, C.c_char_p, I32, C.POINTER(C.c_uint64), C.POINTER(F32), C.POINTER(F32)],
    "zoo_prof_get": [I32, C.c_char_p, I32, C.POINTER(F32)],
    "zoo_test_gemm": [I32, I32, I32, I32, V, I32, V, I32, V, I32],
    "zoo_test_gemm_trace": [I32, I32, I32, I32, I32, I32, I32, V],
    "zoo_test_trace_begin": [I32],
    "zoo_test_trace_stop": [C.POINTER(I32)],
    "zoo_test_trace_get": [I32, C.c_char_p, I32, V],
    "zoo_test_set_tc_stages": [I32],
    "zoo_test_gemm_bench": [I32, I32, I32, I32, I32, I32, I32, I32, I32, C.POINTER(F32)],
    "zoo_test_net_activation": [V, I32, I32, I64, V, I64, C.POINTER(I32)],
}
_OTHER_RESTYPE = {"zoo_last_error": ([], C.c_char_p), "zoo_version": ([], I32), "zoo_device_arch": ([], I32)}

_lib = None


def lib():
    """Loads libzoo_sm100.so once.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZooError(2, "%s not found -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(or ./build.sh); there is no CPU fallback" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, argt in SIGNATURES.items():
            if name.startswith("zoo_test_") and os.environ.get("ZOO_SM100_LIB") and not hasattr(l, name):
                continue  # an older build under A/B may lack a newer test hook; product entry points stay mandatory
            f = getattr(l, name)
            f.argtypes = argt
            f.restype = I32
        for name, (argt, rest) in _OTHER_RESTYPE.items():
            f = getattr(l, name)
            f.argtypes = argt
            f.restype = rest
        _lib = l
    return _lib


def check(status):
    if status != 0:
        raise ZooError(status, lib().zoo_last_error().decode(errors="replace"))


def ptr(a):
    """numpy array (C-contiguous) -> void*; None -> NULL."""
    return None if a is None else a.ctypes.data
