"""Checkpoints: the reference saves ``BSON.@save "policy.bson" policy`` after ``cpu(p)``
(src/experiments/atari/Dopamine_DQN_Atari.jl:123-126).  This module writes the learner's state -- online / target parameters,
optimiser state and step, the learner's counters and hyper-parameters -- as one BSON document (bsonspec.org v1.1) and reads it
back into a learner: parameters, optimiser state, ADAM's beta^t (rebuilt by the same repeated fp64 multiplication the run
carried) and the counter of the on-device tau stream are restored bit-exactly.

Arrays are lowered the way BSON.jl lowers ``Array{Float32,N}``: ``{tag: "array", type: {tag: "datatype", params: [], name:
["Core", "Float32"]}, size: [dims...], data: <binary>}`` with the *Julia* dimensions and column-major bytes (= the ABI bytes), so
a maintainer can ``BSON.load(path)[:policy]`` and ``Flux.loadparams!`` from it.  That lowering is restated from BSON.jl's
documentation; ``julia`` is not available in this image, so compatibility with BSON.jl is UNVERIFIED -- what the tests pin is the
round trip through this codec and the bsonspec framing.  The document is a plain dictionary, not BSON.jl's type-tagged dump of a
``QBasedPolicy`` struct: rebuilding the Julia object is the glue's job."""
from __future__ import annotations

import struct

import numpy as np

_JL_TYPES = {"float32": "Float32", "float64": "Float64", "int32": "Int32", "int64": "Int64", "uint8": "UInt8"}
_NP_TYPES = {v: k for k, v in _JL_TYPES.items()}


# ---- bsonspec subset: double, string, document, array, binary(0), bool, null, int32, int64 ----------------------------------
def _cstr(s):
    b = s.encode("utf-8")
    if b"\x00" in b:
        raise ValueError("BSON keys cannot contain NUL")
    return b + b"\x00"


def _enc_value(v):
    if isinstance(v, bool):
        return 0x08, b"\x01" if v else b"\x00"
    if isinstance(v, (int, np.integer)):
        v = int(v)
        if -(1 << 31) <= v < (1 << 31):
            return 0x10, struct.pack("<i", v)
        return 0x12, struct.pack("<q", v)
    if isinstance(v, (float, np.floating)):
        return 0x01, struct.pack("<d", float(v))
    if isinstance(v, str):
        b = v.encode("utf-8") + b"\x00"
        return 0x02, struct.pack("<i", len(b)) + b
    if isinstance(v, (bytes, bytearray, memoryview)):
        b = bytes(v)
        return 0x05, struct.pack("<i", len(b)) + b"\x00" + b
    if v is None:
        return 0x0A, b""
    if isinstance(v, np.ndarray):
        return 0x03, bson_dumps(lower_array(v))
    if isinstance(v, dict):
        return 0x03, bson_dumps(v)
    if isinstance(v, (list, tuple)):
        return 0x04, bson_dumps({str(i): x for i, x in enumerate(v)})
    raise TypeError("cannot encode %r as BSON" % type(v))


def bson_dumps(doc):
    body = b""
    for k, v in doc.items():
        t, payload = _enc_value(v)
        body += bytes([t]) + _cstr(str(k)) + payload
    return struct.pack("<i", len(body) + 5) + body + b"\x00"


def _dec_doc(buf, off, as_list):
    (n,) = struct.unpack_from("<i", buf, off)
    end = off + n
    if n < 5 or end > len(buf) or buf[end - 1] != 0:
        raise ValueError("corrupt BSON document")
    p = off + 4
    out = {}
    while p < end - 1:
        t = buf[p]
        q = buf.index(b"\x00", p + 1)
        key = buf[p + 1:q].decode("utf-8")
        p = q + 1
        if t == 0x01:
            (val,) = struct.unpack_from("<d", buf, p); p += 8
        elif t == 0x02:
            (ln,) = struct.unpack_from("<i", buf, p)
            val = buf[p + 4:p + 4 + ln - 1].decode("utf-8"); p += 4 + ln
        elif t in (0x03, 0x04):
            val, p = _dec_doc(buf, p, t == 0x04)
        elif t == 0x05:
            (ln,) = struct.unpack_from("<i", buf, p)
            val = bytes(buf[p + 5:p + 5 + ln]); p += 5 + ln
        elif t == 0x08:
            val = buf[p] != 0; p += 1
        elif t == 0x0A:
            val = None
        elif t == 0x10:
            (val,) = struct.unpack_from("<i", buf, p); p += 4
        elif t == 0x12:
            (val,) = struct.unpack_from("<q", buf, p); p += 8
        else:
            raise ValueError("unsupported BSON element type 0x%02x" % t)
        out[key] = val
    if as_list:
        return [out[k] for k in sorted(out, key=int)], end
    if out.get("tag") == "array" and "data" in out:
        return raise_array(out), end
    return out, end


def bson_loads(buf):
    doc, end = _dec_doc(bytes(buf), 0, False)
    if end != len(buf):
        raise ValueError("trailing bytes after the BSON document")
    return doc


# ---- arrays: logical numpy (Dense (out,in), conv OIHW) <-> Julia dims + column-major bytes ------------------------------------
def lower_array(a):
    a = np.asarray(a)
    name = _JL_TYPES[a.dtype.name]
    if a.ndim == 2:  # Dense weight: Julia (out, in), column-major == row-major [in][out]
        size, data = [int(a.shape[0]), int(a.shape[1])], np.ascontiguousarray(a.T).tobytes()
    else:            # vectors; conv OIHW row-major == Julia (kw, kh, cin, cout) column-major
        size, data = [int(x) for x in a.shape[::-1]], np.ascontiguousarray(a).tobytes()
    return {"tag": "array", "type": {"tag": "datatype", "params": [], "name": ["Core", name]}, "size": size, "data": data}


def raise_array(d):
    dt = np.dtype(_NP_TYPES[d["type"]["name"][-1]])
    size = [int(x) for x in d["size"]]
    flat = np.frombuffer(d["data"], dt)
    if int(np.prod(size)) != flat.size:
        raise ValueError("array payload does not match its size")
    if len(size) == 2:
        return np.ascontiguousarray(flat.reshape(size[1], size[0]).T)
    return flat.reshape(size[::-1]).copy()


# ---- learner <-> document -------------------------------------------------------------------------------------------------------
_HYPER = ("gamma", "batch_size", "update_horizon", "min_replay_history", "update_freq", "target_update_freq", "update_step",
          "is_enable_double_DQN", "n_actions", "n_atoms", "Vmax", "Vmin", "N", "N_prime", "K", "kappa", "n_quantile", "ensemble_num")


def _approx_doc(app):
    import ctypes as C

    from . import _lib as L
    d = {"precision": app.precision, "obs_shape": list(app.obs_shape), "params": app.params()}
    if app.opt_h:
        o = app.optimizer
        t = C.c_int64()
        L.check(L.lib().zoo_opt_get_step(app.opt_h, C.byref(t)))
        n_slots = 2 if o.kind == L.OPT_ADAM else 1
        rng = C.c_uint64()
        L.check(L.lib().zoo_opt_get_rng(app.opt_h, C.byref(rng)))
        d["optimizer"] = {"kind": "ADAM" if o.kind == L.OPT_ADAM else "RMSProp", "eta": o.eta, "a": o.a, "b": o.b, "t": int(t.value),
                          "rng": int(rng.value), "state": [app.opt_state(s) for s in range(n_slots)]}
    return d


def policy_document(learner):
    """The dictionary ``save_policy`` writes (numpy arrays in the logical layout of ``NeuralNetworkApproximator.params``)."""
    learner.sync()
    lr = {"type": type(learner).__name__, "loss": float(learner.loss), "approximator": _approx_doc(learner.approximator)}
    if learner.target_approximator is not None:
        lr["target_approximator"] = _approx_doc(learner.target_approximator)
    for k in _HYPER:
        if hasattr(learner, k) and isinstance(getattr(learner, k), (bool, int, float, np.integer, np.floating)):
            lr[k] = getattr(learner, k)
    return {"policy": {"learner": lr}}


def save_policy(path, learner):
    """``BSON.@save path policy`` (Dopamine_DQN_Atari.jl:123-126)."""
    with open(path, "wb") as f:
        f.write(bson_dumps(policy_document(learner)))


def _load_approx(app, d):
    import ctypes as C

    from . import _lib as L
    from .nn import _abi
    if [tuple(np.shape(p)) for p in d["params"]] != [tuple(s) for s in app.shapes]:
        raise ValueError("checkpoint parameters do not match the approximator's model")
    app.set_params(d["params"])
    o = d.get("optimizer")
    if o is not None and app.opt_h:
        want = "ADAM" if app.optimizer.kind == L.OPT_ADAM else "RMSProp"
        if o["kind"] != want:
            raise ValueError("checkpoint optimiser is %s, the approximator's is %s" % (o["kind"], want))
        for slot, tensors in enumerate(o["state"]):
            for i, p in enumerate(tensors):
                a = _abi(p)
                L.check(L.lib().zoo_opt_set_state(app.opt_h, slot, i, a.ctypes.data, a.size))
        L.check(L.lib().zoo_opt_set_step(app.opt_h, C.c_int64(int(o["t"]))))  # also rebuilds beta^t by repeated multiplication
        L.check(L.lib().zoo_opt_set_rng(app.opt_h, C.c_uint64(int(o.get("rng", 0)))))


def load_policy(path, learner):
    """``BSON.@load path policy`` into an existing learner of the same architecture: parameters, optimiser state, counters."""
    with open(path, "rb") as f:
        lr = bson_loads(f.read())["policy"]["learner"]
    if lr["type"] != type(learner).__name__:
        raise ValueError("checkpoint holds a %s, not a %s" % (lr["type"], type(learner).__name__))
    learner.sync()
    _load_approx(learner.approximator, lr["approximator"])
    if learner.target_approximator is not None:
        _load_approx(learner.target_approximator, lr["target_approximator"])
    for k in ("update_step",):
        if k in lr and hasattr(learner, k):
            setattr(learner, k, lr[k])
    learner.loss = np.float32(lr.get("loss", 0.0))
    return learner
