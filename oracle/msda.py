"""ctypes binding of oracle/msda_oracle.c (checker only; see that file's header for citations)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_msda.so")
_SRC = os.path.join(_HERE, "msda_oracle.c")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", _SO, _SRC, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(value, shapes, lsi, loc, attn):
    dt = value.dtype
    assert dt in (np.float32, np.float64)
    value = np.ascontiguousarray(value)
    loc = np.ascontiguousarray(loc, dtype=dt)
    attn = np.ascontiguousarray(attn, dtype=dt)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    lsi = np.ascontiguousarray(lsi, dtype=np.int64)
    B, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    return value, shapes, lsi, loc, attn, (B, S, M, D, L, Lq, P), ("f32" if dt == np.float32 else "f64")


def msda_forward(value, shapes, lsi, loc, attn):
    """numpy in -> numpy out (B, Lq, M*D)."""
    value, shapes, lsi, loc, attn, dims, suf = _prep(value, shapes, lsi, loc, attn)
    B, S, M, D, L, Lq, P = dims
    out = np.empty((B, Lq, M * D), dtype=value.dtype)
    fn = getattr(_load(), f"oracle_msda_fwd_{suf}")
    fn(_ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(attn),
       *[ctypes.c_int(v) for v in dims], _ptr(out))
    return out


def msda_backward(value, shapes, lsi, loc, attn, grad_out):
    value, shapes, lsi, loc, attn, dims, suf = _prep(value, shapes, lsi, loc, attn)
    grad_out = np.ascontiguousarray(grad_out, dtype=value.dtype)
    gv = np.empty_like(value)
    gl = np.empty_like(loc)
    ga = np.empty_like(attn)
    fn = getattr(_load(), f"oracle_msda_bwd_{suf}")
    fn(_ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(attn), _ptr(grad_out),
       *[ctypes.c_int(v) for v in dims], _ptr(gv), _ptr(gl), _ptr(ga))
    return gv, gl, ga
