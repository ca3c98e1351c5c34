"""ctypes binding of oracle/_ref/libref_msda.so -- the REFERENCE's own CUDA kernels recompiled for
sm_100a (see oracle/ref_msda_driver.cu).  Checker / "kernel to beat" only; torch CUDA tensors in/out."""
import ctypes
import os

import torch

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_msda.so")
_lib = None


def available() -> bool:
    return os.path.exists(_SO)


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_SO)
    return _lib


def _dims(value, loc):
    B, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    return [ctypes.c_int(v) for v in (B, S, M, D, L, Lq, P)]


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def forward(value, shapes, lsi, loc, attn):
    suf = "f32" if value.dtype == torch.float32 else "f64"
    B, S, M, D = value.shape
    out = torch.empty((B, loc.shape[1], M * D), dtype=value.dtype, device=value.device)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = getattr(_load(), f"ref_msda_forward_{suf}")(_p(value), _p(shapes), _p(lsi), _p(loc), _p(attn),
                                                      *_dims(value, loc), _p(out), st)
    assert rc == 0, rc
    return out


def backward(value, shapes, lsi, loc, attn, grad_out):
    suf = "f32" if value.dtype == torch.float32 else "f64"
    gv, gl, ga = torch.empty_like(value), torch.empty_like(loc), torch.empty_like(attn)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = getattr(_load(), f"ref_msda_backward_{suf}")(_p(value), _p(shapes), _p(lsi), _p(loc), _p(attn),
                                                       _p(grad_out), *_dims(value, loc), _p(gv), _p(gl), _p(ga), st)
    assert rc == 0, rc
    return gv, gl, ga
