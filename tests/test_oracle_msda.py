"""Pin the C oracle (oracle/msda_oracle.c) against vectors produced by the reference itself.

Fixtures: tests/golden/msda_*.npz, written by tools/gen_golden_msda.py from the reference's
ms_deform_attn_core_pytorch (ops/functions/ms_deform_attn_func.py:41-61) + autograd; the `reftest_*`
cases reproduce the generator of the reference's ops/test.py:21-37.  CPU only.
"""
import glob
import os

import numpy as np
import pytest

from oracle import msda as oracle_msda

CASES = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "msda_*.npz")))


def _tol(dtype):
    # fp64: the reference's own check is torch.allclose default (ops/test.py:40); fp32: rtol 1e-2 atol 1e-3
    # (ops/test.py:56).  We hold the oracle to far tighter bounds.
    return (1e-10, 1e-12) if dtype == np.float64 else (1e-4, 1e-5)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, f"msda_{name}.npz"))
    rtol, atol = _tol(g["value"].dtype)
    out = oracle_msda.msda_forward(g["value"], g["shapes"], g["lsi"], g["loc"], g["attn"])
    np.testing.assert_allclose(out, g["out"], rtol=rtol, atol=atol)
    gv, gl, ga = oracle_msda.msda_backward(g["value"], g["shapes"], g["lsi"], g["loc"], g["attn"], g["grad_out"])
    scale = lambda a: max(1.0, float(np.abs(a).max()))
    np.testing.assert_allclose(gv, g["grad_value"], rtol=rtol, atol=atol * scale(g["grad_value"]))
    np.testing.assert_allclose(gl, g["grad_loc"], rtol=rtol, atol=atol * scale(g["grad_loc"]))
    np.testing.assert_allclose(ga, g["grad_attn"], rtol=rtol, atol=atol * scale(g["grad_attn"]))


def test_fixture_inventory():
    assert {"reftest_f32", "reftest_f64", "d32_f32"} <= set(CASES)


def test_oracle_empty_and_outside():
    """All sample points outside (-1,H)x(-1,W) -> exact zeros everywhere (cuh:288, :365-368)."""
    shapes = np.array([[3, 5]], dtype=np.int64)
    lsi = np.zeros(1, dtype=np.int64)
    value = np.random.RandomState(0).randn(1, 15, 2, 4).astype(np.float32)
    loc = np.full((1, 3, 2, 1, 2, 2), 7.0, dtype=np.float32)
    loc[..., 0, :] = -3.0
    attn = np.full((1, 3, 2, 1, 2), 0.5, dtype=np.float32)
    out = oracle_msda.msda_forward(value, shapes, lsi, loc, attn)
    assert not out.any()
    gv, gl, ga = oracle_msda.msda_backward(value, shapes, lsi, loc, attn, np.ones_like(out))
    assert not gv.any() and not gl.any() and not ga.any()


@pytest.mark.parametrize("name", ["reftest_f64", "d32_f32", "d71_f64"])
def test_torch_port_matches_golden(name, golden_dir):
    """oracle/msda_torch.py (the CPU-baseline port of ms_deform_attn_core_pytorch) against the same vectors."""
    import torch
    from oracle.msda_torch import msda_core_torch
    g = np.load(os.path.join(golden_dir, f"msda_{name}.npz"))
    rtol, atol = _tol(g["value"].dtype)
    v, lo, a = (torch.from_numpy(g[k]).requires_grad_(True) for k in ("value", "loc", "attn"))
    out = msda_core_torch(v, torch.from_numpy(g["shapes"]), lo, a)
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=rtol, atol=atol)
    gv, gl, ga = torch.autograd.grad(out, (v, lo, a), torch.from_numpy(g["grad_out"]))
    np.testing.assert_allclose(gv.numpy(), g["grad_value"], rtol=rtol, atol=atol * 10)
    np.testing.assert_allclose(gl.numpy(), g["grad_loc"], rtol=rtol, atol=atol * 10)
    np.testing.assert_allclose(ga.numpy(), g["grad_attn"], rtol=rtol, atol=atol * 10)
