"""Import the UNMODIFIED reference from /root/reference on CPU (golden-vector generation only).

Nothing under monodetr_b200/ imports this file.  It installs the in-memory shims SURVEY.md 8(c)
lists (no reference file is edited or copied):
  1. torch.nn.modules.linear._LinearWithBias           (ops/modules/ms_deform_attn.py:34 version test)
  2. torch._overrides                                   (ops/modules/ms_deform_attn.py:55-58)
  3. sys.modules['MultiScaleDeformableAttention']       (ops/functions/ms_deform_attn_func.py:18) served
     by the reference's own ms_deform_attn_core_pytorch (same file :41-61) + autograd
  4. backbone.is_main_process -> False                  (backbone.py:102, no weight download)
  5. torch.cuda.current_device -> 0                     (depth_predictor/ddn_loss/ddn_loss.py:32)
"""
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("MONODETR_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "models", "monodetr"))


def install():
    """Returns the imported reference package `lib.models.monodetr` (module)."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    import torch.nn.modules.linear as _lin
    if not hasattr(_lin, "_LinearWithBias"):
        _lin._LinearWithBias = _lin.NonDynamicallyQuantizableLinear
    if "torch._overrides" not in sys.modules:
        ov = types.ModuleType("torch._overrides")
        ov.has_torch_function = torch.overrides.has_torch_function
        ov.handle_torch_function = torch.overrides.handle_torch_function
        sys.modules["torch._overrides"] = ov
        torch._overrides = ov

    if "MultiScaleDeformableAttention" not in sys.modules or \
            not getattr(sys.modules["MultiScaleDeformableAttention"], "_is_ref_shim", False):
        shim = types.ModuleType("MultiScaleDeformableAttention")
        shim._is_ref_shim = True

        def _core():
            from lib.models.monodetr.ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch
            return ms_deform_attn_core_pytorch

        def ms_deform_attn_forward(value, shapes, lsi, loc, attn, im2col_step):
            with torch.no_grad():
                return _core()(value, shapes, loc, attn)

        def ms_deform_attn_backward(value, shapes, lsi, loc, attn, grad_output, im2col_step):
            with torch.enable_grad():
                v = value.detach().requires_grad_(True)
                lo = loc.detach().requires_grad_(True)
                a = attn.detach().requires_grad_(True)
                out = _core()(v, shapes, lo, a)
                gv, gl, ga = torch.autograd.grad(out, (v, lo, a), grad_output)
            return gv, gl, ga

        shim.ms_deform_attn_forward = ms_deform_attn_forward
        shim.ms_deform_attn_backward = ms_deform_attn_backward
        sys.modules["MultiScaleDeformableAttention"] = shim

    torch.cuda.current_device = lambda: 0

    import lib.models.monodetr.backbone as _bb
    _bb.is_main_process = lambda: False
    import lib.models.monodetr as ref_pkg
    return ref_pkg


def load_cfg():
    import yaml
    with open(os.path.join(REF_ROOT, "configs", "monodetr.yaml")) as f:
        cfg = yaml.load(f, Loader=yaml.Loader)
    cfg["model"]["device"] = "cpu"
    return cfg
