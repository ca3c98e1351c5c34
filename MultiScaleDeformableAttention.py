"""Top-level module with the name the reference imports
(`import MultiScaleDeformableAttention as MSDA`, lib/models/monodetr/ops/functions/ms_deform_attn_func.py:18).
Put the repo root on PYTHONPATH and the reference's MSDeformAttnFunction runs on the sm_100a kernels unchanged.
"""
from monodetr_b200.msda import ms_deform_attn_backward, ms_deform_attn_forward  # noqa: F401
