"""monodetr_b200 -- B200 (sm_100a) implementation of MonoDETR's forward/backward hot path behind the reference's
own interfaces: `build_monodetr(cfg)` / `MonoDETR.forward`, the `MSDeformAttn` module and the
`MultiScaleDeformableAttention` extension functions.  Kernels live in csrc/ behind the C ABI of
include/monodetr_b200.h (libmonodetr_b200.so, built by `python -m monodetr_b200.build`)."""


def build_monodetr(cfg, criterion_builder=None):
    """Mirror of lib/models/monodetr/__init__.py:4-5."""
    from .monodetr import build
    return build(cfg, criterion_builder)


def set_deterministic(on=True):
    """Reproducible accumulation for the two large scatter sites (`mdb_set_deterministic`, include/monodetr_b200.h): the
    MSDeformAttn value gradient is accumulated in a fixed order (the reference's kernel and the default path here scatter
    with atomics, ms_deform_im2col_cuda.cuh:125-152) and weight gradients run without split-K.  A test / debugging mode
    (much slower); set it before the forward pass.  Returns the previous setting."""
    from . import _lib
    prev = bool(_lib.lib().mdb_get_deterministic())
    _lib.check(_lib.lib().mdb_set_deterministic(1 if on else 0), "set_deterministic")
    return prev


def is_deterministic():
    from . import _lib
    return bool(_lib.lib().mdb_get_deterministic())


__all__ = ["build_monodetr", "set_deterministic", "is_deterministic"]
