"""monodetr_b200 -- B200 (sm_100a) implementation of MonoDETR's forward/backward hot path behind the reference's
own interfaces: `build_monodetr(cfg)` / `MonoDETR.forward`, the `MSDeformAttn` module and the
`MultiScaleDeformableAttention` extension functions.  Kernels live in csrc/ behind the C ABI of
include/monodetr_b200.h (libmonodetr_b200.so, built by `python -m monodetr_b200.build`)."""


def build_monodetr(cfg, criterion_builder=None):
    """Mirror of lib/models/monodetr/__init__.py:4-5."""
    from .monodetr import build
    return build(cfg, criterion_builder)


__all__ = ["build_monodetr"]
