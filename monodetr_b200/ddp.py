"""Data-parallel gradient exchange: ONE NCCL all-reduce per step over a flat fp32 bucket.

Every gradient-receiving parameter's `.grad` is a view into one contiguous buffer, so backward accumulates
straight into the bucket and the step ends with a single `all_reduce(bucket) / world_size` over NVLink.  The 15
tensors that never receive a gradient (sa_v_proj, query_scale, ref_point_head, label_enc -- SURVEY.md appendix C.2)
are left out of the bucket (their .grad stays None, exactly as in the reference).
"""
import torch
import torch.distributed as dist

_NEVER_USED = ("sa_v_proj", "decoder.query_scale", "decoder.ref_point_head", "label_enc")


class FlatGradBucket:
    def __init__(self, model):
        self.params = [p for n, p in model.named_parameters() if p.requires_grad and not any(s in n for s in _NEVER_USED)]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.numel = n

    def zero(self):
        self.flat.zero_()

    def all_reduce(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())


def broadcast_parameters(model, src=0):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src)
