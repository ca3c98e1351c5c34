"""CPU, world_size 2, gloo: the flat-bucket gradient all-reduce (monodetr_b200/ddp.py) -- host logic of the N>1 path."""
import os
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, store_path, q, views):
    # file rendezvous: no port to race for, nothing to resolve (gloo itself still binds 127.0.0.1)
    os.environ["GLOO_SOCKET_IFNAME"] = "lo"
    dist.init_process_group("gloo", init_method=f"file://{store_path}", rank=rank, world_size=world)
    from monodetr_b200.ddp import FlatGradBucket, broadcast_parameters
    torch.manual_seed(rank)                       # different init per rank -> broadcast must equalise
    model = torch.nn.ModuleDict({
        "a": torch.nn.Linear(8, 4),
        "sa_v_proj": torch.nn.Linear(4, 4),       # never-used tensors stay out of the bucket
        "b": torch.nn.Linear(4, 2),
    })
    broadcast_parameters(model)
    w0 = model["a"].weight.detach().clone()
    bucket = FlatGradBucket(model, views=views)
    assert bucket.param_numel == sum(p.numel() for n, p in model.named_parameters() if "sa_v_proj" not in n)
    assert bucket.numel >= bucket.param_numel and all(o % bucket.ALIGN == 0 for o in bucket.offsets)
    assert model["sa_v_proj"].weight.grad is None
    bucket.zero()
    x = torch.full((3, 8), float(rank + 1))
    model["b"](model["a"](x)).sum().backward()   # views=True: accumulates straight into the flat views
    local = torch.cat([p.grad.reshape(-1) for p in bucket.params]).clone()
    assert local.abs().sum() > 0
    bucket.all_reduce()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = sum(gathered) / world
    reduced = torch.cat([v.reshape(-1) for v in bucket.views])          # (the flat buffer also holds alignment padding)
    ok = torch.allclose(reduced, expect, atol=1e-6) and model["a"].weight.grad.data_ptr() == bucket.flat.data_ptr()
    if not views:
        # CUDA-graph mode: a "replay" rewrites the captured gradient tensors in place; after all_reduce() an optimizer
        # reading p.grad must see the rank mean, and the captured sources must stay what the next replay writes
        bucket.zero()
        model["b"](model["a"](x)).sum().backward()
        bucket.freeze_sources()
        srcs = [g for g in bucket.static_grads]
        for step in range(2):
            for g in srcs:                            # the replay: same tensors, new values
                g.copy_(torch.full_like(g, float((rank + 1) * (step + 1))))
            bucket.all_reduce()
            want = (1 + 2) * (step + 1) / world
            for p in bucket.params:
                ok = ok and torch.allclose(p.grad, torch.full_like(p.grad, want))
            ok = ok and all(a is b for a, b in zip(srcs, bucket.static_grads))
        try:                                          # freezing sources that alias the bucket is refused, not silently wrong
            bucket.freeze_sources()
            ok = False
        except RuntimeError:
            pass
    q.put((rank, bool(ok), w0))
    dist.barrier()                                # nobody tears its sockets down while the peer is still communicating
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("views", [False, True])
def test_flat_bucket_allreduce_world2(views):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    store_path = os.path.join(tempfile.mkdtemp(prefix="mdb_gloo_"), "store")
    procs = [ctx.Process(target=_worker, args=(r, 2, store_path, q, views)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]      # a cold `import torch` in a fresh container can take a minute
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert torch.equal(res[0][2], res[1][2])      # parameters were broadcast from rank 0
