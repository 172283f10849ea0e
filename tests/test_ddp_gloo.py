"""world_size-2 gloo test (CPU) of the data-parallel host logic: flat-gradient all-reduce == single-process gradients
of the concatenated batch, and parameters stay identical after an optimizer step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.LayerNorm(32), torch.nn.Linear(32, 5))


def _worker(rank, world, port, out, bucketed=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ofq_b200.ddp import BucketedGradAllReduce, FlatGradAllReduce, broadcast_parameters
    model = _model()
    if rank == 1:                                   # deliberately different start: the broadcast must fix it
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    broadcast_parameters(model, 0)
    # tiny buckets: the four parameter tensors fall into several of them
    ddp = BucketedGradAllReduce(model, world, bucket_mb=0.001) if bucketed else FlatGradAllReduce(model.parameters(), world)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(100)
    x = torch.randn(8, 16, generator=g)
    y = torch.randint(0, 5, (8,), generator=g)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    for _ in range(2):
        ddp.zero()
        loss = torch.nn.functional.cross_entropy(model(xs), ys)
        ddp.scale_loss(loss).backward()
        ddp.reduce()
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(ddp.params, getattr(ddp, 'views', [p.grad for p in ddp.params])))
        grads = torch.cat([p.grad.flatten() for p in model.parameters()])   # (the flat buffer pads every slice to 128 bytes)
        opt.step()
    out[rank] = (grads, torch.cat([p.detach().flatten() for p in model.parameters()]))
    dist.destroy_process_group()


@pytest.mark.parametrize("bucketed", [False, True])
def test_flat_allreduce_matches_single_process(bucketed):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out, bucketed), nprocs=world, join=True)
    # single-process reference on the full batch
    model = _model()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(100)
    x = torch.randn(8, 16, generator=g)
    y = torch.randint(0, 5, (8,), generator=g)
    for _ in range(2):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(model(x), y).backward()
        ref_g = torch.cat([p.grad.flatten() for p in model.parameters()])
        opt.step()
    ref_p = torch.cat([p.detach().flatten() for p in model.parameters()])
    for r in range(world):
        gr, pr = out[r]
        assert torch.allclose(gr, ref_g, rtol=1e-5, atol=1e-7)      # mean of per-rank means == full-batch mean
        assert torch.allclose(pr, ref_p, rtol=1e-5, atol=1e-6)
    assert torch.equal(out[0][1], out[1][1])                        # replicas stay bit-identical
