"""N-GPU data-parallel gradients == 1-GPU gradients of the concatenated batch, on a real multi-GPU box:
DeiT-T width (192, 3 heads) W2A2 QKR, two blocks (low-bit QAT is chaotic in the last bit - see test_gpu_fullsize_parity.py -
and the two runs quantize with step sizes whose last bit depends on the local batch size, lsq.py:6-9), global batch 8 split
over 2 ranks, NCCL, ofq_b200.ddp (flat and bucketed / overlapped exchange).
Skipped on a single-GPU box (the driver's round-end `-m gpu` run); run with `gpurun --gpus 2`, result recorded in profiles/.

Every gradient must agree to the fp16-operand tolerance of the backward (1e-3). The LSQ step-size gradients carry the
reference's own batch dependence: their gradient scale is 1/sqrt(thd_pos * elements-per-scale) of the LOCAL batch
(lsq.py:582-591): 1/sqrt(B/W) instead of 1/sqrt(B), so the reduced step-size gradient is sqrt(W) times the 1-GPU value (torch
DDP around the reference behaves identically); the test accounts for exactly that factor."""
import os
import socket

import pytest
import torch
import torch.nn.functional as F

import fullsize_common as FC
from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
DEPTH = 2


def _model():
    import ofq_b200.quantization as Q
    from ofq_b200.host.deit import DistilledVisionTransformer
    model = DistilledVisionTransformer(num_classes=1000, embed_dim=192, depth=DEPTH, num_heads=3)
    model = Q.replace_module_by_qmodule_deit(model, Q.make_qconfigs(FC.deit_names(DEPTH), 2, 2), pretrained_initialized=True,
                                             qk_reparam=True)
    values, _ = FC.fill_state(model.state_dict())
    FC.apply_state(model, values)
    model = model.cuda()
    model.eval()
    with torch.no_grad():
        model(FC.det_images().cuda())             # setup_alpha on the full batch: identical step sizes everywhere
    return model.train()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _grads(model, img, labels, ddp=None):
    (cls, dst), _ = model(img)
    loss = F.cross_entropy(cls, labels) + F.cross_entropy(dst, labels)
    (ddp.scale_loss(loss) if ddp is not None else loss).backward()
    if ddp is not None:
        ddp.reduce()
    return {n: p.grad.detach().float().cpu().clone() for n, p in model.named_parameters() if p.grad is not None}


def _worker(rank, world, port, mode, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ofq_b200 import ddp as D
    model = _model()
    D.broadcast_parameters(model, 0)
    per = FC.BATCH // world
    img = FC.det_images()[rank * per:(rank + 1) * per].cuda()
    labels = FC.det_labels(FC.BATCH, 1000)[rank * per:(rank + 1) * per].cuda()
    if mode == "flat":
        ddp = D.FlatGradAllReduce(model.parameters(), world)
    else:
        ddp = D.BucketedGradAllReduce(model, world)
    ddp.zero()
    out[rank] = _grads(model, img, labels, ddp)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["flat", "bucketed"])
def test_two_gpu_gradients_match_single_gpu(mode):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), mode, out), nprocs=world, join=True)
    model = _model()
    ref = _grads(model, FC.det_images().cuda(), FC.det_labels(FC.BATCH, 1000).cuda())
    gmax = max(v.abs().max().item() for v in ref.values())
    worst = 0.0
    bad = []
    for n, r in ref.items():
        for rank in range(world):
            mine = out[rank][n]
            if n.endswith(".s") and "lsqw_fn" not in n:          # activation step sizes: local-batch gradient scale
                mine = mine / (world ** 0.5)
            e = rel_err(mine, r)
            ok = e < 2e-3 or (mine - r).abs().max().item() <= 2e-5 * gmax
            if not ok:
                bad.append((n, rank, e))
            worst = max(worst, e if (mine - r).abs().max().item() > 2e-5 * gmax else 0.0)
        assert torch.equal(out[0][n], out[1][n]), n          # both ranks hold the same reduced gradient
    print(f"2-GPU vs 1-GPU gradient parity ({mode}): worst rel err {worst:.2e}; outside tolerance: {bad[:6]}")
    assert not bad, bad[:10]
