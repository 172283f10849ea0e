"""Host-side arithmetic that the kernels rely on, checked without a GPU."""
import random


def make_fastdiv(d):
    """Mirror of make_fastdiv in ofq_b200/csrc/gemm_tc.cu: q = (umulhi(n, mul) + n) >> shr for 0 <= n < 2^31."""
    d = max(d, 1)
    s = 0
    while s < 31 and (1 << s) < d:
        s += 1
    mul = ((1 << 32) * ((1 << s) - d)) // d + 1
    assert mul < (1 << 32)
    return d, mul, s


def fastdiv(n, f):
    d, mul, s = f
    return ((((n * mul) >> 32) + n) & 0xFFFFFFFF) >> s


def test_multiply_high_division_is_exact_for_31_bit_operands():
    """Tile decode and the wrapped epilogue-vector indices of the GEMM engine divide by run-time constants (tile counts,
    batch counts, split counts, vector periods incl. the 'no wrap' period 2^31 - 1) with one multiply-high and a shift."""
    rng = random.Random(1)
    divisors = [1, 2, 3, 5, 6, 7, 9, 12, 13, 33, 64, 99, 128, 148, 197, 198, 384, 1188, 1782, 25344, (1 << 30), (1 << 30) + 1,
                0x7FFFFFFF] + [rng.randrange(1, 1 << 31) for _ in range(500)]
    for d in divisors:
        f = make_fastdiv(d)
        probes = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, (1 << 31) - 2, (1 << 31) - 1] + [rng.randrange(0, 1 << 31) for _ in range(100)]
        for n in probes:
            if 0 <= n < (1 << 31):
                assert fastdiv(n, f) == n // d, (n, d)


def choose_splits(tiles, kblocks, slots=148, fixed=6.0):
    """Mirror of the splits = 0 rule of ofq_gemm_ex: whole rounds of one work item per SM, minimal rounds x (k-blocks + fixed)."""
    best, choice = -1.0, 1
    sp = 1
    while sp <= 64 and sp <= max(kblocks // 4, 1):
        rounds = (tiles * sp + slots - 1) // slots
        cost = rounds * ((kblocks + sp - 1) // sp + fixed)
        if best < 0 or cost < best * 0.98:
            best, choice = cost, sp
        sp += 1
    return choice


def test_auto_split_k_fills_whole_rounds_for_the_weight_gradient_shapes():
    """DeiT-S batch 128: K = 25344 tokens = 396 k-blocks of 64; tiles = row blocks x column tiles the kernel dispatches."""
    for tiles, expect in ((36, 4), (18, 8), (24, 6), (6, 24)):       # dW_qk (18 x 2), fc1 (3 x 6), fc2 (12 x 2), v / proj (3 x 2)
        sp = choose_splits(tiles, 396)
        assert sp == expect and tiles * sp <= 148


def test_dispatcher_ops_are_registered_with_schemas_and_no_cpu_kernel():
    """the kernels are PyTorch dispatcher operators (`torch.ops.ofq_b200.*`); CUDA-only (calling one with CPU tensors
    raises: no CPU fallback in the product path), with fake implementations for shape propagation."""
    import torch
    import ofq_b200.torch_ops as T
    from torch._subclasses.fake_tensor import FakeTensorMode
    for n in T.OPS:
        schema = str(getattr(torch.ops.ofq_b200, n).default._schema)
        assert schema.startswith(f"ofq_b200::{n}("), schema
    assert "Tensor(a0!) p" in str(torch.ops.ofq_b200.cga_adamw_step.default._schema)        # in-place update is declared
    with FakeTensorMode():
        w = torch.empty(8, 16, device="cuda")
        codes, colscale = torch.ops.ofq_b200.statsq_codes(w, 2)
        assert codes.shape == (8, 16) and codes.dtype == torch.int8 and colscale.shape == (8,)
        out = torch.ops.ofq_b200.qgemm_fwd(codes, colscale, codes, colscale, None)
        assert out.shape == (8, 8) and out.dtype == torch.float32
    import pytest
    with pytest.raises(NotImplementedError):
        torch.ops.ofq_b200.cga_mask(torch.randn(4, 4), 2, 0.005)


def test_flat_gradient_buffer_slots_on_cpu():
    """FlatGradAllReduce host logic (no GPU needed): every slice starts on a 128-byte boundary, .grad views alias the flat buffer,
    take() hands a parameter's zeroed slice out once per step (and never for a parameter that already has a gradient)."""
    import torch
    from ofq_b200.ddp import FlatGradAllReduce
    from ofq_b200.quantization import functional as Fn
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(s)) for s in [(7, 5), (3,), (64, 32), (1,), (10, 10)]]
    ddp = FlatGradAllReduce(ps, 1)
    for p, v in zip(ps, ddp.views):
        assert v.storage_offset() % 32 == 0 and v.shape == p.shape and p.grad.data_ptr() == v.data_ptr()
    assert ddp.flat.numel() >= sum(p.numel() for p in ps) and ddp.flat.numel() % 32 == 0
    ddp.flat.fill_(3.0)
    ddp.zero()
    assert Fn.GRAD_SLOTS is ddp and all(p.grad is None for p in ps)
    slot = ddp.take(ps[2])
    assert slot is not None and slot.data_ptr() == ddp.views[2].data_ptr() and float(slot.abs().max()) == 0.0
    assert ddp.take(ps[2]) is None                       # a weight used twice in one forward: the second use gets its own buffer
    assert ddp.take(torch.nn.Parameter(torch.zeros(2))) is None        # not one of ours
    # a backward that wrote into the slot and returned it: autograd-style adoption, then reduce() leaves it in place
    slot.add_(1.5)
    ps[2].grad = slot
    ps[0].grad = torch.full_like(ps[0], 2.0)             # an ordinary gradient tensor: gathered by reduce()
    ddp.reduce()
    assert Fn.GRAD_SLOTS is None
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(ps, ddp.views))
    assert torch.equal(ps[2].grad, torch.full_like(ps[2], 1.5)) and torch.equal(ps[0].grad, torch.full_like(ps[0], 2.0))
    assert float(ps[4].grad.abs().max()) == 0.0          # took no part in the step: zeroed, not stale


def test_kd_losses_and_teacher_on_cpu_take_the_torch_composition():
    """ofq_b200.quantization.utils.KLLossSoft / KDLossSoftandHard keep the reference's call signatures; off the GPU (or for
    probability targets / other reductions) they evaluate the reference's own composition - checked against the goldens of the
    reference classes. Teacher: frozen, no graph, first output of a (logits, info) pair."""
    import torch
    from conftest import load_golden
    from ofq_b200.kd import Teacher
    from ofq_b200.quantization.utils import KDLossSoftandHard, KLLossSoft, Multi_KLLossSoft
    g = load_golden("kd_loss")
    cls, dist = g["b.cls"].clone().requires_grad_(True), g["b.dist"].clone().requires_grad_(True)
    teacher, teacher_dist, y = g["b.teacher"], g["b.teacher_dist"], g["b.y"].long()
    loss = KDLossSoftandHard()((cls, dist), y, (teacher, teacher_dist))
    loss.backward()
    assert torch.equal(loss.detach(), g["b.sh_tuple.loss"]) and torch.equal(cls.grad, g["b.sh_tuple.dcls"])
    assert torch.equal(dist.grad, g["b.sh_tuple.ddist"])
    assert torch.equal(KLLossSoft()((cls, dist), (teacher, teacher_dist), T=2.5).detach(), g["b.soft_T2.5.loss"])
    assert Multi_KLLossSoft is KLLossSoft
    per_sample = KLLossSoft(reduction="none")(cls.detach(), teacher)
    assert per_sample.shape == (cls.shape[0],) and torch.allclose(per_sample.mean(), g["b.soft_T1.0.loss"])

    class Host(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(4, 3)

        def forward(self, x):
            z = self.lin(x)
            return (z, z * 2), None                     # (class, distillation) logits + attention info, as the DeiT host returns

    t = Teacher(Host())
    out = t(torch.randn(5, 4))
    assert isinstance(out, tuple) and len(out) == 2 and not out[0].requires_grad and out[0].grad_fn is None
    assert all(not p.requires_grad for p in t.model.parameters())
