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
