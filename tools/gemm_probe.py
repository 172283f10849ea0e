"""GPU probe for the tcgen05 GEMM engine: exact int8 and bf16 checks vs torch, small timing."""
import ctypes as C, sys, time
import torch
sys.path.insert(0, ".")
from ofq_b200._lib import Operand, Vec, GemmOut, LIB_PATH

lib = C.CDLL(str(LIB_PATH)); lib.ofq_last_error.restype = C.c_char_p
print("device_ok", lib.ofq_device_ok(), torch.cuda.get_device_name(0))
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

def gemm(kind, A, B, out, M, N, K, rs=None, cs=None, rt=None, ct=None, k2=1, nb1=1, nb2=1, splits=1,
         a_str=None, b_str=None, o_str=None, acc=0, rs_period=0):
    a = Operand(A.data_ptr(), *(a_str or (A.stride(-2), 0, 0, 0)))
    b = Operand(B.data_ptr(), *(b_str or (B.stride(-2), 0, 0, 0)))
    o = GemmOut(out.data_ptr(), *(o_str or (out.stride(-2), 0, 0)), acc)
    def v(t, period=0, bs=(0, 0)):
        return C.byref(Vec(t.data_ptr(), period, *bs)) if t is not None else None
    rc = lib.ofq_gemm(kind, C.byref(a), C.byref(b), C.byref(o), M, N, K, k2, nb1, nb2, splits,
                      v(rs, rs_period), v(cs), v(rt), v(ct), st)
    if rc: raise RuntimeError(lib.ofq_last_error().decode())

torch.manual_seed(0)
ok = True
for (M, N, K) in [(128, 128, 128), (256, 384, 384), (25344, 384, 384), (1000, 1536, 384), (777, 200, 1536), (198, 198, 64)]:
    A = torch.randint(-8, 8, (M, K), dtype=torch.int8, device="cuda")
    B = torch.randint(-15, 16, (N, K), dtype=torch.int8, device="cuda")
    rs = torch.rand(198, device="cuda") + 0.5
    cs = torch.rand(N, device="cuda") + 0.5
    ct = torch.randn(N, device="cuda")
    ld = (N + 3) // 4 * 4
    out = torch.full((M, ld), float("nan"), device="cuda")
    gemm(0, A, B, out, M, N, K, rs=rs, cs=cs, ct=ct, rs_period=198)
    torch.cuda.synchronize()
    acc = (A.double() @ B.double().T)
    ref = acc * rs[torch.arange(M, device="cuda") % 198].double()[:, None] * cs.double()[None] + ct.double()[None]
    err = (out[:, :N].double() - ref).abs().max().item() / ref.abs().max().item()
    # exactness of the integer part
    out2 = torch.empty((M, ld), device="cuda")
    gemm(0, A, B, out2, M, N, K)
    torch.cuda.synchronize()
    exact = torch.equal(out2[:, :N].double(), acc)
    print(f"i8 {M}x{N}x{K}: rel err {err:.2e} exact_int {exact}")
    ok &= exact and err < 1e-6

for (M, N, K) in [(128, 128, 64), (384, 1536, 25344), (25344, 384, 1536), (500, 72, 200)]:
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randint(-3, 4, (N, K), device="cuda").bfloat16()
    ld = (N + 3) // 4 * 4
    out = torch.zeros((M, ld), device="cuda")
    splits = 8 if K > 8192 else 1
    rt = torch.randn(M, device="cuda"); ct = torch.randn(N, device="cuda")
    gemm(1, A, B, out, M, N, K, rt=rt, ct=ct, splits=splits, acc=1 if splits > 1 else 0)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T + rt.double()[:, None] * ct.double()[None]
    err = (out[:, :N].double() - ref).norm().item() / ref.norm().item()
    print(f"bf16 {M}x{N}x{K} splits {splits}: rel err {err:.2e}")
    ok &= err < 1e-5

# batched with k2 (attention-like): A[b][n][c] shared over heads, B[b][d][h][c]
Bt, H, Nt, Cc = 3, 6, 198, 384
qx = torch.randint(-2, 2, (Bt, Nt, Cc), dtype=torch.int8, device="cuda")
qk = torch.randint(-2, 2, (Bt, Nt, H, Cc), dtype=torch.int8, device="cuda")
S = torch.empty(Bt, H, Nt, 208, device="cuda")
gemm(0, qx, qk, S, Nt, Nt, Cc, nb1=H, nb2=Bt, a_str=(Cc, 0, 0, Nt * Cc), b_str=(H * Cc, 0, Cc, Nt * H * Cc),
     o_str=(208, Nt * 208, H * Nt * 208))
torch.cuda.synchronize()
ref = torch.einsum("bnc,bdhc->bhnd", qx.double(), qk.double())
print("batched i8 exact:", torch.equal(S[..., :Nt].double(), ref)); ok &= torch.equal(S[..., :Nt].double(), ref)

# timing of the DeiT-S fc1 forward shape
M, N, K = 25344, 1536, 384
A = torch.randint(-2, 2, (M, K), dtype=torch.int8, device="cuda"); B = torch.randint(-3, 4, (N, K), dtype=torch.int8, device="cuda")
out = torch.empty(M, N, device="cuda")
for _ in range(3): gemm(0, A, B, out, M, N, K)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(20): gemm(0, A, B, out, M, N, K)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 20 * 1e-3
print(f"fc1 i8 gemm {t*1e6:.1f} us  {2*M*N*K/t/1e12:.1f} TOP/s  out-write {M*N*4/t/1e9:.0f} GB/s")
print("ALL OK" if ok else "FAILURES")
