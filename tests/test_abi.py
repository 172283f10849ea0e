"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads without a GPU driver and exports
every symbol include/ofq_b200.h declares; compute entry points refuse to run without a B200."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from ofq_b200 import _lib, build
    build.build()
    return _lib.load()


def header_symbols():
    text = (ROOT / "include" / "ofq_b200.h").read_text()
    return sorted(set(re.findall(r"OFQ_API\s+[\w\s\*]+?\b(ofq_\w+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    from ofq_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 20
    assert sorted(_lib.EXPORTS) == syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ofq_b200.h but not exported"


def test_no_torch_types_in_abi():
    text = (ROOT / "include" / "ofq_b200.h").read_text()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)          # declarations only, comments stripped
    assert "at::" not in code and "torch" not in code.lower() and "Tensor" not in code
    assert 'extern "C"' in code


def test_library_has_no_libcuda_or_torch_dependency():
    import subprocess
    out = subprocess.run(["ldd", str(ROOT / "ofq_b200" / "libofq_b200.so")], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out
    assert "libcuda.so" not in out          # resolved at run time through cudaGetDriverEntryPoint


def test_sass_is_blackwell_native():
    """The GEMM engine must carry tcgen05 MMAs (UTC*MMA), TMEM loads (LDTM) and TMA (UTMALDG/UTMASTG)."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", str(ROOT / "ofq_b200" / "libofq_b200.so")], capture_output=True, text=True).stdout
    for mnem in ("UTCIMMA", "UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG"):
        assert mnem in sass, mnem
    assert "sm_100a" in sass
    # per kernel: the round-2 kernels are tcgen05 / TMEM / TMA kernels themselves, not wrappers around the GEMM engine
    per_fn, cur = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            per_fn[cur] = []
        elif cur is not None:
            per_fn[cur].append(line)
    want = {"qkr_attn_fwd16_kernel": ("UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG"),          # S and P.V MMAs, TMEM both ways
            "qkr_attn_bwd_kernel": ("UTCIMMA", "UTCHMMA", "LDTM", "UTMALDG"),                    # int8 S recompute + fp16 dP
            "gemm_lsq_kernel": ("UTCIMMA", "LDTM", "UTMALDG", "UTMASTG"),
            "gemm_dxlsq_kernel": ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG")}
    for frag, mnems in want.items():
        bodies = ["\n".join(v) for k, v in per_fn.items() if frag in k]
        assert bodies, frag
        for m in mnems:
            assert any(m in b for b in bodies), (frag, m)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a machine without a GPU")
def test_compute_entry_points_fail_loudly_without_gpu(lib):
    assert lib.ofq_device_ok() < 0
    buf = (ctypes.c_float * 16)()
    rc = lib.ofq_lsq_effective_scale(ctypes.addressof(buf), 16, 0.1, ctypes.addressof(buf), None, None)
    assert rc != 0 and len(lib.ofq_last_error()) > 0


def test_ops_refuse_cpu_tensors():
    from ofq_b200 import _lib, ops
    with pytest.raises(_lib.OfqError):
        ops.statsq_codes(torch.randn(4, 8), 2)
