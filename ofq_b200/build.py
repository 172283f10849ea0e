"""Build recipe for libofq_b200.so — plain nvcc, sm_100a only, no torch headers.

The library is the drop-in boundary (include/ofq_b200.h): a C-ABI shared object with cudart linked
statically and libcuda resolved at run time, so it loads (and its exports can be checked) on a machine
without a GPU driver.  Objects are compiled per translation unit and cached by source mtime.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
INCLUDE = ROOT.parent / "include"
BUILD = ROOT / "build"
LIB = ROOT / "libofq_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-I", str(INCLUDE), "-I", str(CSRC)] + os.environ.get("OFQ_NVCC_FLAGS", "").split()


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
    return max(p.stat().st_mtime for p in hdrs)


def _compile(src: Path, verbose: bool) -> Path:
    obj = BUILD / (src.stem + ".o")
    newest = max(src.stat().st_mtime, _deps_mtime())
    if obj.exists() and obj.stat().st_mtime >= newest:
        return obj
    cmd = [NVCC, *ARCH, *CFLAGS, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed for {src.name}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    if force:
        for o in BUILD.glob("*.o"):
            o.unlink()
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if LIB.exists() and all(LIB.stat().st_mtime >= o.stat().st_mtime for o in objs):
        return LIB
    cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
