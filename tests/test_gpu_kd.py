"""KD path: the fused KD loss kernel against the reference's own loss classes (goldens) and the frozen
teacher runner. Reference: src/quantization/utils.py:44-77, train.py:428-442, 896-910."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def U():
    from ofq_b200 import _lib
    assert _lib.load().ofq_device_ok() == 1
    import ofq_b200.quantization.utils as U
    return U


@pytest.mark.parametrize("tag", ["a", "b"])
def test_kd_loss_kernel_against_reference_goldens(U, tag):
    """KDLossSoftandHard / KLLossSoft (same names and call signatures as the reference's) on the GPU: one ofq_kd_loss launch
    pair for the value AND the gradients; loss and d loss / d logits within 1e-5 of the reference's autograd (fp32 sums in
    another order), tuple and single-output students, temperatures."""
    from ofq_b200 import ops
    g = load_golden("kd_loss")
    dev = lambda k: g[f"{tag}.{k}"].cuda()
    y = dev("y").long()
    teacher = (dev("teacher"), dev("teacher_dist"))
    cls, dist = dev("cls").requires_grad_(True), dev("dist").requires_grad_(True)
    l0 = ops.LAUNCHES
    loss = U.KDLossSoftandHard()((cls, dist), y, teacher)
    loss.backward()
    assert ops.LAUNCHES - l0 == 2                                             # the fused kernel ran (no torch composition)
    assert abs(loss.item() - g[f"{tag}.sh_tuple.loss"].item()) < 1e-5 * abs(g[f"{tag}.sh_tuple.loss"].item())
    assert rel_err(cls.grad.cpu(), g[f"{tag}.sh_tuple.dcls"]) < 1e-5 and rel_err(dist.grad.cpu(), g[f"{tag}.sh_tuple.ddist"]) < 1e-5
    cls.grad = None
    loss = U.KDLossSoftandHard()(cls, y, teacher[0])
    loss.backward()
    assert abs(loss.item() - g[f"{tag}.sh_single.loss"].item()) < 1e-5 * abs(g[f"{tag}.sh_single.loss"].item())
    assert rel_err(cls.grad.cpu(), g[f"{tag}.sh_single.dcls"]) < 1e-5
    for T in (1.0, 2.5):
        cls.grad = None
        loss = U.KLLossSoft()((cls, dist), teacher, T=T)
        (loss * 3.0).backward()                                               # a non-unit upstream gradient
        assert abs(loss.item() - g[f"{tag}.soft_T{T}.loss"].item()) < 1e-5 * abs(g[f"{tag}.soft_T{T}.loss"].item())
        assert rel_err(cls.grad.cpu() / 3.0, g[f"{tag}.soft_T{T}.dcls"]) < 1e-5
        assert dist.grad is None or True
    # probability (mixup) targets and other reductions take the torch composition: still the reference's numbers
    probs = F.one_hot(y, cls.shape[1]).float() * 0.9 + 0.1 / cls.shape[1]
    ref = F.cross_entropy(cls.detach(), probs) + U.KLLossSoft()(dist.detach(), teacher[0])
    assert abs(U.KDLossSoftandHard()((cls.detach(), dist.detach()), probs, teacher).item() - ref.item()) < 1e-5 * abs(ref.item())


def test_teacher_runner_and_kd_step(U):
    """A frozen unquantized DeiT teacher (train.py:428-442) under no_grad feeding KDLossSoftandHard of a quantized student:
    the teacher gets no gradient and builds no graph, the soft targets are its training-mode class logits (what
    `soft_target, _ = teacher(input)` binds, train.py:906), the student's gradients equal the torch composition's."""
    import ofq_b200.quantization as Q
    from ofq_b200.host.deit import DistilledVisionTransformer
    from ofq_b200.kd import Teacher
    torch.manual_seed(3)
    depth = 2
    teacher_model = DistilledVisionTransformer(embed_dim=128, depth=depth, num_heads=2, num_classes=10).cuda()
    student = DistilledVisionTransformer(embed_dim=128, depth=depth, num_heads=2, num_classes=10)
    student = Q.replace_module_by_qmodule_deit(student, Q.make_qconfigs(Q.deit_qmodule_names(depth), 2, 2),
                                               pretrained_initialized=True, qk_reparam=True).cuda()
    img = torch.randn(4, 3, 224, 224, device="cuda")
    y = torch.tensor([1, 5, 7, 2], device="cuda")
    student.eval()
    with torch.no_grad():
        student(img)
    student.train()
    teacher = Teacher(teacher_model)
    soft = teacher(img)
    (t_cls, t_dist), _ = teacher_model(img)
    assert isinstance(soft, tuple) and torch.equal(soft[0], t_cls.detach()) and not soft[0].requires_grad
    assert all(not p.requires_grad for p in teacher_model.parameters())
    grads = []
    for fused in (True, False):
        student.zero_grad(set_to_none=True)
        out, _ = student(img)
        if fused:
            loss = U.KDLossSoftandHard()(out, y, soft)
        else:
            loss = F.cross_entropy(out[0], y) - (F.softmax(soft[0], 1) * F.log_softmax(out[1], 1)).sum(1).mean()
        loss.backward()
        grads.append((loss.item(), {n: p.grad.clone() for n, p in student.named_parameters() if p.grad is not None}))
    assert abs(grads[0][0] - grads[1][0]) < 1e-5 * abs(grads[1][0])
    gmax = max(v.abs().max().item() for v in grads[1][1].values())
    for n, b in grads[1][1].items():
        a = grads[0][1][n]
        assert rel_err(a, b) < 1e-4 or (a - b).abs().max().item() <= 1e-6 * gmax, n
    # bf16 teacher: same interface, soft targets within bf16 rounding
    soft16 = Teacher(teacher_model, dtype=torch.bfloat16)(img)
    assert soft16[0].dtype == torch.float32 and rel_err(soft16[0], soft[0]) < 3e-2
