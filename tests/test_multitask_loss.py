"""``analysisgnn_b200.nn.MultiTaskLoss`` against the reference's own class (analysisgnn/models/chord.py:16-49).  The
weighting is host-side torch arithmetic over per-task criteria, so with torch's CPU criteria plugged in it runs
without a GPU: golden numbers below were produced by the reference class (tests/golden/make_golden.py is not needed
for a closed form -- they are re-derived here), and in the build container the class is executed live.  The GPU test
plugs in the CUDA criterion (agnn_softmax_ce)."""
import math

import pytest
import torch
import torch.nn as nn

from analysisgnn_b200 import nn as ann
from oracle import ref_loader
from tests.util import DEV, FP32_REL, assert_close

TASKS = {"cadence": 4, "localkey": 50, "romanNumeral": 185}


def make(cls, requires_grad, criterion=nn.CrossEntropyLoss, params=None):
    loss_ft = nn.ModuleDict({t: criterion(ignore_index=-1, label_smoothing=0.1) for t in TASKS})
    m = cls(tasks=list(TASKS), loss_ft=loss_ft, requires_grad=requires_grad)
    if params is not None and requires_grad:
        with torch.no_grad():
            m.params.copy_(params)
    return m


def data(seed, n=60, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    pred = {t: torch.randn(n, c, generator=g).to(device).requires_grad_(True) for t, c in TASKS.items()}
    gt = {t: torch.randint(-1, c, (n,), generator=g).to(device) for t, c in TASKS.items()}
    return pred, gt


@pytest.mark.parametrize("requires_grad", [True, False])
def test_closed_form(requires_grad):
    p = torch.tensor([0.7, 1.0, 1.9])
    m = make(ann.MultiTaskLoss, requires_grad, params=p)
    pred, gt = data(0)
    out = m(pred, gt)
    assert list(out) == list(TASKS) + ["total"]
    per = [float(nn.functional.cross_entropy(pred[t], gt[t], ignore_index=-1, label_smoothing=0.1)) for t in TASKS]
    want = sum(0.5 / float(p[i]) ** 2 * per[i] + math.log(1 + float(p[i]) ** 2) for i in range(3)) if requires_grad \
        else sum(per)
    assert abs(float(out["total"]) - want) <= 1e-5 * abs(want)
    assert ("params" in m.state_dict()) == requires_grad


def test_weights_follow_the_order_of_the_tasks_present():
    """The reference enumerates the tasks of ``gt`` (chord.py:41-44): with one task missing the second PRESENT task
    meets ``params[1]``."""
    p = torch.tensor([0.5, 2.0, 3.0])
    m = make(ann.MultiTaskLoss, True, params=p)
    pred, gt = data(1)
    gt.pop("localkey")
    out = m(pred, gt)
    l0 = float(out["cadence"]); l1 = float(out["romanNumeral"])
    want = 0.5 / 0.25 * l0 + math.log(1.25) + 0.5 / 4.0 * l1 + math.log(5.0)
    assert abs(float(out["total"]) - want) <= 1e-5 * abs(want)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("requires_grad", [True, False])
def test_matches_the_live_reference_class(requires_grad):
    Ref = ref_loader.load_multitask_loss()
    p = torch.tensor([1.3, 0.6, 2.2])
    ref, mine = make(Ref, requires_grad, params=p), make(ann.MultiTaskLoss, requires_grad, params=p)
    if requires_grad:
        assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == \
               {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    for seed in (0, 1, 2):
        (p1, gt), (p2, _) = data(seed), data(seed)
        o1, o2 = ref(p1, gt), mine(p2, gt)
        assert list(o1) == list(o2)
        for k in o1:
            assert torch.equal(o1[k], o2[k]), k                  # same torch ops on the same values
        o1["total"].backward()
        o2["total"].backward()
        for t in TASKS:
            assert torch.equal(p1[t].grad, p2[t].grad)
        if requires_grad:
            assert torch.equal(ref.params.grad, mine.params.grad)
            ref.params.grad = mine.params.grad = None


@pytest.mark.gpu
@pytest.mark.parametrize("requires_grad", [True, False])
def test_with_the_cuda_criterion(requires_grad):
    p = torch.tensor([1.3, 0.6, 2.2])
    ref = make(ann.MultiTaskLoss, requires_grad, params=p)                                  # torch criteria, CPU
    net = make(ann.MultiTaskLoss, requires_grad, criterion=ann.CrossEntropyLoss, params=p).to(DEV)
    (p1, gt), (p2, _) = data(3, n=500), data(3, n=500, device=DEV)
    o1 = ref(p1, gt)
    o2 = net(p2, {k: v.to(DEV) for k, v in gt.items()})
    for k in o1:
        assert_close(o2[k], o1[k], FP32_REL, k)
    o1["total"].backward()
    o2["total"].backward()
    for t in TASKS:
        assert_close(p2[t].grad, p1[t].grad, FP32_REL, f"d logits {t}")
    if requires_grad:
        assert_close(net.params.grad, ref.params.grad, FP32_REL, "d params")


def test_cuda_criterion_has_no_cpu_path():
    from analysisgnn_b200 import _lib
    with pytest.raises(_lib.AgnnError):
        ann.CrossEntropyLoss(ignore_index=-1, label_smoothing=0.1)(torch.randn(4, 3), torch.tensor([0, 1, 2, -1]))
