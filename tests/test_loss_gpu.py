"""Objective and lookup-table gradients (agnn_softmax_ce_*, agnn_embedding_bwd) vs torch on the CPU.

The reference computes both with ATen: ``CrossEntropyLoss(ignore_index=-1, label_smoothing=0.1)`` per task
(analysisgnn/models/analysis.py:881-908) and ``nn.Embedding`` for pitch spelling / key signature (:427-428)."""
import pytest
import torch
import torch.nn.functional as F

from analysisgnn_b200 import ops
from tests.util import DEV, FP32_REL, assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,cols", [(1, 4), (257, 4), (5000, 50), (50001, 185), (300, 1000)])
@pytest.mark.parametrize("smoothing", [0.0, 0.1])
def test_cross_entropy_matches_torch(rows, cols, smoothing):
    torch.manual_seed(rows + cols)
    x = torch.randn(rows, cols) * 4
    y = torch.randint(0, cols, (rows,))
    if rows > 3:
        y[torch.rand(rows) < 0.3] = -1                       # ignored rows
        y[0] = 0
    x1 = x.clone().requires_grad_(True)
    ref = F.cross_entropy(x1, y, ignore_index=-1, label_smoothing=smoothing)
    (ref * 1.7).backward()
    x2 = x.to(DEV).requires_grad_(True)
    out = ops.cross_entropy(x2, y.to(DEV), ignore_index=-1, label_smoothing=smoothing)
    (out * 1.7).backward()
    assert_close(out, ref, FP32_REL, "loss")
    assert_close(x2.grad, x1.grad, FP32_REL, "dlogits")
    if rows > 3:
        assert float(x2.grad[y.to(DEV) == -1].abs().sum()) == 0.0


def test_cross_entropy_on_a_padded_row_view_and_all_ignored():
    torch.manual_seed(0)
    buf = torch.randn(100, 188, device=DEV)
    x = buf[:, :185]                                          # what the padded GEMM output hands over
    y = torch.randint(0, 185, (100,), device=DEV)
    ref = F.cross_entropy(x.cpu().contiguous(), y.cpu(), ignore_index=-1, label_smoothing=0.1)
    assert_close(ops.cross_entropy(x, y, ignore_index=-1, label_smoothing=0.1), ref, FP32_REL)
    none = ops.cross_entropy(x, torch.full_like(y, -1), ignore_index=-1)
    assert torch.isnan(none)                                  # 0 / 0 rows, as in ATen


@pytest.mark.parametrize("rows,n_emb,dim", [(1, 35, 64), (50000, 35, 64), (4097, 15, 64), (1000, 7, 20)])
def test_small_embedding_forward_backward(rows, n_emb, dim):
    torch.manual_seed(rows)
    w = torch.randn(n_emb, dim)
    idx = torch.randint(0, n_emb, (rows,))
    up = torch.randn(rows, dim)
    w1 = w.clone().requires_grad_(True)
    F.embedding(idx, w1).mul(up).sum().backward()
    w2 = w.to(DEV).requires_grad_(True)
    out = ops.embedding(idx.to(DEV), w2)
    out.mul(up.to(DEV)).sum().backward()
    assert torch.equal(out.cpu(), F.embedding(idx, w))
    assert_close(w2.grad, w1.grad, FP32_REL, "dweight")
    again = w.to(DEV).requires_grad_(True)
    ops.embedding(idx.to(DEV), again).mul(up.to(DEV)).sum().backward()
    assert torch.equal(again.grad, w2.grad)                   # fixed summation order
