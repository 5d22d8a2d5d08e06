"""Row-normalisation kernels and the Linear / LayerNorm drop-ins vs torch on the CPU (fp32, 1e-5)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from analysisgnn_b200 import ops
from analysisgnn_b200.nn.layers import LayerNorm, Linear
from tests.util import DEV, FP32_REL, assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,cols", [(1, 64), (1000, 128), (50001, 256), (333, 512), (77, 1024), (100, 36)])
def test_layer_norm_forward_backward(rows, cols):
    torch.manual_seed(cols)
    x = torch.randn(rows, cols) * 3 + 1
    g, b = torch.randn(cols), torch.randn(cols)
    w = torch.randn(rows, cols)
    x1, g1, b1 = (t.clone().requires_grad_(True) for t in (x, g, b))
    F.layer_norm(x1, (cols,), g1, b1, 1e-5).mul(w).sum().backward()
    x2, g2, b2 = (t.to(DEV).requires_grad_(True) for t in (x, g, b))
    y = ops.layer_norm(x2, g2, b2, 1e-5)
    y.mul(w.to(DEV)).sum().backward()
    assert_close(y, F.layer_norm(x, (cols,), g, b, 1e-5), FP32_REL, "forward")
    assert_close(x2.grad, x1.grad, FP32_REL, "dx")
    assert_close(g2.grad, g1.grad, 2 * FP32_REL, "dgamma")
    assert_close(b2.grad, b1.grad, 2 * FP32_REL, "dbeta")


@pytest.mark.parametrize("relu_first", [False, True])
@pytest.mark.parametrize("rows,cols", [(500, 256), (64, 512), (1001, 16)])
def test_l2norm_relu(relu_first, rows, cols):
    torch.manual_seed(1)
    x = torch.randn(rows, cols)
    x[3] = 0.0                                              # an all-zero row: clamp at eps, zero gradient
    w = torch.randn(rows, cols)
    x1 = x.clone().requires_grad_(True)
    ref = F.normalize(F.relu(x1), p=2, dim=-1) if relu_first else F.relu(F.normalize(x1, p=2, dim=-1))
    ref.mul(w).sum().backward()
    x2 = x.to(DEV).requires_grad_(True)
    y = ops.l2norm_relu(x2, relu_first)
    y.mul(w.to(DEV)).sum().backward()
    assert_close(y, ref, FP32_REL, "forward")
    assert_close(x2.grad, x1.grad, FP32_REL, "dx")


def test_colsum():
    x = torch.randn(12345, 256).abs()
    assert_close(ops.colsum(x.to(DEV)), x.double().sum(0), FP32_REL)
    y = torch.randn(100, 185)                               # not a multiple of 4: library route
    assert_close(ops.colsum(y.to(DEV)), y.sum(0), FP32_REL)


@pytest.mark.parametrize("cin,cout,lead", [(256, 256, (1000,)), (153, 256, (777,)), (128, 185, (300,)), (64, 4, (50,)),
                                           (256, 128, (7, 90))])
def test_linear_module_matches_nn_linear(cin, cout, lead):
    torch.manual_seed(2)
    ref = nn.Linear(cin, cout)
    mine = Linear(cin, cout)
    mine.load_state_dict(ref.state_dict())
    mine.to(DEV)
    x = torch.randn(*lead, cin)
    w = torch.rand(*lead, cout) + 0.25
    x1 = x.clone().requires_grad_(True)
    ref(x1).mul(w).sum().backward()
    x2 = x.to(DEV).requires_grad_(True)
    y = mine(x2)
    y.mul(w.to(DEV)).sum().backward()
    assert_close(y, ref(x), FP32_REL, "forward")
    assert_close(x2.grad, x1.grad, FP32_REL, "dx")
    assert_close(mine.weight.grad, ref.weight.grad, FP32_REL, "dW")
    assert_close(mine.bias.grad, ref.bias.grad, FP32_REL, "db")
    assert list(mine.state_dict()) == list(ref.state_dict())


def test_layernorm_module_state_dict_and_fallbacks():
    ref, mine = nn.LayerNorm(64), LayerNorm(64)
    assert list(mine.state_dict()) == list(ref.state_dict())
    mine.to(DEV)
    x = torch.randn(10, 5, 64)
    assert_close(mine(x.to(DEV)), ref(x), FP32_REL)
    odd = LayerNorm(30).to(DEV)                             # width not a multiple of 4: library kernel
    assert_close(odd(torch.ones(4, 30, device=DEV)), torch.zeros(4, 30), 1.0)


@pytest.mark.parametrize("rows,cols", [(1, 64), (50001, 256), (333, 1024), (77, 36)])
@pytest.mark.parametrize("with_relu", [False, True])
def test_grad_prepare_matches_where_colsum_split(rows, cols, with_relu):
    """agnn_grad_prepare = relu mask + TF32 hi/lo pair + column sums of one gradient matrix in one pass."""
    from analysisgnn_b200 import linalg
    torch.manual_seed(rows + cols)
    g = torch.randn(rows, cols, device=DEV)
    o = torch.relu(torch.randn(rows, cols, device=DEV)) if with_relu else None
    op, cs = ops.prepare_grad(g, o, want_colsum=True)
    ref = torch.where(o > 0, g, torch.zeros((), device=DEV)) if with_relu else g
    want = linalg.split(ref.contiguous())
    assert isinstance(op, linalg.Split)
    assert torch.equal(op.hi, want.hi) and torch.equal(op.lo, want.lo)
    assert_close(cs, ref.double().sum(0).float(), FP32_REL, "column sums")
    op2, none = ops.prepare_grad(g, o, want_colsum=False)
    assert none is None and torch.equal(op2.hi, want.hi)
