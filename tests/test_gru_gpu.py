"""The GRU drop-in (tensor-core projections + agnn_gru_fwd / _bwd recurrence) vs torch.nn.GRU on the CPU."""
import pytest
import torch
import torch.nn as nn

from analysisgnn_b200.nn.layers import GRU
from tests.util import DEV, FP32_REL, assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hidden,cin,layers,bidir,b,t", [
    (128, 256, 2, True, 5, 40), (64, 32, 1, True, 3, 17), (32, 48, 1, False, 4, 9), (128, 128, 1, True, 1, 1),
    (128, 256, 2, True, 100, 64), (64, 64, 2, False, 7, 130),
    (128, 256, 2, True, 130, 128),      # 16 640 rows: the projections take the fp16 operand form (linalg.prepare_auto)
    # hidden sizes of the per-time-step kernels (agnn_gru_supported == AGNN_GRU_STEPWISE): MetricalConvLayer's 512,
    # ragged sequence blocks (66 = 64 + 2 sequences forward, 32 + 32 + 2 backward), a single step, one direction
    (256, 96, 1, True, 5, 23), (512, 512, 1, True, 66, 31), (192, 64, 2, False, 33, 12), (256, 256, 1, True, 3, 1),
    (512, 512, 1, True, 64, 124)])
def test_matches_torch_gru(hidden, cin, layers, bidir, b, t):
    torch.manual_seed(hidden + t)
    ref = nn.GRU(cin, hidden, num_layers=layers, batch_first=True, bidirectional=bidir)
    mine = GRU(cin, hidden, num_layers=layers, batch_first=True, bidirectional=bidir)
    mine.load_state_dict(ref.state_dict())
    assert list(mine.state_dict()) == list(ref.state_dict())
    mine.to(DEV)
    x = torch.randn(b, t, cin)
    w = torch.rand(b, t, hidden * (2 if bidir else 1)) + 0.25
    x1 = x.clone().requires_grad_(True)
    o1, h1 = ref(x1)
    (o1 * w).sum().backward()
    x2 = x.to(DEV).requires_grad_(True)
    o2, h2 = mine(x2)
    (o2 * w.to(DEV)).sum().backward()
    tol = 3 * FP32_REL                       # hundreds of dependent steps
    assert_close(o2, o1, tol, "output")
    assert_close(h2, h1, tol, "h_n")
    assert_close(x2.grad, x1.grad, tol, "dx")
    for (n, p), q in zip(mine.named_parameters(), ref.parameters()):
        assert_close(p.grad, q.grad, tol, n)


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("b,t,bidir", [(5, 40, True), (19, 33, True), (100, 64, True), (9, 21, False), (1, 1, True)])
def test_both_resident_implementations_of_hidden_128(mode, b, t, bidir):
    """agnn_gru_mode: SIMT / tensor-core kernels per pass (8 sequences per CTA: 5 = one ragged tile, 19 = 2 + 3/8)."""
    from analysisgnn_b200 import _lib
    old = _lib.lib().agnn_gru_mode(mode)
    try:
        test_matches_torch_gru(128, 96, 2, bidir, b, t)
    finally:
        _lib.lib().agnn_gru_mode(old)


def test_unsupported_sizes_use_the_library_rnn():
    torch.manual_seed(0)
    ref = nn.GRU(24, 20, batch_first=True, bidirectional=True)
    mine = GRU(24, 20, batch_first=True, bidirectional=True)
    mine.load_state_dict(ref.state_dict())
    mine.to(DEV)
    x = torch.randn(3, 11, 24)
    assert_close(mine(x.to(DEV))[0], ref(x)[0], 3 * FP32_REL)


def test_inter_layer_dropout_only_in_training():
    torch.manual_seed(0)
    g = GRU(32, 32, num_layers=2, batch_first=True, bidirectional=True, dropout=0.5).to(DEV)
    x = torch.randn(4, 12, 32, device=DEV)
    g.eval()
    a, b = g(x)[0], g(x)[0]
    assert torch.equal(a, b)
    g.train()
    c = g(x)[0]
    assert not torch.equal(a, c)
