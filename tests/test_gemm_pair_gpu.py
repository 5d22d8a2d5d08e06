"""agnn_gemm_pair (csrc/gemm2.cu): the large fp32-parity products on CTA pairs (tcgen05.mma.cta_group::2, 256 x 256
tiles) against the fp64 product and against the single-CTA kernel on the same operand pairs."""
import pytest
import torch

from analysisgnn_b200 import _lib, linalg
from tests.util import DEV, rel_err

pytestmark = pytest.mark.gpu


def _run_pair(xs, ws, b_layout, m, n, k, bias=None, relu=False, amax_out=None):
    out = torch.empty((m, n), dtype=torch.float32, device=DEV)
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    ptr = lambda t: t.data_ptr() if t is not None else None
    _lib.check(lib.agnn_gemm_pair(b_layout, m, n, k, xs.hi.data_ptr(), xs.lo.data_ptr(), xs.hi.stride(0),
                                  xs.amax.data_ptr(), ws.hi.data_ptr(), ws.lo.data_ptr(), ws.hi.stride(0),
                                  ws.amax.data_ptr(), out.data_ptr(), out.stride(0), ptr(bias),
                                  _lib.GEMM_RELU if relu else 0, ptr(amax_out), st), "agnn_gemm_pair")
    return out


@pytest.mark.parametrize("m,n,k", [(4096, 256, 64), (50000, 256, 2560), (8192, 512, 256), (4100, 300, 200),
                                   (20000, 2560, 256)])
def test_pair_forward_layout(m, n, k):
    """Y = X W^T + b (A K-major, B K-major), ragged M / N / K edges included."""
    g = torch.Generator().manual_seed(m + n)
    x, w, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g) * 0.1, torch.randn(n, generator=g)
    xs, ws = linalg.split_f16(x.to(DEV)), linalg.split_f16(w.to(DEV))
    am = torch.zeros(1, device=DEV)
    got = _run_pair(xs, ws, _lib.K_MAJOR, m, n, k, b.to(DEV), relu=True, amax_out=am)
    want = (x.double() @ w.double().t() + b.double()).relu()
    assert rel_err(got, want) <= 4e-6
    assert float(am) == float(got.abs().max())
    single = linalg.linear(xs, ws, b.to(DEV), relu=True)          # the single-CTA kernel on the same pairs
    assert rel_err(got, single.double().cpu()) <= 1e-6


@pytest.mark.parametrize("m,n,k", [(50000, 2560, 256), (8192, 256, 384), (4100, 328, 200)])
def test_pair_grad_input_layout(m, n, k):
    """dX = dY W (A K-major, B stored [K, N]: MN-major)."""
    g = torch.Generator().manual_seed(m + k)
    a, wt = torch.randn(m, k, generator=g), torch.randn(k, n, generator=g) * 0.1
    as_, ws = linalg.split_f16(a.to(DEV)), linalg.split_f16(wt.to(DEV))
    got = _run_pair(as_, ws, _lib.MN_MAJOR, m, n, k)
    assert rel_err(got, a.double() @ wt.double()) <= 4e-6


def test_pair_route_inside_linalg(monkeypatch):
    """With the route switched on, linalg sends the big single products to the pair kernel and everything else to the
    grouped kernel; same numbers either way."""
    g = torch.Generator().manual_seed(3)
    x, w = torch.randn(30000, 512, generator=g).to(DEV), (torch.randn(256, 512, generator=g) * 0.1).to(DEV)
    xs = linalg.split_f16(x)
    monkeypatch.setattr(linalg, "GEMM_PAIR", False)
    ref = linalg.linear(xs, w)
    monkeypatch.setattr(linalg, "GEMM_PAIR", True)
    before = linalg.stats.get("gemm_pair_launches", 0)
    got = linalg.linear(xs, w)
    assert linalg.stats.get("gemm_pair_launches", 0) == before + 1
    assert rel_err(got, ref.double().cpu()) <= 1e-6
    assert not lib_supported_small()


def lib_supported_small():
    return bool(_lib.lib().agnn_gemm_pair_supported(_lib.GEMM_F16X3, _lib.K_MAJOR, 300, 256, 256, 0))
