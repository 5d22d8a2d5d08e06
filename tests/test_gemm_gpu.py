"""agnn_gemm (tcgen05 / TMEM / TMA) vs fp64 matmul on the CPU.

TF32X3 is the fp32 parity mode: its error against the fp64 product must be of the order of an
fp32 GEMM's (asserted as <= 4e-6 relative to the output scale, and <= 3x what torch's fp32 CPU
matmul shows on the same operands); BF16 is the stated bf16 mode (2e-2)."""
import pytest
import torch

from analysisgnn_b200 import _lib, linalg
from tests.util import DEV, BF16_REL, rel_err

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 32), (256, 256, 256), (300, 200, 100), (1000, 256, 2560), (77, 640, 512), (4097, 128, 36),
          (130, 52, 64)]


def _ops(m, n, k, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(m, k, generator=g), torch.randn(n, k, generator=g) * 0.1, torch.randn(n, generator=g)


def _check_fp32(got, want64, cpu32):
    err = rel_err(got, want64)
    floor = rel_err(cpu32, want64)
    assert err <= max(4e-6, 3 * floor), (err, floor)


@pytest.mark.parametrize("m,n,k", SHAPES)
def test_linear_tf32x3(m, n, k):
    x, w, b = _ops(m, n, k)
    want = x.double() @ w.double().t() + b.double()
    got = linalg.linear(x.to(DEV), w.to(DEV), b.to(DEV))
    _check_fp32(got, want, x @ w.t() + b)
    got = linalg.linear(x.to(DEV), w.to(DEV), None, relu=True)
    _check_fp32(got, (x.double() @ w.double().t()).relu(), (x @ w.t()).relu())


@pytest.mark.parametrize("m,n,k", SHAPES)
def test_mm_grad_input_layout(m, n, k):
    """dX = dY W: A K-major, B stored [K, N] (MN-major)."""
    if n % 4:
        pytest.skip("row stride of B must be a multiple of 16 bytes (falls back to the library)")
    x, w, _ = _ops(m, n, k)
    wt = w.t().contiguous()                                    # [k, n]
    want = x.double() @ wt.double()
    _check_fp32(linalg.mm(x.to(DEV), wt.to(DEV)), want, x @ wt)
    base = torch.randn(m, n)
    out = base.to(DEV)
    linalg.mm(x.to(DEV), wt.to(DEV), out=out, accumulate=True)
    _check_fp32(out, want + base.double(), x @ wt + base)


@pytest.mark.parametrize("r,m,n", [(5000, 256, 2560), (333, 128, 128), (50000, 64, 256), (1030, 100, 52), (64, 256, 768)])
def test_mm_tn_grad_weight_layout_split_k(r, m, n):
    """dW = dY^T X: both operands MN-major, reduction over rows, deterministic split-K."""
    g = torch.Generator().manual_seed(1)
    a, b = torch.randn(r, m, generator=g), torch.randn(r, n, generator=g)
    want = a.double().t() @ b.double()
    got = linalg.mm_tn(a.to(DEV), b.to(DEV))
    _check_fp32(got, want, a.t() @ b)
    again = linalg.mm_tn(a.to(DEV), b.to(DEV))
    assert torch.equal(got, again)                            # fixed reduction order


def test_presplit_operands_are_reused():
    x, w, b = _ops(640, 256, 512)
    xs = linalg.split(x.to(DEV))
    assert torch.equal((xs.hi.view(torch.int32) & 0x1FFF), torch.zeros_like(xs.hi, dtype=torch.int32))   # TF32-exact
    assert rel_err(xs.hi.double() + xs.lo.double(), x.double()) < 3e-7
    y1 = linalg.linear(xs, w.to(DEV), b.to(DEV))
    y2 = linalg.linear(x.to(DEV), w.to(DEV), b.to(DEV))
    assert torch.equal(y1, y2)
    gw = linalg.mm_tn(linalg.split(torch.randn(640, 128, device=DEV)), xs)
    assert gw.shape == (128, 512)


@pytest.mark.parametrize("m,n,k", [(256, 256, 256), (1000, 256, 2560), (300, 200, 96)])
def test_bf16_mode(m, n, k):
    x, w, b = _ops(m, n, k)
    xb, wb = x.to(DEV, torch.bfloat16), w.to(DEV, torch.bfloat16)
    want = xb.double().cpu() @ wb.double().cpu().t() + b.double()
    got = linalg.linear(xb, wb, b.to(DEV))
    assert got.dtype == torch.bfloat16
    assert rel_err(got.float(), want) <= BF16_REL
    wt = wb.t().contiguous()
    assert rel_err(linalg.mm(xb, wt).float(), xb.double().cpu() @ wt.double().cpu()) <= BF16_REL
    a = torch.randn(2000, m, device=DEV).to(torch.bfloat16)
    c = torch.randn(2000, n, device=DEV).to(torch.bfloat16)
    assert rel_err(linalg.mm_tn(a, c).float(), a.double().cpu().t() @ c.double().cpu()) <= BF16_REL


def test_single_tf32_pass_is_not_the_parity_mode():
    """AGNN_GEMM_TF32 (one product) is ~1e-3; the split mode must be orders of magnitude closer."""
    x, w, _ = _ops(512, 256, 1024)
    xs, ws = linalg.split(x.to(DEV)), linalg.split(w.to(DEV))
    want = x.double() @ w.double().t()
    out = torch.empty(512, 256, device=DEV)
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.agnn_gemm(_lib.GEMM_TF32, 0, 0, 512, 256, 1024, xs.hi.data_ptr(), None, 1024, ws.hi.data_ptr(), None,
                             1024, out.data_ptr(), 256, None, 0, 1, None, 0, st))
    e1 = rel_err(out, want)
    e3 = rel_err(linalg.linear(xs, ws), want)
    assert 1e-5 < e1 < 5e-3 and e3 < 4e-6, (e1, e3)


def test_unsupported_strides_fall_back_to_the_library():
    x, w, b = _ops(100, 185, 153)                              # project_dict's 25+128 input: 612-byte rows
    want = x.double() @ w.double().t() + b.double()
    got = linalg.linear(x.to(DEV), w.to(DEV), b.to(DEV))
    assert rel_err(got, want) < 4e-6


def test_argument_errors():
    lib = _lib.lib()
    assert lib.agnn_gemm(7, 0, 0, 8, 8, 8, None, None, 8, None, None, 8, None, 8, None, 0, 1, None, 0, None) == -1
    assert b"gemm" in lib.agnn_last_error()
