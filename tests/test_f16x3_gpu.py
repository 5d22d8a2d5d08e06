"""The fp16 hi/lo operand form of the fp32 parity mode (AGNN_GEMM_F16X3, include/agnn.h): same nn.Linear sites and
the same tolerance as the 3xTF32 mode -- fp16 carries TF32's 11-bit significand, the per-tensor scale is a power of
two and is undone exactly -- at twice the tensor-core rate.

* the GEMM against an fp64 product, every operand layout, split-K, bias / ReLU / accumulate, operands whose
  magnitudes sit far from 1 (the scale must absorb them) and a tensor with a wide dynamic range;
* the producers of operand pairs (agnn_split_f16, agnn_gather_reduce_f16, agnn_grad_prepare_f16) against the fp32
  values they encode;
* the fused message-passing layers with ``linalg.set_parity_operands("f16")``: the oracle comparisons of
  tests/test_hetero_gpu.py, unchanged tolerances."""
import pytest
import torch

from analysisgnn_b200 import linalg
from tests import test_hetero_gpu as th
from tests.util import DEV, rel_err

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 64), (256, 256, 256), (300, 200, 104), (1000, 256, 2560), (77, 640, 512), (4097, 128, 40),
          (130, 56, 64)]


@pytest.fixture(params=["f16", "tf32"])
def f16_mode(request):
    """Both operand forms of the parity mode, whichever is the library default."""
    old_p = linalg.parity_operands()
    linalg.set_parity_operands(request.param)
    yield request.param
    linalg.set_parity_operands(old_p)


def _ops(m, n, k, seed=0, sx=1.0, sw=0.1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(m, k, generator=g) * sx, torch.randn(n, k, generator=g) * sw, torch.randn(n, generator=g) * sx * sw


def _check_fp32(got, want64, cpu32):
    err = rel_err(got, want64)
    floor = rel_err(cpu32, want64)
    assert err <= max(4e-6, 3 * floor), (err, floor)


def test_split_f16_encodes_fp32_to_22_bits():
    g = torch.Generator().manual_seed(3)
    for scale in (1.0, 3e-6, 7e4):
        x = (torch.randn(513, 264, generator=g) * scale).to(DEV)
        s = linalg.split_f16(x)
        assert s.hi.dtype == torch.float16 and s.lo.dtype == torch.float16
        sc = float(linalg.f16_scale(s.amax))
        top = float(x.abs().max()) * sc
        assert 2.0 ** 13 <= top < 2.0 ** 14, top
        back = (s.hi.double() + s.lo.double()) / sc
        # elements within 2^-17 of the amax keep 22 bits; the rest are exact to 2^-25 of the scaled range
        err = (back - x.double()).abs()
        bound = torch.maximum(x.double().abs() * 2.0 ** -21, torch.full_like(err, 2.0 ** -24 / sc))
        assert bool((err <= bound).all())
        assert rel_err(linalg.plain(s), back) < 1e-7


@pytest.mark.parametrize("m,n,k", SHAPES)
@pytest.mark.parametrize("sx,sw", [(1.0, 0.1), (2e-5, 30.0)])
def test_linear_f16x3(m, n, k, sx, sw):
    x, w, b = _ops(m, n, k, sx=sx, sw=sw)
    want = x.double() @ w.double().t() + b.double()
    xs = linalg.split_f16(x.to(DEV))
    got = linalg.linear(xs, w.to(DEV), b.to(DEV))
    _check_fp32(got, want, x @ w.t() + b)
    got = linalg.linear(xs, w.to(DEV), None, relu=True)
    _check_fp32(got, (x.double() @ w.double().t()).relu(), (x @ w.t()).relu())


@pytest.mark.parametrize("m,n,k", SHAPES)
def test_mm_grad_input_layout_f16x3(m, n, k):
    x, w, _ = _ops(m, n, k, seed=1, sx=1e-3)
    wt = w.t().contiguous()                                    # [k, n]
    want = x.double() @ wt.double()
    xs = linalg.split_f16(x.to(DEV))
    _check_fp32(linalg.mm(xs, wt.to(DEV)), want, x @ wt)
    base = torch.randn(m, n) * 1e-4
    out = base.to(DEV)
    linalg.mm(xs, wt.to(DEV), out=out, accumulate=True)
    _check_fp32(out, want + base.double(), x @ wt + base)


@pytest.mark.parametrize("r,m,n", [(5000, 256, 2560), (333, 128, 128), (50000, 64, 256), (1030, 104, 56), (64, 256, 768)])
def test_mm_tn_grad_weight_layout_split_k_f16x3(r, m, n):
    g = torch.Generator().manual_seed(1)
    a, b = torch.randn(r, m, generator=g) * 1e-4, torch.randn(r, n, generator=g)
    want = a.double().t() @ b.double()
    sa, sb = linalg.split_f16(a.to(DEV)), linalg.split_f16(b.to(DEV))
    got = linalg.mm_tn(sa, sb)
    _check_fp32(got, want, a.t() @ b)
    assert torch.equal(got, linalg.mm_tn(sa, sb))             # fixed reduction order


def test_wide_dynamic_range_rows():
    """Rows 1e-4 of the tensor's amax still come out fp32-grade relative to THEIR OWN scale (22 bits hold down to
    2^-17 of the amax)."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(512, 256, generator=g)
    x[::2] *= 1e-4
    w = torch.randn(128, 256, generator=g) * 0.1
    got = linalg.linear(linalg.split_f16(x.to(DEV)), w.to(DEV))
    want = x.double() @ w.double().t()
    assert rel_err(got[::2], want[::2]) <= 4e-6
    assert rel_err(got[1::2], want[1::2]) <= 4e-6


def test_sage_layers_in_f16_mode(f16_mode):
    th.test_sage_layer("sum")
    th.test_sage_layer("mean")
    th.test_sage_layer_with_missing_relations_and_types()


def test_sage_stack_3x256_in_f16_mode(f16_mode):
    th.test_sage_stack_3x256()


def test_sage_stack_trim_to_layer_in_f16_mode(f16_mode):
    th.test_sage_stack_trim_to_layer()


def test_analysis_encoder_shell_in_f16_mode(f16_mode):
    th.test_analysis_encoder_shell("hybridgnn")


@pytest.mark.parametrize("rows,k,n", [(20000, 153, 256), (16384, 128, 185), (30000, 512, 128)])
def test_linear_module_fwd_bwd(f16_mode, rows, k, n):
    """nn.Linear drop-in (ops._Linear): odd widths are padded to the 16-byte row rule of the operand type; from
    16 384 rows on the f16 mode runs all three GEMMs on fp16 pairs."""
    from analysisgnn_b200 import ops
    g = torch.Generator().manual_seed(7)
    x, w, b = torch.randn(rows, k, generator=g), torch.randn(n, k, generator=g) * 0.05, torch.randn(n, generator=g)
    gy = torch.randn(rows, n, generator=g) * 1e-3
    xd, wd, bd = (t.double().requires_grad_(True) for t in (x, w, b))
    (xd @ wd.t() + bd).backward(gy.double())
    xg, wg, bg = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    y = ops.linear(xg, wg, bg)
    fn = y.grad_fn
    while fn is not None and not hasattr(fn, "saved_tensors"):      # through the reshape node
        fn = fn.next_functions[0][0] if fn.next_functions else None
    assert fn is not None
    assert any(t is not None and t.dtype == torch.float16 for t in fn.saved_tensors) == (f16_mode == "f16")
    y.backward(gy.to(DEV))
    assert rel_err(y, x.double() @ w.double().t() + b.double()) <= 4e-6
    for got, want, what in ((xg.grad, xd.grad, "dx"), (wg.grad, wd.grad, "dw"), (bg.grad, bd.grad, "db")):
        assert rel_err(got, want) <= 4e-6, what


def test_intree_layers_in_both_operand_forms(f16_mode):
    """In-tree HeteroConv{SageConvScatter} / MetricalGNN (ops._IntreeSageLayer): reference goldens and the oracle."""
    from tests import test_intree_gpu as ti
    ti.test_golden_heteroconv(0)
    ti.test_golden_sage_and_empty_branch(1)
    ti.test_golden_metricalgnn(2, "eval")
    ti.test_heteroconv_vs_oracle("mean", 128, 512)
    ti.test_heteroconv_vs_oracle("sum", 256, 256)
    ti.test_metricalgnn_config4_shape_small_batch()


def test_f16_mode_uses_the_f16_kernels(f16_mode):
    """The mode is not a silent no-op: the layer's saved operands are fp16 pairs."""
    if f16_mode != "f16":
        pytest.skip("checks the f16 form")
    from analysisgnn_b200 import nn as ann, synth
    b = synth.hetero_batch(2, 60, 3)
    net = ann.HeteroSAGELayer(b["metadata"][1], 32, 32).to(DEV)
    x = {k: v.to(DEV).requires_grad_(True) for k, v in th._features(b, 32).items()}
    out = net(x, {k: v.to(DEV) for k, v in b["edge_index_dict"].items()})
    fn = next(iter(out.values())).grad_fn
    while fn is not None and not hasattr(fn, "x_amax"):
        fn = fn.next_functions[0][0] if fn.next_functions else None
    assert fn is not None and fn.x_amax is not None
    assert any(t is not None and t.dtype == torch.float16 for t in fn.saved_tensors)
