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


def test_unaligned_operands_are_repacked_not_sent_to_a_library():
    """Rows that are not 16-byte multiples (project_dict's 25 + 128 = 153 inputs: 612-byte rows) are copied into
    padded buffers and still run on agnn_gemm (counted; refused in strict mode) -- there is no library GEMM."""
    x, w, b = _ops(100, 185, 153)
    want = x.double() @ w.double().t() + b.double()
    before = linalg.stats["repacked_gemms"]
    got = linalg.linear(x.to(DEV), w.to(DEV), b.to(DEV))
    assert rel_err(got, want) < 4e-6 and got.shape == (100, 185)
    assert linalg.stats["repacked_gemms"] == before + 1
    wt = w.t().contiguous()                                    # [153, 185]: MN-major B with 740-byte rows
    _check_fp32(linalg.mm(x.to(DEV), wt.to(DEV)), x.double() @ wt.double(), x @ wt)
    a = torch.randn(100, 37)
    _check_fp32(linalg.mm_tn(a.to(DEV), x.to(DEV)), a.double().t() @ x.double(), a.t() @ x)
    base = torch.randn(100, 185)
    out = base.to(DEV)
    linalg.mm(x.to(DEV), wt.to(DEV), out=out, accumulate=True)
    _check_fp32(out, x.double() @ wt.double() + base.double(), x @ wt + base)
    old = _lib.strict()
    _lib.set_strict(True)
    try:
        with pytest.raises(_lib.AgnnError):
            linalg.linear(x.to(DEV), w.to(DEV), b.to(DEV))
    finally:
        _lib.set_strict(old)


@pytest.mark.parametrize("operands", ["tf32", "f16"])
def test_grouped_launch_matches_single_launches(operands):
    """agnn_gemm_grouped: independent problems of different sizes (node types, task heads), one of them empty, in
    ONE launch -- bit-identical to launching them one by one, in all three layout combinations."""
    g = torch.Generator().manual_seed(3)
    rows = [5000, 1133, 287, 0, 50000 if operands == "f16" else 640]
    mk = (lambda t: linalg.split_f16(t.to(DEV)) if t.shape[0] else t.to(DEV)) if operands == "f16" else \
        (lambda t: t.to(DEV))
    xs = [torch.randn(r, 256, generator=g) for r in rows]
    ws = [torch.randn(n, 256, generator=g) * 0.1 for n in (256, 128, 64, 32, 192)]
    bs = [torch.randn(w.shape[0], generator=g) for w in ws]
    xd = [mk(x) for x in xs]
    wd = [mk(w) if operands == "f16" else w.to(DEV) for w in ws]
    before = linalg.stats.get("gemm_launches", 0)
    outs = linalg.linear_group(xd, wd, [b.to(DEV) for b in bs], relu=True)
    assert linalg.stats["gemm_launches"] == before + 1
    for x, w, b, xo, wo, o in zip(xs, ws, bs, xd, wd, outs):
        assert o.shape == (x.shape[0], w.shape[0])
        if x.shape[0] == 0:
            continue
        _check_fp32(o, (x.double() @ w.double().t() + b.double()).relu(), (x @ w.t() + b).relu())
        assert torch.equal(o, linalg.linear(xo, wo, b.to(DEV), relu=True))
    # grad-weight products (split-K inside the launch) and grad-input products of the same group
    gs = [torch.randn(x.shape[0], w.shape[0], generator=g) for x, w in zip(xs, ws)]
    gd = [mk(t) for t in gs]
    dws = linalg.mm_tn_group(gd, xd)
    dxs = linalg.mm_group(gd, wd)
    for x, w, gg, go, xo, wo, dw, dx in zip(xs, ws, gs, gd, xd, wd, dws, dxs):
        if x.shape[0] == 0:
            continue
        _check_fp32(dw, gg.double().t() @ x.double(), gg.t() @ x)
        _check_fp32(dx, gg.double() @ w.double(), gg @ w)
        # (the group picks its split counts for the group as a whole: the grad-weight sums are ordered differently
        # than in a launch of their own, so only the unsplit products are compared bit for bit)
        assert torch.equal(dx, linalg.mm(go, wo))


@pytest.mark.parametrize("in_kernel", [False, True])
def test_split_k_reduction_is_deterministic(in_kernel, monkeypatch):
    """Split-K partials are added in split order -- by one grouped reduce kernel behind the launch (default) or, with
    ticket counters, inside the launch by the CTA that stores a tile's last partial (AGNN_SPLITK=tickets): bit-identical
    from run to run and to each other, ticket counters back at zero."""
    monkeypatch.setattr(linalg, "SPLITK_IN_KERNEL", in_kernel)
    g = torch.Generator().manual_seed(4)
    a, b = torch.randn(50000, 384, generator=g).to(DEV), torch.randn(50000, 256, generator=g).to(DEV)
    bias = torch.randn(256, generator=g).to(DEV)
    launches = _lib.launches()
    got = linalg.mm_tn(a, b)
    # two operand splits + the gemm launch (+ ONE reduce launch in the two-kernel form)
    assert _lib.launches() - launches <= (3 if in_kernel else 4)
    for _ in range(3):
        assert torch.equal(got, linalg.mm_tn(a, b))
    dev = torch.device(DEV)
    dev = torch.device("cuda", torch.cuda.current_device()) if dev.index is None else dev
    assert int(linalg._tickets(dev).abs().sum()) == 0
    _check_fp32(got, a.double().cpu().t() @ b.double().cpu(), a.cpu().t() @ b.cpu())
    # the legacy entry point (no ticket array: partials reduced by splitk_reduce_kernel) gives the same bits
    lib = _lib.lib()
    sa, sb = linalg.split(a), linalg.split(b)
    m, n, k = 384, 256, 50000
    import ctypes as C
    one = lambda v: (C.c_int64 * 1)(v)
    split = (C.c_int32 * 1)()
    _lib.check(lib.agnn_gemm_group_split_k(_lib.GEMM_TF32X3, 1, one(m), one(n), one(k), split))
    sk = int(split[0])                                         # the split count the grouped launch chose
    assert sk > 1
    ws = torch.empty(lib.agnn_gemm_workspace(_lib.GEMM_TF32X3, m, n, k, sk), dtype=torch.uint8, device=DEV)
    out = torch.empty(m, n, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.agnn_gemm(_lib.GEMM_TF32X3, 1, 1, m, n, k, sa.hi.data_ptr(), sa.lo.data_ptr(), 384, sb.hi.data_ptr(),
                             sb.lo.data_ptr(), 256, out.data_ptr(), 256, None, 0, sk, ws.data_ptr(), ws.numel(), st))
    assert torch.equal(out, got)
    del bias


@pytest.mark.parametrize("m,n,k,relu", [(5000, 256, 256, True), (300, 200, 96, False), (77, 52, 640, False)])
def test_epilogue_amax_out(m, n, k, relu):
    """amax_out: max |C| from the GEMM epilogue (valid rows / columns only), also through split-K."""
    x, w, b = _ops(m, n, k)
    am = torch.zeros(1, device=DEV)
    y = linalg.linear(x.to(DEV), w.to(DEV), b.to(DEV), relu=relu, amax_out=am)
    assert float(am) == float(y.abs().max())
    g = torch.Generator().manual_seed(9)
    a, c = torch.randn(20000, 64, generator=g).to(DEV), torch.randn(20000, n if n % 4 == 0 else 56, generator=g).to(DEV)
    am2 = torch.zeros(1, device=DEV)
    dw = linalg._group(_lib.MN_MAJOR, _lib.MN_MAJOR, [dict(a=a, b=c, m=64, n=c.shape[1], k=20000, amax_out=am2)])[0]
    assert float(am2) == float(dw.abs().max())


def test_argument_errors():
    lib = _lib.lib()
    assert lib.agnn_gemm(7, 0, 0, 8, 8, 8, None, None, 8, None, None, 8, None, 8, None, 0, 1, None, 0, None) == -1
    assert b"gemm" in lib.agnn_last_error()


def test_presplit_of_several_weights_in_one_launch():
    """agnn_split_f16_multi (a cluster of 8 CTAs per matrix, partial maxima through distributed shared memory):
    bit-identical to agnn_amax + agnn_split_f16 on each matrix, exact amax, one launch."""
    g = torch.Generator().manual_seed(6)
    shapes = [(256, 2560), (256, 768), (64, 128), (185, 64), (8, 8), (384, 256), (1000, 40)]
    ws = [(torch.randn(r, c, generator=g) * (10.0 ** (i - 3))).to(DEV) for i, (r, c) in enumerate(shapes)]
    ws[4].zero_()                                              # an all-zero matrix: amax 0, scale 1
    linalg.begin_step()
    before = _lib.launches()
    linalg.presplit_f16(ws)
    assert _lib.launches() - before == 1
    for w in ws:
        got = linalg._cached_split_f16(w)                      # served from the cache
        want = linalg.split_f16(w)
        assert float(got.amax) == float(w.abs().max()) == float(want.amax)
        assert torch.equal(got.hi, want.hi) and torch.equal(got.lo, want.lo)
    assert _lib.launches() - before == 1 + 2 * len(ws)         # only the reference splits launched anything more
