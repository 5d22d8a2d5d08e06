"""Fused projection stages (analysisgnn_b200/fused.py, nn/layers.py::MLP): [LayerNorm ->] [Dropout ->] Linear [-> ReLU]
for groups of independent inputs -- against torch's own modules in fp64, forward and every gradient, in the fp16 operand
form (>= 16 384 rows: LayerNorm writes the operand pair directly) and the TF32 form (small inputs); the counter-based
dropout mask (same mask forward and backward, new mask per step); amax tags (no extra passes over tagged tensors)."""
import pytest
import torch
import torch.nn as nn

from analysisgnn_b200 import _lib, fused, linalg
from analysisgnn_b200.nn.layers import MLP, LayerNorm, Linear
from tests.util import DEV, rel_err

pytestmark = pytest.mark.gpu

TOL = 4e-6      # fp32-GEMM level (tests/test_gemm_gpu.py) through one LayerNorm + one projection


def _ref_stage(x, w, b, norm, relu, mask=None, p=0.0, relu_mask=None):
    """``relu_mask``: the CUDA side's activation pattern -- a pre-activation within rounding of zero may land on either
    side in two correct implementations, and the gradient of that unit is then 0 or 1 (tests/util.py)."""
    y = x
    if norm is not None:
        y = nn.functional.layer_norm(y, (y.shape[-1],), norm[0], norm[1], norm[2])
    if mask is not None:
        y = y * mask / (1.0 - p)
    y = y @ w.t() + (b if b is not None else 0.0)
    if relu_mask is not None:
        return y * relu_mask
    return y.relu() if relu else y


@pytest.mark.parametrize("rows", [[20000, 3100, 0], [300, 77, 5]])
@pytest.mark.parametrize("with_norm,relu", [(True, True), (True, False), (False, True)])
def test_stage_group_matches_torch(rows, with_norm, relu):
    g = torch.Generator().manual_seed(11)
    k, n = 256, 128
    xs = [torch.randn(r, k, generator=g) * 2 + 0.3 for r in rows]
    ws = [torch.randn(n, k, generator=g) * 0.05 for _ in rows]
    bs = [torch.randn(n, generator=g) for _ in rows]
    norms = [(torch.rand(k, generator=g) + 0.5, torch.randn(k, generator=g) * 0.1, 1e-5) if with_norm else None
             for _ in rows]
    gys = [torch.randn(r, n, generator=g) * 1e-2 for r in rows]

    def leaf(t, dev, dt):
        return t.to(dev, dt).requires_grad_(True)

    ref_in = [[leaf(x, "cpu", torch.float64), leaf(w, "cpu", torch.float64), leaf(b, "cpu", torch.float64)] +
              ([leaf(nm[0], "cpu", torch.float64), leaf(nm[1], "cpu", torch.float64)] if nm else [])
              for x, w, b, nm in zip(xs, ws, bs, norms)]
    dev_in = [[leaf(x, DEV, torch.float32), leaf(w, DEV, torch.float32), leaf(b, DEV, torch.float32)] +
              ([leaf(nm[0], DEV, torch.float32), leaf(nm[1], DEV, torch.float32)] if nm else [])
              for x, w, b, nm in zip(xs, ws, bs, norms)]
    before = linalg.stats.get("gemm_launches", 0)
    outs = fused.stage_group([t[0] for t in dev_in], [t[1] for t in dev_in], [t[2] for t in dev_in],
                             [(t[3], t[4], nm[2]) if nm else None for t, nm in zip(dev_in, norms)], relu=relu)
    assert linalg.stats["gemm_launches"] - before == 1          # all members in one grouped launch
    live = [(o, gy) for o, gy in zip(outs, gys) if o.numel()]
    torch.autograd.backward([o for o, _ in live], [gy.to(DEV) for _, gy in live])
    for t, gy, nm, o in zip(ref_in, gys, norms, outs):          # reference gradients on the CUDA side's ReLU pattern
        y = _ref_stage(t[0], t[1], t[2], (t[3], t[4], nm[2]) if nm else None, relu,
                       relu_mask=(o.detach() > 0).double().cpu() if relu else None)
        if y.numel():
            y.backward(gy.double())
    for o, t_ref, t_dev, nm in zip(outs, ref_in, dev_in, norms):
        if o.shape[0] == 0:
            continue
        want = _ref_stage(t_ref[0], t_ref[1], t_ref[2], (t_ref[3], t_ref[4], nm[2]) if nm else None, relu)
        assert rel_err(o, want) <= TOL
        assert float(linalg.known_amax(o)) == float(o.abs().max())      # the epilogue's amax tag is exact
        for a, r, what in zip(t_dev, t_ref, ("x", "w", "b", "gamma", "beta")):
            err = rel_err(a.grad, r.grad)
            if err > 2 * TOL:                                     # where: which rows / columns are off
                d = (a.grad.detach().double().cpu() - r.grad).abs()
                bad = (d > 2 * TOL * float(r.grad.abs().max())).nonzero()
                raise AssertionError(f"{what} of a member with {o.shape[0]} rows: err {err:.3e}, {bad.shape[0]} bad "
                                     f"entries, rows {bad[:, 0].min().item()}..{bad[:, 0].max().item()}, cols "
                                     f"{bad[:, -1].min().item()}..{bad[:, -1].max().item()}, first {bad[:5].tolist()}")


def test_counter_dropout_masks():
    """keep(element) depends on (seed, step, site, index) only: the same call repeats its mask (that is how the
    backward recomputes it), another site or another step draws a new one; the keep rate is 1 - p."""
    linalg.begin_step()
    x = torch.ones(4096, 256, device=DEV)
    p = 0.3
    y1 = fused.dropout_apply(x, p, 1)
    y2 = fused.dropout_apply(x, p, 1)
    y3 = fused.dropout_apply(x, p, 2)
    assert torch.equal(y1, y2) and not torch.equal(y1, y3)
    keep = float((y1 != 0).float().mean())
    assert abs(keep - (1 - p)) < 3e-3
    assert torch.allclose(y1[y1 != 0], torch.full((1,), 1 / (1 - p), device=DEV))
    # rows and columns are both mixed: no constant row / column pattern
    assert float((y1 != 0).float().mean(0).std()) < 0.02 and float((y1 != 0).float().mean(1).std()) < 0.05
    linalg.begin_step()                                          # next step
    assert not torch.equal(fused.dropout_apply(x, p, 1), y1)
    am = torch.zeros(1, device=DEV)
    y = fused.dropout_apply(torch.randn(1000, 64, device=DEV), p, 5, am)
    assert float(am) == float(y.abs().max())


@pytest.mark.parametrize("rows", [20000, 500])
def test_stage_with_dropout_uses_one_mask_forward_and_backward(rows):
    g = torch.Generator().manual_seed(12)
    k, n, p = 128, 64, 0.3
    x, w, b = torch.randn(rows, k, generator=g), torch.randn(n, k, generator=g) * 0.1, torch.randn(n, generator=g)
    gam, bet = torch.rand(k, generator=g) + 0.5, torch.randn(k, generator=g) * 0.1
    gy = torch.randn(rows, n, generator=g) * 1e-2
    for with_norm in (True, False):
        linalg.begin_step()                                      # the stage below gets dropout site 1 of this step
        mask = (fused.dropout_apply(torch.ones(rows, k, device=DEV), p, 1) != 0).double().cpu()
        t_ref = [t.double().requires_grad_(True) for t in (x, w, b, gam, bet)]
        want = _ref_stage(t_ref[0], t_ref[1], t_ref[2], (t_ref[3], t_ref[4], 1e-5) if with_norm else None, False, mask, p)
        want.backward(gy.double())
        t_dev = [t.to(DEV).requires_grad_(True) for t in (x, w, b, gam, bet)]
        out = fused.stage_group([t_dev[0]], [t_dev[1]], [t_dev[2]], [(t_dev[3], t_dev[4], 1e-5) if with_norm else None],
                                dropout=p, training=True)[0]
        out.backward(gy.to(DEV))
        assert rel_err(out, want) <= TOL
        for a, r, what in list(zip(t_dev, t_ref, ("x", "w", "b", "gamma", "beta")))[:5 if with_norm else 3]:
            assert rel_err(a.grad, r.grad) <= 2 * TOL, (what, with_norm)
        # eval mode: no dropout
        out_eval = fused.stage_group([t_dev[0]], [t_dev[1]], [t_dev[2]], [(t_dev[3], t_dev[4], 1e-5) if with_norm else None],
                                     dropout=p, training=False)[0]
        assert rel_err(out_eval, _ref_stage(t_ref[0], t_ref[1], t_ref[2], (t_ref[3], t_ref[4], 1e-5) if with_norm else None,
                                            False)) <= TOL


def _mlp_pair(kind):
    torch.manual_seed(5)
    if kind == "project_enc":        # analysis.py:474-485
        mods = lambda L, N: [N(512), L(512, 256), nn.ReLU(), N(256), nn.Dropout(0.3), L(256, 128), nn.ReLU(), N(128),
                             nn.Dropout(0.3), L(128, 128)]
        width = 512
    elif kind == "head":             # analysis.py:486-496 (185 classes: odd output width)
        mods = lambda L, N: [L(128, 64), nn.ReLU(), N(64), L(64, 185)]
        width = 128
    else:                            # project_dict, analysis.py:429-443 (25 + 128 inputs: odd input width)
        mods = lambda L, N: [L(153, 256), nn.ReLU(), N(256), nn.Dropout(0.3), L(256, 256)]
        width = 153
    ref = nn.Sequential(*mods(nn.Linear, nn.LayerNorm))
    net = MLP(*mods(Linear, LayerNorm))
    net.load_state_dict(ref.state_dict())
    return ref.double().eval(), net.to(DEV).eval(), width


@pytest.mark.parametrize("kind", ["project_enc", "head", "project_dict"])
@pytest.mark.parametrize("rows", [20000, 333])
def test_mlp_container_matches_sequential(kind, rows):
    """Same children, same state_dict, same numbers as nn.Sequential of torch modules (eval mode), all gradients."""
    ref, net, width = _mlp_pair(kind)
    assert list(ref.state_dict()) == list(net.state_dict())
    assert net.stages() is not None
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, width, generator=g)
    x1 = x.double().requires_grad_(True)
    y1 = ref(x1)
    gy = torch.randn(y1.shape, generator=g) * 1e-2
    y1.backward(gy.double())
    x2 = x.to(DEV).requires_grad_(True)
    y2 = net(x2)
    y2.backward(gy.to(DEV))
    assert rel_err(y2, y1) <= 3 * TOL                              # three stages deep
    assert rel_err(x2.grad, x1.grad) <= 3 * TOL
    for (name, p1), p2 in zip(ref.named_parameters(), net.parameters()):
        assert rel_err(p2.grad, p1.grad) <= 3 * TOL, name


def test_mlp_group_of_heads_is_one_launch_per_stage():
    """clf_dict: three heads on the same input run stage by stage in one grouped launch each (forward), and the amax
    tags keep the backward free of extra passes over the gradients."""
    torch.manual_seed(2)
    heads = [MLP(Linear(128, 64), nn.ReLU(), LayerNorm(64), Linear(64, c)).to(DEV) for c in (4, 50, 185)]
    x = torch.randn(20000, 128, device=DEV, requires_grad=True)
    before = linalg.stats.get("gemm_launches", 0)
    outs = MLP.forward_group(heads, [x, x, x])
    assert linalg.stats["gemm_launches"] - before == 2
    assert [o.shape[1] for o in outs] == [4, 50, 185]
    for h, o in zip(heads, outs):
        want = nn.Sequential(*h.children())(x)                     # module by module (unfused route)
        assert rel_err(o, want) <= 2 * TOL
    passes = linalg.stats.get("amax_passes", 0)
    before = linalg.stats.get("gemm_launches", 0)
    torch.autograd.backward(outs, [torch.randn_like(o) for o in outs])
    # backward: per stage one grouped grad-weight + one grouped grad-input launch
    assert linalg.stats["gemm_launches"] - before <= 4
    assert x.grad is not None and x.grad.shape == x.shape
    # amax passes: the three incoming gradients (random tensors without a tag) and the three zero-padded copies of
    # the odd-width head weights -- none over the intermediate gradients, which carry their producers' tags
    assert linalg.stats.get("amax_passes", 0) - passes <= 6


@pytest.mark.parametrize("b,t,h,n_dir", [(700, 30, 128, 2), (5, 1, 128, 1), (33, 7, 64, 2)])
def test_shifted_split_is_the_previous_state_operand(b, t, h, n_dir):
    """agnn_split_f16_shifted: the GRU's h_{t-1} (forward) / h_{t+1} (reverse) operand straight from the output
    [B, T, n_dir H] -- bit-equal to splitting the materialised shifted copy."""
    g = torch.Generator().manual_seed(5)
    out = (torch.rand(b, t, n_dir * h, generator=g) * 2 - 1).to(DEV)
    one = linalg.const_amax(out.device, 1.0)
    out2 = out.view(b * t, n_dir * h)
    for d in range(n_dir):
        hd = out[:, :, d * h:(d + 1) * h]
        prev = torch.zeros(b, t, h, device=DEV)
        if t > 1:
            if d == 0:
                prev[:, 1:] = hd[:, :-1]
            else:
                prev[:, :-1] = hd[:, 1:]
        want = linalg.split_f16(prev.reshape(b * t, h), one)
        got = linalg.split_f16(out2[:, d * h:(d + 1) * h], one, shift=(t, 1 if d == 0 else -1))
        assert torch.equal(got.hi, want.hi) and torch.equal(got.lo, want.lo)


@pytest.mark.parametrize("k,n,f,scale", [(1, 256, 256, 1.0), (3, 256, 256, 1.0), (9, 128, 64, 0.5)])
def test_sage_weight_assembly(k, n, f, scale):
    """agnn_sage_weights (one launch) against stack / sum / cat, forward and the parameter gradients."""
    from analysisgnn_b200 import ops
    g = torch.Generator().manual_seed(9)
    mk = lambda *s: [torch.randn(*s, generator=g).to(DEV).requires_grad_() for _ in range(k)]
    wr, wl, bl = mk(n, f), mk(n, f), mk(n)
    wcat, bias = ops.sage_weights(wr, wl, bl, scale)
    ref_w = torch.cat([torch.stack([w.double() for w in wr]).sum(0)] + [w.double() for w in wl], dim=1) * scale
    ref_b = torch.stack([b.double() for b in bl]).sum(0) * scale
    assert rel_err(wcat, ref_w) < 1e-6 and rel_err(bias, ref_b) < 1e-6
    gw, gb = torch.randn(wcat.shape, generator=g).to(DEV), torch.randn(n, generator=g).to(DEV)
    grads = torch.autograd.grad([wcat, bias], wr + wl + bl, [gw, gb])
    for j in range(k):
        assert torch.equal(grads[j], gw[:, :f] * scale)
        assert torch.equal(grads[k + j], gw[:, (j + 1) * f:(j + 2) * f] * scale)
        assert torch.equal(grads[2 * k + j], gb * scale)
