"""CUDA in-tree layers (analysisgnn_b200/nn/intree.py) vs the pinned oracle
(oracle/intree.py) and vs the golden vectors generated from the reference's own
modules (tests/golden/intree_seed*.pt).  fp32, tolerance 1e-5 relative."""
import pytest
import torch

from analysisgnn_b200 import synth
from analysisgnn_b200 import nn as ann
from oracle import intree as oi
from tests.util import (DEV, FP32_REL, ActivationPatterns, assert_close, feeds_relu,
                        first_seed_with_equal_patterns, golden_intree, grads_of)

pytestmark = pytest.mark.gpu


def _dev(v):
    return v.to(DEV) if torch.is_tensor(v) else v


def _compare(net, ref_out, ref_pg, ref_xg, out, x, tol=FP32_REL):
    assert_close(out, ref_out, tol, "forward")
    pg, ig = grads_of(net, out, [x])
    assert set(pg) == set(ref_pg), set(pg) ^ set(ref_pg)
    for k in pg:
        assert_close(pg[k], ref_pg[k], tol, f"grad {k}")
    assert_close(ig[0], ref_xg, tol, "grad x")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_golden_sage_and_empty_branch(seed):
    rec = golden_intree(seed)
    conv = ann.SageConvScatter(8, 16).to(DEV)
    conv.load_state_dict(rec["sage"]["state"])
    x = rec["batch"]["x"].to(DEV).requires_grad_(True)
    ei = rec["sage"]["edge_index"].to(DEV)
    _compare(conv, rec["sage"]["out"], rec["sage"]["param_grads"], rec["sage"]["x_grad"], conv(x, ei), x)
    e0 = ei[:, :0].contiguous()
    _compare(conv, rec["sage_empty"]["out"], rec["sage_empty"]["param_grads"], rec["sage_empty"]["x_grad"],
             conv(x, e0), x)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_golden_heteroconv(seed):
    rec = golden_intree(seed)
    b = rec["batch"]
    hc = ann.HeteroConv(8, 16, rec["etypes"]).to(DEV)
    hc.load_state_dict(rec["hetero"]["state"])
    x = b["x"].to(DEV).requires_grad_(True)
    out = hc(x, b["edge_index"].to(DEV), b["edge_type"].to(DEV))
    _compare(hc, rec["hetero"]["out"], rec["hetero"]["param_grads"], rec["hetero"]["x_grad"], out, x)


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_golden_metricalgnn(seed, mode):
    rec = golden_intree(seed)
    b = {k: _dev(v) for k, v in rec["batch"].items()}
    g = rec[f"metrical_{mode}"]
    net = ann.MetricalGNN(8, 16, 16, rec["etypes"], num_layers=3, dropout=0.0, metrical=True).to(DEV)
    net.load_state_dict(g["state"])
    net.train(mode == "train")
    x = b["x"].clone().requires_grad_(True)
    out = net(x, b["edge_index"], b["edge_type"], b["beat_nodes"], b["measure_nodes"], b["beat_edges"],
              b["measure_edges"], beat_lengths=b["beat_lengths"], measure_lengths=b["measure_lengths"])
    # the GRU / BatchNorm inside are library kernels (cuDNN / ATen) outside this repo's kernels: BatchNorm's
    # division by the batch deviation of an 80..116-row golden amplifies their rounding, hence 1e-4 here
    if mode == "eval":          # cuDNN cannot run the RNN backward in eval mode (a cuDNN restriction)
        assert_close(out, g["out"], 10 * FP32_REL, "forward")
        return
    _compare(net, g["out"], g["param_grads"], g["x_grad"], out, x, tol=10 * FP32_REL)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_golden_plain_metricalgnn(seed):
    rec = golden_intree(seed)
    b = rec["batch"]
    net = ann.MetricalGNN(8, 16, 16, rec["etypes"], num_layers=2, dropout=0.0, metrical=False).to(DEV)
    net.load_state_dict(rec["plain"]["state"])
    x = b["x"].to(DEV).requires_grad_(True)
    out = net(x, b["edge_index"].to(DEV), b["edge_type"].to(DEV))
    _compare(net, rec["plain"]["out"], rec["plain"]["param_grads"], rec["plain"]["x_grad"], out, x)


@pytest.mark.parametrize("reduction", ["mean", "sum"])
@pytest.mark.parametrize("f_in,f_out", [(64, 64), (256, 256), (128, 512)])
def test_heteroconv_vs_oracle(reduction, f_in, f_out):
    b = synth.intree_batch(4, 150, 21, in_features=f_in, metrical=False)
    torch.manual_seed(1)
    ref = oi.HeteroConv(f_in, f_out, b["etypes"], reduction=reduction)
    net = ann.HeteroConv(f_in, f_out, b["etypes"], reduction=reduction).to(DEV)
    net.load_state_dict(ref.state_dict())
    x1 = b["x"].clone().requires_grad_(True)
    o1 = ref(x1, b["edge_index"], b["edge_type"])
    pg, ig = grads_of(ref, o1, [x1])
    x2 = b["x"].to(DEV).requires_grad_(True)
    o2 = net(x2, b["edge_index"].to(DEV), b["edge_type"].to(DEV))
    _compare(net, o1, pg, ig[0], o2, x2)


def test_heteroconv_noncontiguous_codes_and_unknown_types():
    """etypes whose codes are not 0..R-1, plus edges of a type the layer does not know."""
    b = synth.intree_batch(2, 120, 22, in_features=32, metrical=False)
    remap = torch.tensor([5, 9, 2, 11, 0, 7, 3])
    et = remap[b["edge_type"]]
    etypes = {k: int(remap[v]) for k, v in b["etypes"].items() if k != "during_rev"}   # code 7 now unknown
    torch.manual_seed(2)
    ref = oi.HeteroConv(32, 48, etypes)
    net = ann.HeteroConv(32, 48, etypes).to(DEV)
    net.load_state_dict(ref.state_dict())
    x1 = b["x"].clone().requires_grad_(True)
    o1 = ref(x1, b["edge_index"], et)
    pg, ig = grads_of(ref, o1, [x1])
    x2 = b["x"].to(DEV).requires_grad_(True)
    o2 = net(x2, b["edge_index"].to(DEV), et.to(DEV))
    _compare(net, o1, pg, ig[0], o2, x2)


def test_sage_edge_features_and_neigh_feats():
    b = synth.intree_batch(2, 100, 23, in_features=32, metrical=False)
    ei = b["edge_index"][:, b["edge_type"] == 2]
    torch.manual_seed(3)
    ef = torch.randn(ei.shape[1], 6)
    nf = torch.randn(b["x"].shape[0], 32)
    ref = oi.SageConvScatter(32, 40, in_edge_features=6)
    net = ann.SageConvScatter(32, 40, in_edge_features=6).to(DEV)
    net.load_state_dict(ref.state_dict())
    x1 = b["x"].clone().requires_grad_(True)
    o1 = ref(x1, ei, ef, nf)
    pg, ig = grads_of(ref, o1, [x1])
    x2 = b["x"].to(DEV).requires_grad_(True)
    o2 = net(x2, ei.to(DEV), ef.to(DEV), nf.to(DEV))
    _compare(net, o1, pg, ig[0], o2, x2)


@pytest.mark.parametrize("name", ["ResGatedGraphConv", "RelEdgeConv"])
def test_alternative_conv_blocks(name):
    """The other ``conv_block``s of the reference (gnn.py:79-106, 212-258), alone and inside HeteroConv."""
    b = synth.intree_batch(2, 110, 25, in_features=32, metrical=False)
    ei = b["edge_index"][:, b["edge_type"] == 2]
    torch.manual_seed(6)
    ref = getattr(oi, name)(32, 48)
    net = getattr(ann, name)(32, 48).to(DEV)
    net.load_state_dict(ref.state_dict())
    x1 = b["x"].clone().requires_grad_(True)
    o1 = ref(x1, ei)
    pg, ig = grads_of(ref, o1, [x1])
    x2 = b["x"].to(DEV).requires_grad_(True)
    _compare(net, o1, pg, ig[0], net(x2, ei.to(DEV)), x2)
    ref_h = oi.HeteroConv(32, 48, b["etypes"], module=getattr(oi, name))
    net_h = ann.HeteroConv(32, 48, b["etypes"], module=getattr(ann, name)).to(DEV)
    net_h.load_state_dict(ref_h.state_dict())
    x1 = b["x"].clone().requires_grad_(True)
    o1 = ref_h(x1, b["edge_index"], b["edge_type"])
    pg, ig = grads_of(ref_h, o1, [x1])
    x2 = b["x"].to(DEV).requires_grad_(True)
    _compare(net_h, o1, pg, ig[0], net_h(x2, b["edge_index"].to(DEV), b["edge_type"].to(DEV)), x2)


@pytest.mark.parametrize("uniform", [True, False])
def test_metrical_conv_layer_uniform_and_ragged(uniform):
    torch.manual_seed(4)
    n_graphs, t = 3, 10
    sizes = [t] * n_graphs if uniform else [7, 12, 9]
    n_m = sum(sizes)
    n = 90
    note_to_m = torch.sort(torch.randint(0, n_m, (n,)))[0]
    edges = torch.stack((torch.arange(n), note_to_m))
    lengths = torch.tensor(sizes) if uniform else torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)))
    ref = oi.MetricalConvLayer(16, 24, activation=torch.relu, dropout=0.0)
    net = ann.MetricalConvLayer(16, 24, activation=torch.relu, dropout=0.0).to(DEV)
    net.load_state_dict(ref.state_dict())
    xm, x = torch.randn(n_m, 16), torch.randn(n, 16)
    x1 = x.clone().requires_grad_(True)
    o1, h1 = ref(xm, x1, edges, lengths)
    x2 = x.to(DEV).requires_grad_(True)
    o2, h2 = net(xm.to(DEV), x2, edges.to(DEV), lengths.to(DEV))
    assert_close(o2, o1, 5 * FP32_REL, "notes")
    assert_close(h2, h1, 5 * FP32_REL, "metrical")
    pg, ig = grads_of(ref, o1 + 0, [x1])
    pg2, ig2 = grads_of(net, o2 + 0, [x2])
    for k in pg:
        assert_close(pg2[k], pg[k], 5 * FP32_REL, k)
    assert_close(ig2[0], ig[0], 5 * FP32_REL, "x")


def test_metricalgnn_config4_shape_small_batch():
    """BASELINE config 4 architecture (4 layers, hidden 512, 7 relations, metrical) on a small batch."""
    args = ("edge_index", "edge_type", "beat_nodes", "measure_nodes", "beat_edges", "measure_edges")
    kw = ("beat_lengths", "measure_lengths")

    def run(seed):
        b = synth.intree_batch(3, 120, 24 + seed, in_features=64, metrical=True)
        torch.manual_seed(5)
        ref = oi.MetricalGNN(64, 512, 512, b["etypes"], num_layers=4, dropout=0.0, metrical=True)
        net = ann.MetricalGNN(64, 512, 512, b["etypes"], num_layers=4, dropout=0.0, metrical=True).to(DEV)
        net.load_state_dict(ref.state_dict())
        pr, pn = ActivationPatterns(ref, feeds_relu), ActivationPatterns(net, feeds_relu)
        x1 = b["x"].clone().requires_grad_(True)
        o1 = ref(x1, *[b[k] for k in args], **{k: b[k] for k in kw})
        x2 = b["x"].to(DEV).requires_grad_(True)
        o2 = net(x2, *[b[k].to(DEV) for k in args], **{k: b[k].to(DEV) for k in kw})
        assert_close(o2, o1, 5 * FP32_REL, "forward")
        mism = pr.mismatches(pn)
        pr.close(), pn.close()
        if mism:
            return mism, None
        pg, ig = grads_of(ref, o1, [x1])
        _compare(net, o1, pg, ig[0], o2, x2, tol=5 * FP32_REL)
        return 0, None

    first_seed_with_equal_patterns(run)


def test_state_dict_keys_are_the_reference_keys():
    rec = golden_intree(0)
    net = ann.MetricalGNN(8, 16, 16, rec["etypes"], num_layers=3, dropout=0.0, metrical=True)
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in rec["metrical_train"]["state"].items()}
