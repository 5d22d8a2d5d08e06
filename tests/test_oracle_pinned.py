"""Pins the CPU oracle (oracle/intree.py) to the reference's own code:

* against the committed golden vectors generated from the reference modules
  (tests/golden/make_golden.py), and
* where /root/reference is present (the build container), against the reference
  modules executed live.
"""
import pytest
import torch

from oracle import intree, ref_loader
from tests.util import assert_close, golden_intree, grads_of

TOL = 2e-6   # same arithmetic, different kernels (index_add_ vs scatter_add_): rounding only


def _check(module, rec, out, x):
    assert_close(out, rec["out"], TOL, "forward")
    pg, ig = grads_of(module, out, [x])
    assert set(pg) == set(rec["param_grads"]), set(pg) ^ set(rec["param_grads"])
    for k, g in pg.items():
        assert_close(g, rec["param_grads"][k], 5e-6, f"grad {k}")
    assert_close(ig[0], rec["x_grad"], 5e-6, "grad x")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_sage_matches_reference_golden(seed):
    rec = golden_intree(seed)
    x = rec["batch"]["x"].clone().requires_grad_(True)
    conv = intree.SageConvScatter(8, 16)
    conv.load_state_dict(rec["sage"]["state"])
    _check(conv, rec["sage"], conv(x, rec["sage"]["edge_index"]), x)
    _check(conv, rec["sage_empty"], conv(x, rec["sage"]["edge_index"][:, :0]), x)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_heteroconv_matches_reference_golden(seed):
    rec = golden_intree(seed)
    b = rec["batch"]
    x = b["x"].clone().requires_grad_(True)
    hc = intree.HeteroConv(8, 16, rec["etypes"])
    hc.load_state_dict(rec["hetero"]["state"])
    _check(hc, rec["hetero"], hc(x, b["edge_index"], b["edge_type"]), x)


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_metricalgnn_matches_reference_golden(seed, mode):
    rec = golden_intree(seed)
    b = rec["batch"]
    x = b["x"].clone().requires_grad_(True)
    net = intree.MetricalGNN(8, 16, 16, rec["etypes"], num_layers=3, dropout=0.0, metrical=True)
    net.load_state_dict(rec[f"metrical_{mode}"]["state"])
    net.train(mode == "train")
    out = net(x, b["edge_index"], b["edge_type"], b["beat_nodes"], b["measure_nodes"], b["beat_edges"],
              b["measure_edges"], beat_lengths=b["beat_lengths"], measure_lengths=b["measure_lengths"])
    _check(net, rec[f"metrical_{mode}"], out, x)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_plain_metricalgnn_matches_reference_golden(seed):
    rec = golden_intree(seed)
    b = rec["batch"]
    x = b["x"].clone().requires_grad_(True)
    net = intree.MetricalGNN(8, 16, 16, rec["etypes"], num_layers=2, dropout=0.0, metrical=False)
    net.load_state_dict(rec["plain"]["state"])
    _check(net, rec["plain"], net(x, b["edge_index"], b["edge_type"]), x)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference():
    from analysisgnn_b200 import synth
    gnn, hgnn = ref_loader.load_core()
    b = synth.intree_batch(3, 70, 11, voices=4, in_features=12, reverse=True, metrical=True)
    torch.manual_seed(5)
    ref = hgnn.MetricalGNN(12, 20, 20, b["etypes"], num_layers=4, dropout=0.0, metrical=True,
                           conv_block=gnn.SageConvScatter)
    mine = intree.MetricalGNN(12, 20, 20, b["etypes"], num_layers=4, dropout=0.0, metrical=True)
    mine.load_state_dict(ref.state_dict())
    args = (b["edge_index"], b["edge_type"], b["beat_nodes"], b["measure_nodes"], b["beat_edges"], b["measure_edges"])
    kw = dict(beat_lengths=b["beat_lengths"], measure_lengths=b["measure_lengths"])
    x1 = b["x"].clone().requires_grad_(True)
    x2 = b["x"].clone().requires_grad_(True)
    o1, o2 = ref(x1, *args, **kw), mine(x2, *args, **kw)
    assert_close(o2, o1, TOL, "forward")
    g1, i1 = grads_of(ref, o1, [x1])
    g2, i2 = grads_of(mine, o2, [x2])
    assert set(g1) == set(g2)
    for k in g1:
        assert_close(g2[k], g1[k], 5e-6, k)
    assert_close(i2[0], i1[0], 5e-6, "x grad")
    # state_dict keys and shapes are the reference's
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in mine.state_dict().items()}


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("name", ["ResGatedGraphConv", "RelEdgeConv"])
def test_alternative_conv_blocks_match_live_reference(name):
    from analysisgnn_b200 import synth
    gnn, _ = ref_loader.load_core()
    b = synth.intree_batch(2, 60, 3, in_features=12, metrical=False)
    ei = b["edge_index"][:, b["edge_type"] == 2]
    torch.manual_seed(1)
    ref = getattr(gnn, name)(12, 20)
    mine = getattr(intree, name)(12, 20)
    mine.load_state_dict(ref.state_dict())
    x1 = b["x"].clone().requires_grad_(True)
    x2 = b["x"].clone().requires_grad_(True)
    o1, o2 = ref(x1, ei), mine(x2, ei)
    assert_close(o2, o1, TOL, "forward")
    g1, i1 = grads_of(ref, o1, [x1])
    g2, i2 = grads_of(mine, o2, [x2])
    for k in g1:
        assert_close(g2[k], g1[k], 5e-6, k)
    assert_close(i2[0], i1[0], 5e-6, "x grad")
