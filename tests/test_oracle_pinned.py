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


# ---------------------------------------------------------------- GATConvLayer / OnsetEmbedding (gnn.py:154-209, 294-311)

ATTENTION_PARAMS = ("attnl", "attnr", "el.weight", "el.bias", "er.weight", "er.bias")


def golden_convblocks():
    import os
    from tests.util import GOLDEN
    return torch.load(os.path.join(GOLDEN, "convblocks.pt"), weights_only=False)


def check_gat_grads(pg, want, what):
    """``linear`` carries the whole gradient; the attention parameters only see the rounding residue of
    ``1 - sum(softmax)`` (1e-8 of the ``linear`` gradients in the reference itself): asserted as noise, not compared."""
    scale = float(want["linear.weight"].abs().max())
    for k in ("linear.weight", "linear.bias"):
        assert_close(pg[k], want[k], 5e-6, f"{what} grad {k}")
    for k in ATTENTION_PARAMS:
        assert float(want[k].abs().max()) <= 1e-6 * scale, k
        if k in pg and pg[k] is not None:
            assert float(pg[k].abs().max()) <= 1e-6 * scale, k


@pytest.mark.parametrize("seed", [0, 1])
def test_gat_matches_reference_golden(seed):
    rec = golden_convblocks()[f"gat{seed}"]
    x = rec["x"].clone().requires_grad_(True)
    gat = intree.GATConvLayer(12, 20, num_heads=3, dropout=0.0)
    gat.load_state_dict(rec["state"])
    out = gat(x, rec["edge_index"])
    assert_close(out, rec["out"], TOL, "forward")
    pg, ig = grads_of(gat, out, [x])
    check_gat_grads(pg, rec["param_grads"], "gat")
    assert_close(ig[0], rec["x_grad"], 5e-6, "grad x")
    # softmax over the heads, then their mean: attention dropout cannot change the output
    gat.train()
    gat.attndrop.p = 0.5
    assert_close(gat(x, rec["edge_index"]), rec["out_train_dropout"], TOL, "train mode, attention dropout 0.5")
    assert_close(rec["out_train_dropout"], rec["out"], TOL, "the reference itself")


@pytest.mark.parametrize("seed", [0, 1])
@pytest.mark.parametrize("loops", [True, False])
def test_onset_embedding_matches_reference_golden(seed, loops):
    rec = golden_convblocks()[f"onset{seed}_{int(loops)}"]
    x = rec["x"].clone().requires_grad_(True)
    emb = intree.OnsetEmbedding(12, 20, add_self_loops=loops)
    emb.load_state_dict(rec["state"])
    _check(emb, rec, emb(x, rec["edge_index"]), x)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_gat_and_onset_embedding_match_live_reference():
    from analysisgnn_b200 import synth
    gnn, _ = ref_loader.load_core()
    b = synth.intree_batch(3, 45, 8, in_features=16, metrical=False)
    ei = b["edge_index"][:, b["edge_type"] <= 1]
    torch.manual_seed(2)
    ref, mine = gnn.GATConvLayer(16, 24, num_heads=4, dropout=0.0), intree.GATConvLayer(16, 24, num_heads=4, dropout=0.0)
    mine.load_state_dict(ref.state_dict())
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    x1, x2 = b["x"].clone().requires_grad_(True), b["x"].clone().requires_grad_(True)
    o1, o2 = ref(x1, ei), mine(x2, ei)
    assert_close(o2, o1, TOL, "gat forward")
    g1, i1 = grads_of(ref, o1, [x1])
    g2, i2 = grads_of(mine, o2, [x2])
    check_gat_grads(g2, g1, "gat")
    assert_close(i2[0], i1[0], 5e-6, "gat x grad")
    ref, mine = gnn.OnsetEmbedding(16, 24), intree.OnsetEmbedding(16, 24)
    mine.load_state_dict(ref.state_dict())
    x1, x2 = b["x"].clone().requires_grad_(True), b["x"].clone().requires_grad_(True)
    o1, o2 = ref(x1, ei), mine(x2, ei)
    assert_close(o2, o1, TOL, "onset embedding forward")
    g1, i1 = grads_of(ref, o1, [x1])
    g2, i2 = grads_of(mine, o2, [x2])
    for k in g1:
        assert_close(g2[k], g1[k], 5e-6, k)
    assert_close(i2[0], i1[0], 5e-6, "onset embedding x grad")
