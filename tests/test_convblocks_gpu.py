"""GATConvLayer and OnsetEmbedding (analysisgnn/models/core/gnn.py:154-209, 294-311) on the CUDA path against golden
vectors the REFERENCE produced (tests/golden/convblocks.pt, tests/golden/make_golden.py) and against the pinned
oracle (oracle/intree.py) at a larger width; fp32, 1e-5."""
import os

import pytest
import torch

from analysisgnn_b200 import nn as ann
from analysisgnn_b200 import synth
from oracle import intree as oi
from tests.util import DEV, FP32_REL, GOLDEN, assert_close, grads_of

pytestmark = pytest.mark.gpu

ATTENTION_PARAMS = ("attnl", "attnr", "el.weight", "el.bias", "er.weight", "er.bias")


def golden():
    return torch.load(os.path.join(GOLDEN, "convblocks.pt"), weights_only=False)


def check_gat(net, out, x, want_out, want_pg, want_xg):
    assert_close(out, want_out, FP32_REL, "forward")
    pg, ig = grads_of(net, out, [x])
    for k in ("linear.weight", "linear.bias"):
        assert_close(pg[k], want_pg[k], FP32_REL, f"grad {k}")
    assert_close(ig[0], want_xg, FP32_REL, "grad x")
    # the attention parameters: noise in the reference (1e-8 of the linear gradients), untouched here
    scale = float(want_pg["linear.weight"].abs().max())
    for k in ATTENTION_PARAMS:
        assert float(want_pg[k].abs().max()) <= 1e-6 * scale
        assert k not in pg or float(pg[k].abs().max()) <= 1e-6 * scale


@pytest.mark.parametrize("seed", [0, 1])
def test_gat_golden(seed):
    rec = golden()[f"gat{seed}"]
    net = ann.GATConvLayer(12, 20, num_heads=3, dropout=0.0).to(DEV)
    net.load_state_dict(rec["state"])
    x = rec["x"].to(DEV).requires_grad_(True)
    ei = rec["edge_index"].to(DEV)
    check_gat(net, net(x, ei), x, rec["out"], rec["param_grads"], rec["x_grad"])
    net.train()                                   # the reference's output under attention dropout 0.5 is the same
    assert_close(net(x, ei), rec["out_train_dropout"], FP32_REL, "train mode")
    assert_close(net(x, ei[:, :0].contiguous()), net.linear(x), 0.0, "no edges: h itself")


@pytest.mark.parametrize("seed", [0, 1])
@pytest.mark.parametrize("loops", [True, False])
def test_onset_embedding_golden(seed, loops):
    rec = golden()[f"onset{seed}_{int(loops)}"]
    net = ann.OnsetEmbedding(12, 20, add_self_loops=loops).to(DEV)
    net.load_state_dict(rec["state"])
    x = rec["x"].to(DEV).requires_grad_(True)
    out = net(x, rec["edge_index"].to(DEV))
    assert_close(out, rec["out"], FP32_REL, "forward")
    pg, ig = grads_of(net, out, [x])
    assert set(pg) == set(rec["param_grads"])
    for k in pg:
        assert_close(pg[k], rec["param_grads"][k], FP32_REL, f"grad {k}")
    assert_close(ig[0], rec["x_grad"], FP32_REL, "grad x")


def test_against_the_oracle_at_width_128():
    b = synth.intree_batch(3, 200, 41, in_features=128, metrical=False)
    ei = b["edge_index"][:, b["edge_type"] <= 2]
    torch.manual_seed(3)
    ref, net = oi.GATConvLayer(128, 96, num_heads=4, dropout=0.0), ann.GATConvLayer(128, 96, num_heads=4, dropout=0.0)
    net.load_state_dict(ref.state_dict())
    net.to(DEV)
    x1 = b["x"].clone().requires_grad_(True)
    o1 = ref(x1, ei)
    pg, ig = grads_of(ref, o1, [x1])
    x2 = b["x"].to(DEV).requires_grad_(True)
    check_gat(net, net(x2, ei.to(DEV)), x2, o1, pg, ig[0])
    ref, net = oi.OnsetEmbedding(128, 96), ann.OnsetEmbedding(128, 96)
    net.load_state_dict(ref.state_dict())
    net.to(DEV)
    x1 = b["x"].clone().requires_grad_(True)
    o1 = ref(x1, ei)
    pg, ig = grads_of(ref, o1, [x1])
    x2 = b["x"].to(DEV).requires_grad_(True)
    o2 = net(x2, ei.to(DEV))
    assert_close(o2, o1, FP32_REL, "forward")
    pg2, ig2 = grads_of(net, o2, [x2])
    for k in pg:
        assert_close(pg2[k], pg[k], FP32_REL, f"grad {k}")
    assert_close(ig2[0], ig[0], FP32_REL, "grad x")
