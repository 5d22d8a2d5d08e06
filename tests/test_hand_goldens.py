"""Hand goldens for the PyG-shaped operators (goldens #5 and #6): literal numbers derived with
scalar Python from the published operator definitions (tests/golden/hand_goldens.py -- no torch, nothing shared with
oracle/ or analysisgnn_b200/).  They are the pin under oracle/pyg.py, whose third-party originals (torch_geometric,
graphmuse) cannot be installed here: the CPU suite holds oracle/pyg.py to these numbers, the GPU suite holds the CUDA
modules to the same numbers.

Cases: SAGEConv mean with zero-in-degree rows and a duplicate edge; HeteroConv(aggr='sum') with a relation that
misses a target and a node type that receives nothing; an HGTConv target with incoming edges from THREE relations
(joint softmax across relations, two source types), an isolated target (attention output zero) and the per-relation
softmax variant; trim_to_layer over two sampled hops (node / edge counts per layer and the final values)."""
import pytest
import torch

from oracle import pyg as op
from tests.golden import hand_goldens as hg
from tests.util import DEV

TOL = 2e-6      # literals are rounded to 9 decimals; values are O(1); fp32 arithmetic on <= 24-term sums

SAGE = [[-0.5, 0.375, -1.125, -0.25], [-2.0, 0.0, -0.09375, -1.75], [-3.53125, 1.890625, 0.859375, -3.28125],
        [0.8125, -0.1875, -1.875, 1.0625]]
HETERO_SUM = [[-2.0, 0.5, 2.3125, -1.5], [-1.71875, 0.5, 2.3125, -1.21875], [-2.21875, 1.875, 1.0625, -1.71875]]
HGT_JOINT = [
    [-0.575912985, -0.91807727, 0.185163612, 0.146012851, -0.196151434, 0.129015284, 0.867938687, -0.795617939],
    [-0.5, 0.188770334, 0.877540669, -0.533155502, 0.155614833, 0.066311003, -0.566311003, 0.122459331],
    [0.376047567, -1.408432402, 0.468782139, 0.909203069, -0.8752769, 0.223863477, 0.12096623, -0.342121399]]
HGT_PER_RELATION = [
    [-0.018980476, -1.985772177, 0.527910746, 0.70294536, -1.263846341, 0.471762418, 1.424871196, -1.863312845],
    [-0.5, 0.188770334, 0.877540669, -0.533155502, 0.155614833, 0.066311003, -0.566311003, 0.122459331],
    [0.376047567, -1.408432402, 0.468782139, 0.909203069, -0.8752769, 0.223863477, 0.12096623, -0.342121399]]
TRIM_SIZES = [(9, 9), (5, 4), (2, 0)]
TRIM_OUT = [[0.0, 0.525634766, 0.0, 0.130615234, 0.0, 0.09375, 0.380615234, 0.0],
            [0.0, 0.666015625, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]]


def T(rows):
    return torch.tensor(rows, dtype=torch.float32)


def ei(edges):
    return torch.tensor(edges, dtype=torch.long).t().contiguous() if edges else torch.zeros((2, 0), dtype=torch.long)


def close(got, want):
    got = got.detach().float().cpu()
    want = T(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    err = float((got - want).abs().max())
    assert err <= TOL, (err, got, want)


def set_sage(conv, salt):
    with torch.no_grad():
        conv.lin_l.weight.copy_(T(hg.matrix(conv.lin_l.weight.shape[0], conv.lin_l.weight.shape[1], hg.weight, salt)))
        conv.lin_l.bias.copy_(T([hg.bias(o, salt) for o in range(conv.lin_l.bias.shape[0])]))
        conv.lin_r.weight.copy_(T(hg.matrix(conv.lin_r.weight.shape[0], conv.lin_r.weight.shape[1], hg.weight, salt + 1)))


def set_hgt(conv):
    p = hg.hgt_params()
    with torch.no_grad():
        for t in hg.HGT_TYPES:
            conv.kqv_lin[t].weight.copy_(T(p["kqv_w"][t]))
            conv.kqv_lin[t].bias.copy_(T(p["kqv_b"][t]))
            conv.out_lin[t].weight.copy_(T(p["out_w"][t]))
            conv.out_lin[t].bias.copy_(T(p["out_b"][t]))
            conv.skip[t].fill_(hg.HGT_SKIP[t])
        conv.k_rel.copy_(T(p["k_rel"]))
        conv.v_rel.copy_(T(p["v_rel"]))
        for et in hg.HGT_RELS:
            conv.p_rel[op.rel_key(et)].copy_(T([hg.HGT_PREL[et]]))


def test_the_derivation_script_reproduces_the_literals():
    """The literals above are what tests/golden/hand_goldens.py computes (scalar Python, no torch)."""
    r = lambda rows: [[round(v, 9) for v in row] for row in rows]
    assert r(hg.case_sage()) == SAGE and r(hg.case_hetero_sum()) == HETERO_SUM
    assert r(hg.case_hgt(True)) == HGT_JOINT and r(hg.case_hgt(False)) == HGT_PER_RELATION
    sizes, out = hg.case_trim()
    assert [tuple(s) for s in sizes] == TRIM_SIZES and r(out) == TRIM_OUT
    # spot checks a reader can redo on paper: node 0 of the SAGE case has no incoming edge, so its row is
    # b_l + W_r x_0; the isolated HGT target is sigma(skip) * b_out + (1 - sigma(skip)) * x
    x0 = [hg.feat(0, f) for f in range(hg.F_IN)]
    want = [hg.bias(o, 1) + sum(hg.weight(o, f, 2) * x0[f] for f in range(hg.F_IN)) for o in range(hg.F_OUT)]
    assert want == SAGE[0]
    import math
    g = 1.0 / (1.0 + math.exp(-0.5))
    assert [round(g * hg.bias(o, 20) + (1 - g) * hg.feat(1, o, 5), 9) for o in range(hg.HGT_C)] == HGT_JOINT[1]


# ----------------------------------------------------------------------------------------- builders shared by both suites

def run_sage(mod, dev):
    conv = mod.SAGEConv(hg.F_IN, hg.F_OUT)
    set_sage(conv, 1)
    conv.to(dev)
    x = T(hg.matrix(4, hg.F_IN, hg.feat)).to(dev)
    return conv(x, x, ei(hg.SAGE_EDGES).to(dev))


def run_hetero_sum(mod, dev):
    ets = list(hg.HET_EDGES) + [("a", "r9", "b")]             # a relation type that is absent from the batch
    layer = mod.HeteroSAGELayer(ets, hg.F_IN, hg.F_OUT, "sum")
    set_sage(layer.convs["a__r1__a"], 5)
    set_sage(layer.convs["b__r2__a"], 7)
    layer.to(dev)
    x = {"a": T(hg.matrix(3, hg.F_IN, hg.feat, 3)).to(dev), "b": T(hg.matrix(2, hg.F_IN, hg.feat, 4)).to(dev)}
    return layer(x, {et: ei(e).to(dev) for et, e in hg.HET_EDGES.items()})


def run_hgt(mod, dev, joint):
    conv = mod.HGTConv(hg.HGT_C, hg.HGT_C, (hg.HGT_TYPES, hg.HGT_RELS), hg.HGT_H, joint_softmax=joint)
    set_hgt(conv)
    conv.to(dev)
    x = {"a": T(hg.matrix(3, hg.HGT_C, hg.feat, 5)).to(dev), "b": T(hg.matrix(2, hg.HGT_C, hg.feat, 6)).to(dev)}
    return conv(x, {et: ei(e).to(dev) for et, e in hg.HGT_EDGES.items()})


def run_trim(mod, dev):
    et = ("note", "to", "note")
    stack = mod.HeteroSAGEStack([et], hg.TRIM_F, hg.TRIM_F, hg.TRIM_LAYERS)
    for layer, hl in enumerate(stack.convs):
        conv = hl.convs["note__to__note"]
        with torch.no_grad():
            conv.lin_l.weight.copy_(T(hg.matrix(hg.TRIM_F, hg.TRIM_F, hg.weight, 50 + layer)))
            conv.lin_l.bias.copy_(T([hg.bias(o, 50 + layer) for o in range(hg.TRIM_F)]))
            conv.lin_r.weight.copy_(T(hg.matrix(hg.TRIM_F, hg.TRIM_F, hg.weight, 60 + layer)))
    stack.to(dev)
    x = {"note": T(hg.matrix(9, hg.TRIM_F, hg.feat, 9)).to(dev)}
    collect = []
    out = stack(x, {et: ei(hg.TRIM_EDGES).to(dev)}, {"note": hg.TRIM_NODES_PER_HOP}, {et: hg.TRIM_EDGES_PER_HOP},
                collect)
    return out, collect


def check_all(mod, dev):
    close(run_sage(mod, dev), SAGE)
    out = run_hetero_sum(mod, dev)
    assert set(out) == {"a"}                                  # type b receives nothing: absent (PyG HeteroConv)
    close(out["a"], HETERO_SUM)
    for joint, want in ((True, HGT_JOINT), (False, HGT_PER_RELATION)):
        out = run_hgt(mod, dev, joint)
        assert set(out) == {"a"}
        close(out["a"], want)
    out, collect = run_trim(mod, dev)
    assert [c["note"].shape[0] for c in collect] == [n for n, _ in TRIM_SIZES]      # 9 -> 5 -> 2 nodes
    close(out["note"], TRIM_OUT)


def test_oracle_pyg_matches_the_hand_goldens():
    check_all(op, "cpu")


def test_oracle_trim_to_layer_counts():
    x = {"note": torch.zeros(9, 1)}
    e = {("note", "to", "note"): ei(hg.TRIM_EDGES)}
    sizes = []
    for layer in range(3):
        x, e = op.trim_to_layer(layer, {"note": hg.TRIM_NODES_PER_HOP}, {("note", "to", "note"): hg.TRIM_EDGES_PER_HOP},
                                x, e)
        sizes.append((x["note"].shape[0], e[("note", "to", "note")].shape[1]))
    assert sizes == TRIM_SIZES


@pytest.mark.gpu
@pytest.mark.parametrize("operands", ["f16", "tf32"])
def test_cuda_modules_match_the_hand_goldens(operands):
    from analysisgnn_b200 import linalg
    from analysisgnn_b200 import nn as ann

    class Mods:
        SAGEConv, HeteroSAGELayer, HGTConv, HeteroSAGEStack = (ann.SAGEConv, ann.HeteroSAGELayer, ann.HGTConv,
                                                                ann.HeteroSAGEStack)
    old = linalg.parity_operands()
    linalg.set_parity_operands(operands)
    try:
        check_all(Mods, DEV)
    finally:
        linalg.set_parity_operands(old)
