"""Generates the committed golden vectors from the REFERENCE's own code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

* ``edges_*.npz``   -- edge lists produced by the reference's
  ``hetero_graph_from_note_array`` (analysisgnn/utils/hgraph.py:214-300) and its
  ``add_beat_nodes`` / ``add_measure_nodes`` (:41-73) on small note arrays.
* ``decode_*.pt``   -- outputs of the reference's ``onsetwise_logit_aggregation``
  (analysisgnn/models/analysis.py:44-101, torch_scatter shim) on ``synth.decode_case`` inputs.
* ``intree_*.pt``   -- inputs, state_dict, forward output and all gradients of the
  reference's ``SageConvScatter`` / ``HeteroConv`` / ``MetricalConvLayer`` /
  ``MetricalGNN`` (analysisgnn/models/core/{gnn,hgnn}.py) executed through
  ``oracle/ref_loader.py`` (torch_scatter shim), seeds 0-2.

* ``convblocks.pt`` -- the same for the reference's ``GATConvLayer`` and ``OnsetEmbedding`` (gnn.py:154-209, 294-311).

The files are small on purpose; tests compare the oracle restatement and the
CUDA path with them.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from analysisgnn_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402


def hand_score():
    """12 notes, 2 voices, one rest gap in the upper voice (hand-checkable)."""
    rows = [  # onset_div, duration_div, pitch, voice
        (0, 4, 48, 0), (0, 2, 60, 1), (2, 2, 62, 1), (4, 4, 50, 0), (4, 1, 64, 1), (6, 2, 65, 1),
        (8, 4, 52, 0), (8, 4, 67, 1), (12, 2, 53, 0), (12, 4, 69, 1), (14, 2, 55, 0), (16, 4, 57, 0),
    ]
    na = np.zeros(len(rows), dtype=synth.NOTE_DTYPE)
    for i, (o, d, p, v) in enumerate(rows):
        na[i] = (o, d, o / 4, d / 4, 4, p, v)
    return na


def edge_goldens():
    build = ref_loader.load_edge_builder()
    add_beats, add_measures = ref_loader.load_metrical_edge_builders()
    cases = {"hand12": hand_score()}
    for seed, n, voices in ((0, 60, 4), (1, 97, 2), (2, 120, 8), (3, 500, 4)):
        cases[f"synth_s{seed}_n{n}_v{voices}"] = synth.synth_note_array(n, seed, voices)
    for name, na in cases.items():
        out = build(na, pot_edge_dist=0)
        edges = out[1]                       # (nodes, edges) when pot_edge_dist == 0 (hgraph.py:300)
        ns = types.SimpleNamespace(note_array=na, name=name)
        add_beats(ns)
        measures = synth.measure_bounds(na)
        add_measures(ns, measures)
        np.savez_compressed(os.path.join(HERE, f"edges_{name}.npz"), note_array=na,
                            edges=np.asarray(edges, dtype=np.int64),
                            beat_nodes=np.asarray(ns.beat_nodes), beat_edges=np.asarray(ns.beat_edges, dtype=np.int64),
                            measures=measures, measure_nodes=np.asarray(ns.measure_nodes),
                            measure_edges=np.asarray(ns.measure_edges, dtype=np.int64))
        print(name, "edges", np.asarray(edges).shape)


def _grads(module, out, inputs):
    weights = torch.linspace(0.25, 1.25, out.numel(), dtype=out.dtype).view_as(out)
    loss = (out * weights).sum()
    params = [p for p in module.parameters() if p.requires_grad]
    names = [n for n, p in module.named_parameters() if p.requires_grad]
    got = torch.autograd.grad(loss, params + inputs, allow_unused=True)
    pg = {n: g for n, g in zip(names, got[:len(params)]) if g is not None}
    ig = [g for g in got[len(params):]]
    return pg, ig


def intree_goldens():
    gnn, hgnn = ref_loader.load_core()
    for seed in (0, 1, 2):
        torch.manual_seed(seed)
        b = synth.intree_batch(2, 40 + 9 * seed, seed, voices=4, in_features=8, reverse=True, metrical=True)
        x = b["x"].clone().requires_grad_(True)
        ei, et = b["edge_index"], b["edge_type"]
        record = {"batch": {k: v for k, v in b.items() if k != "etypes"}, "etypes": b["etypes"]}
        # SageConvScatter on the 'consecutive' relation, and on an empty edge set
        conv = gnn.SageConvScatter(8, 16)
        pick = et == 1
        out = conv(x, ei[:, pick])
        pg, ig = _grads(conv, out, [x])
        record["sage"] = {"state": conv.state_dict(), "edge_index": ei[:, pick], "out": out.detach(),
                          "param_grads": pg, "x_grad": ig[0]}
        out0 = conv(x, ei[:, :0])
        pg0, ig0 = _grads(conv, out0, [x])
        record["sage_empty"] = {"out": out0.detach(), "param_grads": pg0, "x_grad": ig0[0]}
        # HeteroConv, 7 relations (rests_rev may be empty in small graphs)
        hc = hgnn.HeteroConv(8, 16, b["etypes"], module=gnn.SageConvScatter)
        out = hc(x, ei, et)
        pg, ig = _grads(hc, out, [x])
        record["hetero"] = {"state": hc.state_dict(), "out": out.detach(), "param_grads": pg, "x_grad": ig[0]}
        # MetricalGNN, metrical branch on, ragged beat / measure sequences, eval mode (dropout off, BN running stats)
        for mode in ("train", "eval"):
            torch.manual_seed(seed + 10)
            net = hgnn.MetricalGNN(8, 16, 16, b["etypes"], num_layers=3, dropout=0.0, metrical=True,
                                   conv_block=gnn.SageConvScatter)
            net.train(mode == "train")
            state = {k: v.clone() for k, v in net.state_dict().items()}
            out = net(x, ei, et, b["beat_nodes"], b["measure_nodes"], b["beat_edges"], b["measure_edges"],
                      beat_lengths=b["beat_lengths"], measure_lengths=b["measure_lengths"])
            pg, ig = _grads(net, out, [x])
            record[f"metrical_{mode}"] = {"state": state, "out": out.detach(), "param_grads": pg, "x_grad": ig[0]}
        torch.manual_seed(seed + 20)
        net = hgnn.MetricalGNN(8, 16, 16, b["etypes"], num_layers=2, dropout=0.0, metrical=False,
                               conv_block=gnn.SageConvScatter)
        out = net(x, ei, et, None, None, None, None)
        pg, ig = _grads(net, out, [x])
        record["plain"] = {"state": net.state_dict(), "out": out.detach(), "param_grads": pg, "x_grad": ig[0]}
        torch.save(record, os.path.join(HERE, f"intree_seed{seed}.pt"))
        print("intree seed", seed, "nodes", x.shape[0], "edges", ei.shape[1])


def convblock_goldens():
    """GATConvLayer (gnn.py:154-209) and OnsetEmbedding (:294-311) of the reference, alone, on the 'during' relation."""
    gnn, _ = ref_loader.load_core()
    record = {}
    for seed in (0, 1):
        torch.manual_seed(40 + seed)
        b = synth.intree_batch(2, 50 + 13 * seed, 30 + seed, voices=4, in_features=12, metrical=False)
        ei = b["edge_index"][:, b["edge_type"] == 2]
        x = b["x"].clone().requires_grad_(True)
        gat = gnn.GATConvLayer(12, 20, num_heads=3, dropout=0.0)
        out = gat(x, ei)
        pg, ig = _grads(gat, out, [x])
        record[f"gat{seed}"] = {"x": b["x"], "edge_index": ei, "state": gat.state_dict(), "out": out.detach(),
                                "param_grads": pg, "x_grad": ig[0]}
        gat.train()                                  # attention dropout on: the output must not change (see oracle)
        gat.attndrop.p = 0.5
        record[f"gat{seed}"]["out_train_dropout"] = gat(x, ei).detach()
        for loops in (True, False):
            emb = gnn.OnsetEmbedding(12, 20, add_self_loops=loops)
            out = emb(x, ei)
            pg, ig = _grads(emb, out, [x])
            record[f"onset{seed}_{int(loops)}"] = {"x": b["x"], "edge_index": ei, "state": emb.state_dict(),
                                                   "out": out.detach(), "param_grads": pg, "x_grad": ig[0]}
    torch.save(record, os.path.join(HERE, "convblocks.pt"))
    print("convblocks", sorted(record))


DECODE_CASES = {   # name -> synth.decode_case kwargs
    "single": dict(n_notes=240, seed=0),
    "single_extra_nodes": dict(n_notes=200, seed=1, extra_nodes=40),
    "single_valid_mask": dict(n_notes=220, seed=2, valid_fraction=0.8),
    "single_tpc": dict(n_notes=260, seed=3, with_tpc=True),
    "two_scores": dict(n_notes=180, seed=4, n_scores=2),
    "one_note": dict(n_notes=1, seed=5),
}


def decode_goldens():
    from oracle import decode as odecode
    fn = ref_loader.load_onsetwise_decode()
    for name, kw in DECODE_CASES.items():
        case = synth.decode_case(**kw)
        logits = {k: v.clone() for k, v in case["logits"].items()}
        originals = dict(logits)                                  # the reference also mutates these tensors in place
        graph = odecode.note_store(case["x"], case["batch"], case["onset_div"], case["edge_index_dict"])
        out = fn(logits, graph, batch_size=case["batch_size"], valid_label_mask=case["valid_label_mask"])
        torch.save({"kwargs": kw, "out": {k: v.clone() for k, v in out.items()},
                    "mutated_inputs": {k: v.clone() for k, v in originals.items()}},
                   os.path.join(HERE, f"decode_{name}.pt"))
        print("decode", name, {k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    edge_goldens()
    intree_goldens()
    convblock_goldens()
    decode_goldens()
