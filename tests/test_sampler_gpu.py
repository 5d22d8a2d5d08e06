"""GPU samplers vs the CPU definition (oracle/graph.py): bit-exact for a given seed."""
import numpy as np
import pytest
import torch

from analysisgnn_b200 import sampler, scoregraph, synth
from oracle import graph as og
from tests.util import DEV

pytestmark = pytest.mark.gpu


def _corpus(sizes, seed=0, voices=4):
    arrays = [synth.synth_note_array(n, seed * 100 + k, voices) for k, n in enumerate(sizes)]
    edges, ptr = scoregraph.score_graph_edges(arrays, DEV)
    x = torch.randn(int(ptr[-1]), 8, device=DEV)
    return arrays, sampler.Corpus(x, edges, ptr.cpu().tolist())


def test_counter_rng_matches_the_oracle():
    for args in [(0, 0, 0, 0, 0), (7, 1, 2, 3, 4), (2 ** 63 + 5, 9, 9, 9, 9)]:
        assert sampler.rng_u64(*args) == og.rng_u64(*args)
    assert sampler.window_start(3, 5, 900, 100) == og.window_start(3, 5, 900, 100)


@pytest.mark.parametrize("size", [100, 500, 10_000])
def test_window_subgraphs_match_the_recipe(size):
    sizes = [700, 120, 1500, 64, 501]
    arrays, corpus = _corpus(sizes, seed=1)
    ids = [2, 0, 4, 1, 3]
    starts = [sampler.window_start(9, g, sizes[g], size) for g in ids]
    edges, eid, node_index, slot_ptr = sampler.window_subgraphs(corpus.edges[0], corpus.edges[1], corpus.edges[2],
                                                                corpus.node_ptr, corpus.edge_ptr, ids, starts, size)
    want_e, want_nodes, off = [], [], 0
    for g, st in zip(ids, starts):
        e = og.score_graph_edges(arrays[g])
        ei, et, _ = og.window_subgraph(e[:2], e[2], sizes[g], st, size)
        w = min(size, sizes[g] - st)
        want_e.append(np.concatenate((ei + off, et[None]), axis=0))
        want_nodes.append(np.arange(corpus.node_ptr[g] + st, corpus.node_ptr[g] + st + w))
        off += w
    np.testing.assert_array_equal(edges.cpu().numpy(), np.concatenate(want_e, axis=1))
    np.testing.assert_array_equal(node_index.cpu().numpy(), np.concatenate(want_nodes))
    assert slot_ptr[-1] == off
    # edge ids point at the corpus edges the batch edges came from
    src_back = corpus.edges[0][eid] - node_index[edges[0]]
    assert int(src_back.abs().max()) == 0


@pytest.mark.parametrize("fanouts", [[3, 3], [5], [-1], [2, 4, 1], [32, 32]])
def test_neighbor_sample_matches_the_oracle(fanouts):
    arrays, corpus = _corpus([400], seed=4, voices=8)
    n = 400
    e = corpus.edges.cpu().numpy()
    rowptr, col, perm = og.csr_build(e[1], e[0], n, e[2], 4)
    seeds = np.arange(100, 140)
    want = og.neighbor_sample(rowptr, col, n, seeds, fanouts, seed=9, n_rel=4)
    got = sampler.neighbor_sample(corpus.csr_by_destination(), n, torch.as_tensor(seeds, device=DEV), fanouts, seed=9)
    np.testing.assert_array_equal(got["node"].cpu().numpy(), want["node"])
    assert got["num_sampled_nodes"] == want["num_sampled_nodes"]
    for k in range(4):
        assert got["num_sampled_edges"][k] == want["num_sampled_edges"][k]
        for key in ("src", "dst", "edge"):
            np.testing.assert_array_equal(got[key][k].cpu().numpy(), want[key][k], err_msg=f"{key}[{k}]")


def test_neighbor_sample_depends_on_the_seed_only():
    arrays, corpus = _corpus([300, 300], seed=5, voices=8)
    seeds = torch.arange(50, 150, device=DEV)
    a = sampler.neighbor_sample(corpus.csr_by_destination(), 600, seeds, [2, 2], seed=1)
    b = sampler.neighbor_sample(corpus.csr_by_destination(), 600, seeds, [2, 2], seed=1)
    c = sampler.neighbor_sample(corpus.csr_by_destination(), 600, seeds, [2, 2], seed=2)
    assert torch.equal(a["node"], b["node"]) and all(torch.equal(x, y) for x, y in zip(a["edge"], b["edge"]))
    assert any(x.shape != y.shape or not torch.equal(x, y) for x, y in zip(a["edge"], c["edge"]))


def test_loader_batches_feed_the_encoder_and_shard_across_ranks():
    from analysisgnn_b200 import nn as ann
    sizes = [600, 800, 520, 700, 900, 610, 530, 750]
    arrays, corpus = _corpus(sizes, seed=6)
    plain = sampler.ScoreGraphLoader(corpus, subgraph_size=500, batch_size=4, seed=3)
    b = plain.batch(0, 0)
    assert b["batch_size"] == 2000 and b["x_dict"]["note"].shape == (2000, 8)
    assert b["batch_dict"]["note"].bincount().tolist() == [500] * 4
    # data-parallel sharding: ranks take disjoint subgraphs of the same global batch
    r0 = sampler.ScoreGraphLoader(corpus, 500, 4, seed=3, rank=0, world_size=2).batch(0, 0)
    r1 = sampler.ScoreGraphLoader(corpus, 500, 4, seed=3, rank=1, world_size=2).batch(0, 0)
    assert sorted(r0["graph_ids"] + r1["graph_ids"]) == sorted(b["graph_ids"])
    assert torch.equal(torch.cat((r0["node_index"], r1["node_index"])).sort()[0], b["node_index"].sort()[0])
    # sampled hops: PyG layout (targets first, per-hop counts) straight into the trimmed SAGE stack
    hop = sampler.ScoreGraphLoader(corpus, subgraph_size=200, batch_size=3, num_neighbors=[4, 4], seed=3).batch(1, 1)
    n_all = hop["x_dict"]["note"].shape[0]
    assert sum(hop["num_sampled_nodes_dict"]["note"]) == n_all and hop["batch_size"] == 600
    stack = ann.HeteroSAGEStack(list(hop["edge_index_dict"].keys()), 8, 16, 3).to(DEV)
    out = stack(hop["x_dict"], hop["edge_index_dict"], hop["num_sampled_nodes_dict"], hop["num_sampled_edges_dict"])
    assert out["note"].shape[0] == hop["batch_size"] + 0 * n_all or out["note"].shape[0] <= n_all
    assert torch.isfinite(out["note"]).all()


def test_corpus_file_round_trip_feeds_the_same_batches(tmp_path):
    """corpusfile: device corpus -> file -> pinned staging -> device; the loader's batches (plain windows and sampled
    hops) from the reloaded corpus are bit-identical to those from the original."""
    from analysisgnn_b200 import corpusfile
    sizes = [300, 420, 260, 510, 333]
    arrays, corpus = _corpus(sizes, seed=8)
    corpus.extras["onset_div"] = torch.cat([torch.as_tensor(a["onset_div"].astype(np.int64)) for a in arrays]).to(DEV)
    path = str(tmp_path / "corpus.agc")
    corpusfile.save_corpus(path, corpus)
    loaded = corpusfile.load_corpus(path, device=DEV)
    assert loaded.x.is_cuda and loaded.edges.dtype == torch.int64 and loaded.node_ptr == corpus.node_ptr
    assert torch.equal(loaded.x, corpus.x) and torch.equal(loaded.edges, corpus.edges)
    for kwargs in (dict(), dict(num_neighbors=[3, 2])):
        a = sampler.ScoreGraphLoader(corpus, subgraph_size=200, batch_size=3, seed=5, **kwargs).batch(1, 0)
        b = sampler.ScoreGraphLoader(loaded, subgraph_size=200, batch_size=3, seed=5, **kwargs).batch(1, 0)
        assert a["graph_ids"] == b["graph_ids"] and a["batch_size"] == b["batch_size"]
        assert torch.equal(a["node_index"], b["node_index"])
        assert torch.equal(a["x_dict"]["note"], b["x_dict"]["note"])
        assert torch.equal(a["extras"]["onset_div"], b["extras"]["onset_div"])
        for et in a["edge_index_dict"]:
            assert torch.equal(a["edge_index_dict"][et], b["edge_index_dict"][et]), et


def test_loader_iterates_hetero_batches():
    """``for batch in loader`` yields what the reference's training step reads (sampler.HeteroBatch); a new ``iter()``
    is a new epoch with another order."""
    sizes = [300, 420, 260, 510, 333]
    arrays, corpus = _corpus(sizes, seed=9)
    corpus.extras["pitch_spelling"] = torch.randint(0, 35, (corpus.node_ptr[-1],), device=DEV)
    loader = sampler.ScoreGraphLoader(corpus, subgraph_size=200, batch_size=2, seed=4)
    first = list(loader)
    assert len(first) == len(loader) == 3
    seen = []
    for b in first:
        n = b["note"].batch_size
        assert n == 200 * len(b.graph_ids) and b.x_dict["note"].shape == (n, 8)
        assert b["note"].pitch_spelling.shape == (n,) and "pitch_spelling" in b["note"].keys()
        assert torch.equal(b["note"].pitch_spelling, corpus.extras["pitch_spelling"][b.node_index])
        assert b.batch_dict["note"].bincount().tolist() == [200] * len(b.graph_ids)
        assert b.num_sampled_nodes_dict is None and set(b.edge_types) == set(b.edge_index_dict)
        seen += b.graph_ids
    assert sorted(seen) == list(range(5))
    second = [b.graph_ids for b in loader]
    assert sorted(g for ids in second for g in ids) == list(range(5))
    assert second == [loader.batch_ids(1, i) for i in range(3)] and [b.graph_ids for b in first] == \
        [loader.batch_ids(0, i) for i in range(3)]


# ------------------------------------------------------------------------------------------ static-shape batches

def _static_setup(n_scores=12, batch=4, size=200, seed=3, ratio=None):
    c = synth.corpus(n_scores, lambda g: size + 40 + 37 * (g % 5), seed, in_features=8)
    corpus = sampler.Corpus(c["x"].to(DEV), c["edges"].to(DEV), c["node_ptr"], extras={k: v.to(DEV) for k, v in c["extras"].items()})
    sb = sampler.StaticBatcher(corpus, size, batch, beat_of=c["beat_of"].to(DEV), measure_of=c["measure_of"].to(DEV))
    loader = sampler.ScoreGraphLoader(corpus, size, batch, seed=5, subgraph_sample_ratio=ratio)
    return c, corpus, sb, loader


def _reference_batch(c, ids, starts, size):
    """The same batch built on the host, score by score, the way synth.hetero_batch collates (compact, no padding)."""
    e = c["edges"].numpy()
    ptr = c["node_ptr"]
    out = {"nn": [], "beat": [], "measure": []}
    off = {"note": 0, "beat": 0, "measure": 0}
    counts = {"beat": [], "measure": []}
    for g, st in zip(ids, starts):
        lo, hi = ptr[g] + st, ptr[g] + st + size
        sel = (e[0] >= lo) & (e[0] < hi) & (e[1] >= lo) & (e[1] < hi)
        # corpus edges are stored score by score: the selection keeps corpus order
        out["nn"].append(np.stack((e[0][sel] - lo + off["note"], e[1][sel] - lo + off["note"], e[2][sel])))
        for name, of in (("beat", c["beat_of"].numpy()), ("measure", c["measure_of"].numpy())):
            ids_v = of[lo:hi]
            first, n_v = ids_v[0], ids_v[-1] - ids_v[0] + 1
            out[name].append(np.stack((np.arange(size) + off["note"], ids_v - first + off[name])))
            counts[name].append(n_v)
            off[name] += n_v
        off["note"] += size
    return {k: np.concatenate(v, axis=1) for k, v in out.items()}, counts


@pytest.mark.parametrize("n_scores,ratio", [(12, None), (2, 1.0)])
def test_static_batches_match_the_host_collation(n_scores, ratio):
    """StaticBatcher: fixed shapes, -1 padding; after dropping the padding the edges are exactly (values AND order) the
    induced window subgraphs collated score by score, beat / measure nodes are the windows' contiguous id ranges.
    Second case: ``subgraph_sample_ratio`` with more visits than batches -- both scores appear twice in the one batch
    of the epoch, each visit with its own window (slots are independent: nothing is shared between them)."""
    size, batch = 200, 4
    c, corpus, sb, loader = _static_setup(n_scores=n_scores, batch=batch, size=size, ratio=ratio)
    shapes = None
    for index in range(len(loader)):
        sel_host = sb.select(loader, 0, index).clone()
        out = sb.batch(sel_host.to(DEV))
        ids, starts = sel_host[0].tolist(), sel_host[1].tolist()
        assert ids == loader.batch_ids(0, index) and starts == loader.window_starts(0, index)
        if ratio is not None:
            assert sorted(ids) == [0, 0, 1, 1] and len(set(zip(ids, starts))) == 4
        want, counts = _reference_batch(c, ids, starts, size)
        eid = out["edge_index_dict"]
        grp = eid.groups[0]
        ei, et = grp.edge_index.cpu().numpy(), grp.edge_type.cpu().numpy()
        cap = sb.edge_cap
        fwd = et[:cap] >= 0
        np.testing.assert_array_equal(np.stack((ei[0, :cap][fwd], ei[1, :cap][fwd], et[:cap][fwd])), want["nn"])
        assert int(out["n_edges"]) == want["nn"].shape[1] and (ei[:, :cap][:, ~fwd] == -1).all()
        # reversed copies of types 1..3 as types 4..6
        rev = et[cap:] >= 0
        keep = want["nn"][2] > 0
        np.testing.assert_array_equal(np.stack((ei[0, cap:][rev], ei[1, cap:][rev], et[cap:][rev])),
                                      np.stack((want["nn"][1][keep], want["nn"][0][keep], want["nn"][2][keep] + 3)))
        for name in ("beat", "measure"):
            up = eid[("note", "connects", name)].cpu().numpy()
            np.testing.assert_array_equal(up, want[name])
            np.testing.assert_array_equal(eid[(name, "connects_rev", "note")].cpu().numpy(), want[name][::-1])
            nxt = eid[(name, "next", name)].cpu().numpy()
            valid = nxt[0] >= 0
            ends = np.cumsum(counts[name])
            want_next = np.array([k for k in range(ends[-1]) if k + 1 not in ends and k + 1 < ends[-1]])
            np.testing.assert_array_equal(nxt[0][valid], want_next)
            np.testing.assert_array_equal(nxt[1][valid], want_next + 1)
            assert out["x_dict"][name].shape[0] == sb.virtual[name][1] >= ends[-1]
        # per-type views of the typed group: what generic consumers (onset pooling) index
        onset = eid[("note", "onset", "note")].cpu().numpy()
        np.testing.assert_array_equal(onset[:, onset[0] >= 0], want["nn"][:2][:, want["nn"][2] == 0])
        # features / labels follow the windows
        nodes = np.concatenate([np.arange(c["node_ptr"][g] + st, c["node_ptr"][g] + st + size) for g, st in zip(ids, starts)])
        np.testing.assert_array_equal(out["node_index"].cpu().numpy(), nodes)
        assert torch.equal(out["x_dict"]["note"].cpu(), c["x"][nodes])
        assert torch.equal(out["extras"]["cadence"].cpu(), c["extras"]["cadence"][nodes])
        sig = {k: tuple(v.shape) for k, v in out["x_dict"].items()}, grp.edge_index.shape
        assert shapes is None or sig == shapes                 # the same shapes for every batch
        shapes = sig


def test_static_batch_gives_the_same_encoder_output_as_the_compact_batch():
    """Padding is invisible to the model: the encoder shell on a static batch (typed COO with -1 slots, isolated padding
    beats / measures) returns the note logits of the same batch in compact PyG form."""
    from analysisgnn_b200 import nn as ann
    size, batch = 200, 4
    c, corpus, sb, loader = _static_setup(batch=batch, size=size)
    sel = sb.select(loader, 1, 0).clone()
    out = sb.batch(sel.to(DEV))
    want, counts = _reference_batch(c, sel[0].tolist(), sel[1].tolist(), size)
    tasks = c["tasks"]
    torch.manual_seed(0)
    net = ann.AnalysisEncoder(sb.metadata, 8, 32, 16, tasks, 2, dropout=0.0).to(DEV)
    run = lambda x_dict, ei, bd: net(out["extras"]["pitch_spelling"], out["extras"]["key_signature"], x_dict, ei, bd,
                                     out["batch_size"], None, None)
    got = run(out["x_dict"], out["edge_index_dict"], out["batch_dict"])
    names = ["onset", "consecutive", "during", "rest"]
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=DEV)
    ei = {("note", nm, "note"): t(want["nn"][:2][:, want["nn"][2] == k]) for k, nm in enumerate(names)}
    for k, nm in enumerate(names[1:], 1):
        ei[("note", nm + "_rev", "note")] = ei[("note", nm, "note")].flip(0)
    x_dict = {"note": out["x_dict"]["note"]}
    bd = {"note": out["batch_dict"]["note"]}
    for name in ("beat", "measure"):
        n_v = int(sum(counts[name]))
        ei[("note", "connects", name)] = t(want[name])
        ei[(name, "connects_rev", "note")] = t(want[name][::-1])
        ends = np.cumsum(counts[name])
        nx = np.array([k for k in range(n_v) if k + 1 not in ends and k + 1 < n_v])
        ei[(name, "next", name)] = t(np.stack((nx, nx + 1)))
        x_dict[name] = out["x_dict"][name][:n_v]
        bd[name] = out["batch_dict"][name][:n_v]
    ref = run(x_dict, ei, bd)
    for task in tasks:
        assert torch.allclose(got[task], ref[task], rtol=0, atol=2e-6 * float(ref[task].abs().max())), task
