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
