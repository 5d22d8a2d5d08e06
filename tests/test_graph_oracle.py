"""Integer side of the oracle (oracle/graph.py) and the host-side synthetic graph
generator (analysisgnn_b200/synth.py), pinned to edge lists produced by the
reference's own builder (tests/golden/edges_*.npz; analysisgnn/utils/hgraph.py:41-73,
214-300).  Everything here is bit-exact."""
import numpy as np
import pytest

from analysisgnn_b200 import synth
from oracle import graph as og
from oracle import ref_loader
from tests.util import EDGE_CASES, golden_edges


@pytest.mark.parametrize("name", EDGE_CASES)
def test_score_graph_edges_match_reference(name):
    g = golden_edges(name)
    na = g["note_array"]
    want = g["edges"]
    np.testing.assert_array_equal(og.score_graph_edges(na), want)        # oracle restatement
    np.testing.assert_array_equal(synth.score_graph_edges(na), want)     # vectorised host generator


@pytest.mark.parametrize("name", EDGE_CASES)
def test_beat_measure_edges_match_reference(name):
    g = golden_edges(name)
    na = g["note_array"]
    nb, be = og.beat_edges(na)
    assert nb == len(g["beat_nodes"])
    np.testing.assert_array_equal(be, g["beat_edges"])
    nb2, be2 = synth.beat_edges(na, reference_quirk=True)
    assert nb2 == nb
    np.testing.assert_array_equal(be2, g["beat_edges"])
    nm, me = og.measure_edges(na, g["measures"])
    assert nm == len(g["measure_nodes"])
    np.testing.assert_array_equal(me, g["measure_edges"])
    nm2, me2 = synth.measure_edges(na, g["measures"])
    np.testing.assert_array_equal(me2, g["measure_edges"])


def test_hand_score_edges_by_hand():
    """The 12-note score of make_golden.hand_score, checked by hand: note 0 (onset 0, dur 4)
    shares its onset with note 1, is followed by notes 3 and 4 (onset 4) and sounds during
    note 2 (onset 2)."""
    e = golden_edges("hand12")["edges"]
    from0 = {(int(d), int(t)) for s, d, t in e.T if s == 0}
    assert from0 == {(1, 0), (3, 1), (4, 1), (2, 2)}
    # note 4 (onset 4, dur 1) ends at 5 where nothing starts: a rest edge to the next onset (note 5 at 6)
    rest = {(int(s), int(d)) for s, d, t in e.T if t == 3}
    assert (4, 5) in rest


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("seed,n,voices", [(5, 200, 4), (6, 333, 3), (7, 64, 8)])
def test_edges_match_live_reference(seed, n, voices):
    build = ref_loader.load_edge_builder()
    na = synth.synth_note_array(n, seed, voices)
    want = np.asarray(build(na, pot_edge_dist=0)[1], dtype=np.int64)
    np.testing.assert_array_equal(synth.score_graph_edges(na), want)
    np.testing.assert_array_equal(og.score_graph_edges(na), want)


def test_survey_edge_statistics():
    """~4.6 forward edges per note at 4 voices."""
    na = synth.synth_note_array(500, 0, 4)
    e = synth.score_graph_edges(na)
    assert 4.0 < e.shape[1] / 500 < 5.2
    assert set(np.unique(e[2])) <= {0, 1, 2, 3}


def test_csr_build_oracle_is_a_stable_sort():
    rng = np.random.default_rng(0)
    n, e, r = 50, 400, 3
    row, col = rng.integers(0, n, e), rng.integers(0, n, e)
    et = rng.integers(-1, r + 1, e)          # includes out-of-range types, which are dropped
    rowptr, c, perm = og.csr_build(row, col, n, et, r)
    rp = rowptr.reshape(r, n + 1)
    kept = np.flatnonzero((et >= 0) & (et < r))
    assert rp[-1, -1] == len(kept) == len(perm)
    assert sorted(perm.tolist()) == kept.tolist()
    np.testing.assert_array_equal(c, col[perm])
    for k in range(r):
        assert rp[k, 0] == (rp[k - 1, -1] if k else 0)
        for i in range(n):
            seg = perm[rp[k, i]:rp[k, i + 1]]
            assert np.all(et[seg] == k) and np.all(row[seg] == i)
            assert np.all(np.diff(seg) > 0)          # input order kept inside a row


def test_csr_build_oracle_empty():
    rowptr, c, perm = og.csr_build(np.zeros(0, np.int64), np.zeros(0, np.int64), 4, None, 2)
    assert rowptr.tolist() == [0] * 10 and len(c) == 0 and len(perm) == 0


def test_window_subgraph_follows_reference_recipe():
    """analysisgnn/data/datasets/chord.py:217-229: arange window, isin filter, subtract start."""
    na = synth.synth_note_array(300, 3, 4)
    e = synth.score_graph_edges(na)
    start, size = 57, 100
    ei, et, ids = og.window_subgraph(e[:2], e[2], 300, start, size)
    nodes = np.arange(start, start + size)
    keep = np.isin(e[0], nodes) & np.isin(e[1], nodes)
    np.testing.assert_array_equal(ei, e[:2, keep] - start)
    np.testing.assert_array_equal(et, e[2, keep])
    np.testing.assert_array_equal(ids, np.flatnonzero(keep))
    assert og.window_start(1, 0, 80, 100) == 0
    s = og.window_start(1, 7, 300, 100)
    assert 0 <= s <= 200 and s == og.window_start(1, 7, 300, 100)


def test_neighbor_sample_invariants():
    na = synth.synth_note_array(400, 4, 8)
    e = synth.score_graph_edges(na)
    n, r = 400, 4
    rowptr, col, perm = og.csr_build(e[1], e[0], n, e[2], r)      # reduce side = destination
    seeds = np.arange(100, 140)
    out = og.neighbor_sample(rowptr, col, n, seeds, [3, 3], seed=9, n_rel=r)
    again = og.neighbor_sample(rowptr, col, n, seeds, [3, 3], seed=9, n_rel=r)
    other = og.neighbor_sample(rowptr, col, n, seeds, [3, 3], seed=10, n_rel=r)
    np.testing.assert_array_equal(out["node"], again["node"])
    assert any(len(a) != len(b) or np.any(a != b) for a, b in zip(out["edge"], other["edge"]))
    np.testing.assert_array_equal(out["node"][:40], seeds)
    assert len(set(out["node"].tolist())) == len(out["node"])
    assert sum(out["num_sampled_nodes"]) == len(out["node"])
    rp = rowptr.reshape(r, n + 1)
    for k in range(r):
        assert sum(out["num_sampled_edges"][k]) == len(out["edge"][k])
        src_g, dst_g = out["node"][out["src"][k]], out["node"][out["dst"][k]]
        np.testing.assert_array_equal(col[out["edge"][k]], src_g)          # sampled slots are real edges
        for p, d in zip(out["edge"][k], dst_g):
            assert rp[k, d] <= p < rp[k, d + 1]
        assert len(set(out["edge"][k].tolist())) == len(out["edge"][k])    # no slot taken twice
        # fan-out respected per hop: hop h's edges are the h-th block of the relation's list
        start = 0
        for cnt_h in out["num_sampled_edges"][k]:
            dsts = out["dst"][k][start:start + cnt_h]
            if cnt_h:
                assert np.unique(dsts, return_counts=True)[1].max() <= 3
            start += cnt_h
    full = og.neighbor_sample(rowptr, col, n, seeds, [-1], seed=9, n_rel=r)
    for k in range(r):
        deg = sum(rp[k, s + 1] - rp[k, s] for s in seeds)
        assert len(full["edge"][k]) == deg
