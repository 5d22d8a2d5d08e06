"""GPU score-graph builder vs the edge lists the REFERENCE's builder produced (tests/golden/edges_*.npz,
analysisgnn/utils/hgraph.py:214-300) and vs the oracle restatement: bit-exact, order included."""
import numpy as np
import pytest

from analysisgnn_b200 import scoregraph, synth
from oracle import graph as og
from tests.util import DEV, EDGE_CASES, golden_edges

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", EDGE_CASES)
def test_matches_reference_golden(name):
    g = golden_edges(name)
    edges, _ = scoregraph.score_graph_edges(g["note_array"], DEV)
    np.testing.assert_array_equal(edges.cpu().numpy(), g["edges"])


def test_reference_signature():
    g = golden_edges("hand12")
    nodes, edges = scoregraph.hetero_graph_from_note_array(g["note_array"], pot_edge_dist=0, device=DEV)
    assert len(nodes) == 12 and edges.dtype == np.int64
    np.testing.assert_array_equal(edges, g["edges"])


@pytest.mark.parametrize("seed,n,voices", [(11, 1, 1), (12, 2, 2), (13, 777, 3), (14, 3000, 6), (15, 500, 1)])
def test_matches_oracle_on_random_scores(seed, n, voices):
    na = synth.synth_note_array(n, seed, voices)
    edges, _ = scoregraph.score_graph_edges(na, DEV)
    np.testing.assert_array_equal(edges.cpu().numpy(), og.score_graph_edges(na))


def test_quirks_zero_duration_and_no_later_onset():
    """dur == 0 notes (a consecutive edge to themselves) and the reference's 'no later onset'
    branch that links a rest source to EVERY note (hgraph.py:277-279)."""
    na = np.zeros(7, dtype=synth.NOTE_DTYPE)
    na["onset_div"] = [0, 0, 2, 4, 4, 6, 6]
    na["duration_div"] = [2, 0, 1, 2, 3, 1, 4]
    edges, _ = scoregraph.score_graph_edges(na, DEV)
    np.testing.assert_array_equal(edges.cpu().numpy(), og.score_graph_edges(na))
    assert (edges[2] == 3).sum() > 0


def test_batch_of_scores_is_the_collated_concatenation():
    arrays = [synth.synth_note_array(n, 20 + k, v) for k, (n, v) in enumerate([(120, 4), (1, 1), (333, 2), (64, 8)])]
    edges, ptr = scoregraph.score_graph_edges(arrays, DEV)
    want, off = [], 0
    for na in arrays:
        e = og.score_graph_edges(na)
        want.append(e + np.array([[off], [off], [0]]))
        off += len(na)
    np.testing.assert_array_equal(edges.cpu().numpy(), np.concatenate(want, axis=1))
    assert ptr.cpu().tolist() == [0, 120, 121, 454, 518]


def test_full_size_properties():
    """BASELINE config 5 size: one 200 000-note score.  Equal to the vectorised host generator (itself
    pinned to the reference), sorted by source within the note block, types in range."""
    na = synth.synth_note_array(200_000, 5, 4)
    edges, _ = scoregraph.score_graph_edges(na, DEV)
    e = edges.cpu().numpy()
    np.testing.assert_array_equal(e, synth.score_graph_edges(na))
    assert 4.0 < e.shape[1] / 200_000 < 5.2


def test_unsorted_input_is_rejected():
    na = synth.synth_note_array(10, 0, 2)
    na["onset_div"][3] = 1000
    with pytest.raises(ValueError):
        scoregraph.score_graph_edges(na, DEV)
