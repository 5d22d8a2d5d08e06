"""agnn_csr_build vs the CPU oracle (oracle/graph.py::csr_build): bit-exact."""
import numpy as np
import pytest
import torch

from analysisgnn_b200 import graph, synth
from oracle import graph as og

from tests.util import DEV  # noqa: E402
pytestmark = pytest.mark.gpu


def _build(row, col, n_rows, n_cols, et, n_rel):
    t = lambda a: None if a is None else torch.as_tensor(np.asarray(a), dtype=torch.long, device=DEV)
    seg = graph.Segment(t(row), t(col), n_rows, n_cols, t(et), n_rel)
    out = graph.build_csr([seg], device=torch.device(DEV))[0]
    torch.cuda.synchronize()
    return out


def _check(row, col, n_rows, n_cols, et, n_rel):
    got = _build(row, col, n_rows, n_cols, et, n_rel)
    rowptr, c, perm = og.csr_build(row, col, n_rows, et, n_rel)
    np.testing.assert_array_equal(got.rowptr.cpu().numpy().reshape(-1), rowptr)
    kept = len(perm)
    np.testing.assert_array_equal(got.col.cpu().numpy()[:kept], c)
    np.testing.assert_array_equal(got.perm.cpu().numpy()[:kept], perm)


@pytest.mark.parametrize("n,e,r,seed", [(1, 0, 1, 0), (7, 0, 3, 0), (50, 400, 3, 1), (1000, 20000, 7, 2),
                                         (4097, 50000, 1, 3), (33, 5000, 2, 4), (100000, 300000, 4, 5)])
def test_random_graphs(n, e, r, seed):
    rng = np.random.default_rng(seed)
    row, col = rng.integers(0, n, e), rng.integers(0, n, e)
    et = rng.integers(0, r, e) if r > 1 else None
    _check(row, col, n, n, et, r)


def test_dropped_relation_codes():
    rng = np.random.default_rng(7)
    n, e, r = 200, 5000, 3
    row, col = rng.integers(0, n, e), rng.integers(0, n, e)
    et = rng.integers(-2, r + 2, e)
    _check(row, col, n, n, et, r)


def test_long_rows_take_the_block_sort_path():
    """Rows above 32 entries go through the bitonic block sort (shared memory up to 8192
    entries, in place beyond); all must still come out in input order."""
    rng = np.random.default_rng(8)
    n = 64
    row = np.concatenate([np.full(20000, 3), np.full(700, 5), np.full(33, 9), rng.integers(0, n, 3000)])
    rng.shuffle(row)
    col = rng.integers(0, n, len(row))
    _check(row, col, n, n, None, 1)


def test_rectangular_and_out_of_range():
    rng = np.random.default_rng(9)
    n_rows, n_cols, e = 30, 500, 2000
    row, col = rng.integers(0, n_rows, e), rng.integers(0, n_cols, e)
    _check(row, col, n_rows, n_cols, None, 1)
    bad = row.copy()
    bad[17] = n_rows + 3
    seg = graph.Segment(torch.as_tensor(bad, device=DEV), torch.as_tensor(col, device=DEV), n_rows, n_cols)
    with pytest.raises(ValueError):
        graph.build_csr([seg], validate=True)


def test_many_segments_in_one_call():
    """More than AGNN_MAX_SEG segments (chunked), mixed sizes, one output arena."""
    rng = np.random.default_rng(10)
    segs, want = [], []
    for k in range(40):
        n, e, r = int(rng.integers(1, 300)), int(rng.integers(0, 2000)), int(rng.integers(1, 4))
        row, col, et = rng.integers(0, n, e), rng.integers(0, n, e), rng.integers(0, r, e)
        segs.append(graph.Segment(torch.as_tensor(row, device=DEV), torch.as_tensor(col, device=DEV), n, n,
                                  torch.as_tensor(et, device=DEV), r))
        want.append(og.csr_build(row, col, n, et, r))
    got = graph.build_csr(segs)
    torch.cuda.synchronize()
    for g, (rowptr, c, perm) in zip(got, want):
        base = int(g.rowptr.reshape(-1)[0])
        np.testing.assert_array_equal(g.rowptr.cpu().numpy().reshape(-1) - base, rowptr)
        np.testing.assert_array_equal(g.col.cpu().numpy(), c)
        np.testing.assert_array_equal(g.perm.cpu().numpy(), perm)


def test_score_graph_batch_typed_and_transposed():
    b = synth.intree_batch(6, 200, 3, in_features=8)
    ei, et = b["edge_index"].to(DEV), b["edge_type"].to(DEV)
    n = b["x"].shape[0]
    csr = graph.TypedCSR(ei, et, n, n_rel=7)
    torch.cuda.synchronize()
    e = b["edge_index"].numpy()
    for side, (r_, c_) in ((csr.fwd, (e[0], e[1])), (csr.bwd, (e[1], e[0]))):
        rowptr, c, perm = og.csr_build(r_, c_, n, b["edge_type"].numpy(), 7)
        np.testing.assert_array_equal(side.rowptr.cpu().numpy().reshape(-1) - int(side.rowptr.reshape(-1)[0]), rowptr)
        np.testing.assert_array_equal(side.col.cpu().numpy(), c)
        np.testing.assert_array_equal(side.perm.cpu().numpy(), perm)


def test_full_size_properties():
    """BASELINE config 1 size (100 x 500 notes, 7 relations): sortedness, permutation, stability."""
    b = synth.intree_batch(100, 500, 0, in_features=8, metrical=False)
    ei, et = b["edge_index"].to(DEV), b["edge_type"].to(DEV)
    n = b["x"].shape[0]
    csr = graph.TypedCSR(ei, et, n, n_rel=7).fwd
    rowptr = csr.rowptr.cpu().numpy().astype(np.int64)
    perm = csr.perm.cpu().numpy().astype(np.int64)
    col = csr.col.cpu().numpy()
    e = b["edge_index"].numpy()
    t = b["edge_type"].numpy()
    assert np.array_equal(np.sort(perm), np.arange(e.shape[1]))
    assert np.array_equal(col, e[1][perm])
    key = t[perm] * n + e[0][perm]
    assert np.all(np.diff(key) >= 0)
    assert np.all(np.diff(perm)[np.diff(key) == 0] > 0)
    counts = np.bincount(t * n + e[0], minlength=7 * n).reshape(7, n)
    assert np.array_equal(np.diff(rowptr, axis=1), counts)
