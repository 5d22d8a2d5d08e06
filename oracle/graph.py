"""CPU oracle for the integer side of the hot path: score-graph edges, CSR build,
window / neighbour subgraph sampling and batch collation.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  numpy, bit-exact.

* ``score_graph_edges``   restates analysisgnn/utils/hgraph.py:214-300 (pinned
  against the reference body via ``ref_loader.load_edge_builder``).
* ``beat_edges`` / ``measure_edges`` restate hgraph.py:61-73 / :41-59.
* ``csr_build``           = the "stable sort by reduce-side row" every PyG-style
  sampler and this repo's kernels assume.
* ``window_subgraph``     restates analysisgnn/data/datasets/chord.py:217-229
  (contiguous window, isin filter, subtract start).
* ``neighbor_sample``     pyg-lib ``hetero_neighbor_sample`` semantics with this
  repo's counter-based RNG (pyg-lib's stream cannot be reproduced: "same seed"
  parity is defined against THIS function).
"""
from __future__ import annotations

import numpy as np

MASK64 = (1 << 64) - 1


# ---------------------------------------------------------------- score graph

def score_graph_edges(note_array) -> np.ndarray:
    """Edge list int64 [3,E] (src, dst, type); types onset=0, consecutive=1,
    during=2, rest=3.  Follows analysisgnn/utils/hgraph.py:232-285 for the case
    rest_array=None, pot_edge_dist=0 -- one pass per note, emission order kept."""
    onset = np.asarray(note_array["onset_div"])
    dur = np.asarray(note_array["duration_div"])
    n = len(onset)
    rows = []
    for i in range(n):
        o, e = onset[i], onset[i] + dur[i]
        for j in np.flatnonzero(onset == o):          # hgraph.py:233-237
            if j != i:
                rows.append((i, j, 0))
        for j in np.flatnonzero(onset == e):          # hgraph.py:244-247
            rows.append((i, j, 1))
        for j in np.flatnonzero((o < onset) & (e > onset)):   # hgraph.py:255-259
            rows.append((i, j, 2))
    end = onset + dur                                   # hgraph.py:273-285
    for et in np.sort(np.unique(end))[:-1]:
        if et in onset:
            continue
        srcs = np.flatnonzero(end == et)
        gap = onset - et
        gap = np.where(gap > 0, gap, np.inf)
        dsts = np.flatnonzero(gap == gap.min())
        for i in srcs:
            for j in dsts:
                rows.append((i, j, 3))
    if not rows:
        return np.zeros((3, 0), dtype=np.int64)
    return np.asarray(rows, dtype=np.int64).T.copy()


def beat_edges(note_array):
    """(n_beats, [2,E]) per hgraph.py:61-73 (``int(max)`` beats: the last partial
    beat has no node)."""
    ob = np.asarray(note_array["onset_beat"])
    n_beats = int(ob.max())
    cols = []
    for b in range(n_beats):
        idx = np.flatnonzero((ob >= b) & (ob < b + 1))
        if idx.size:
            cols.append(np.vstack((idx, np.full(idx.size, b))))
    e = np.hstack(cols).astype(np.int64) if cols else np.zeros((2, 0), dtype=np.int64)
    return n_beats, e


def measure_edges(note_array, measures):
    """(n_measures, [2,E]) per hgraph.py:41-59."""
    onset = np.asarray(note_array["onset_div"])
    cols = []
    for m in range(len(measures)):
        idx = np.flatnonzero((onset >= measures[m, 0]) & (onset < measures[m, 1]))
        if idx.size:
            cols.append(np.vstack((idx, np.full(idx.size, m))))
    e = np.hstack(cols).astype(np.int64) if cols else np.zeros((2, 0), dtype=np.int64)
    return len(measures), e


# ------------------------------------------------------------------------ CSR

def csr_build(row, col, n_rows, etype=None, n_rel=1):
    """Relation-major CSR of a COO edge list, reduce side = ``row``.

    Returns (rowptr int32 [n_rel*(n_rows+1)], col int32 [E], perm int32 [E]):
    edges are ordered by (relation, row) with ties kept in input order (stable);
    ``rowptr[r*(n_rows+1)+i]`` indexes the concatenated ``col``/``perm`` arrays;
    ``perm[k]`` is the input position of the k-th CSR entry.  Edges whose type is
    outside [0, n_rel) are dropped (rowptr's last entry is the kept count).
    """
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    et = np.zeros(len(row), dtype=np.int64) if etype is None else np.asarray(etype, dtype=np.int64)
    keep = np.flatnonzero((et >= 0) & (et < n_rel))
    key = et[keep] * n_rows + row[keep]
    order = np.argsort(key, kind="stable")
    perm = keep[order]
    counts = np.bincount(key, minlength=n_rel * n_rows).reshape(n_rel, n_rows)
    rowptr = np.zeros((n_rel, n_rows + 1), dtype=np.int64)
    base = 0
    for r in range(n_rel):
        rowptr[r, 0] = base
        rowptr[r, 1:] = base + np.cumsum(counts[r])
        base = rowptr[r, -1]
    return rowptr.reshape(-1).astype(np.int32), col[perm].astype(np.int32), perm.astype(np.int32)


# -------------------------------------------------------------------- samplers

def mix64(x: int) -> int:
    """splitmix64 finaliser."""
    x &= MASK64
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & MASK64
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & MASK64
    x ^= x >> 31
    return x


def rng_u64(seed: int, a: int, b: int, c: int, d: int) -> int:
    """Counter-based RNG: one 64-bit draw keyed on (seed, a, b, c, d)."""
    h = mix64(seed + 0x9E3779B97F4A7C15)
    for v in (a, b, c, d):
        h = mix64(h ^ ((v + 0x9E3779B97F4A7C15) & MASK64))
    return h


def window_start(seed: int, graph_id: int, n_nodes: int, size: int, draw: int = 0) -> int:
    """``start = randint(0, n - size)`` (chord.py:219) with the counter RNG; ``draw`` = which of several windows of
    the same score in one batch (0 for the first)."""
    if n_nodes <= size:
        return 0
    return rng_u64(seed, 0x57494E, graph_id, draw, 0) % (n_nodes - size + 1)


def window_subgraph(edge_index, edge_type, n_nodes, start, size):
    """Node-induced subgraph of the contiguous window [start, start+size):
    keep edges with both ends inside, in input order, re-indexed by ``- start``
    (analysisgnn/data/datasets/chord.py:217-229, utils/hgraph.py:404-452)."""
    edge_index = np.asarray(edge_index, dtype=np.int64)
    stop = min(start + size, n_nodes)
    keep = ((edge_index[0] >= start) & (edge_index[0] < stop) &
            (edge_index[1] >= start) & (edge_index[1] < stop))
    ids = np.flatnonzero(keep)
    return edge_index[:, ids] - start, np.asarray(edge_type)[ids], ids


def sample_row(seed, hop, rel, dst_global, neigh_pos, fanout):
    """Choose ``fanout`` of the ``len(neigh_pos)`` CSR slots of a row without
    replacement; all of them (in CSR order) if deg <= fanout.

    Floyd-free formulation chosen so one GPU thread can do it with O(fanout)
    state: partial Fisher-Yates over a virtual identity array, recording swaps
    in a tiny open list; draw t uses ``rng_u64(seed, hop, rel, dst, t)``.
    Chosen slots are returned in draw order."""
    deg = len(neigh_pos)
    if fanout < 0 or deg <= fanout:
        return list(neigh_pos)
    swapped_idx, swapped_val = [], []

    def get(i):
        for a, v in zip(swapped_idx, swapped_val):
            if a == i:
                return v
        return i

    def put(i, v):
        for k, a in enumerate(swapped_idx):
            if a == i:
                swapped_val[k] = v
                return
        swapped_idx.append(i)
        swapped_val.append(v)

    out = []
    for t in range(fanout):
        j = t + rng_u64(seed, hop, rel, dst_global, t) % (deg - t)
        vt, vj = get(t), get(j)
        put(j, vt)
        put(t, vj)
        out.append(neigh_pos[vj])
    return out


def neighbor_sample(rowptr, col, n_nodes, seeds, fanouts, seed, n_rel=1):
    """k-hop uniform neighbour sampling on a relation-major CSR (reduce side =
    destination; ``col`` holds sources), single node type.

    Per hop, for relation r = 0..R-1, for each frontier node in order: take its
    in-neighbours (all if deg <= fanout, else ``sample_row``); a source not seen
    before is appended to the node list in discovery order.  Emits local
    (src, dst) pairs per relation, global ``node`` ids, CSR slot ids ``edge``,
    and ``num_sampled_nodes`` / ``num_sampled_edges`` per hop (PyG
    ``num_sampled_*`` layout: entry 0 of nodes = #seeds).
    """
    rowptr = np.asarray(rowptr, dtype=np.int64).reshape(n_rel, n_nodes + 1)
    col = np.asarray(col, dtype=np.int64)
    local = -np.ones(n_nodes, dtype=np.int64)
    nodes = []
    for s in seeds:
        if local[s] < 0:
            local[s] = len(nodes)
            nodes.append(int(s))
    n_per_hop = [len(nodes)]
    e_per_hop = [[] for _ in range(n_rel)]
    src_l = [[] for _ in range(n_rel)]
    dst_l = [[] for _ in range(n_rel)]
    eid = [[] for _ in range(n_rel)]
    lo = 0
    for hop, k in enumerate(fanouts):
        hi = len(nodes)
        for r in range(n_rel):
            before = len(eid[r])
            for li in range(lo, hi):
                g = nodes[li]
                slots = list(range(rowptr[r, g], rowptr[r, g + 1]))
                for p in sample_row(seed, hop, r, g, slots, k):
                    s = int(col[p])
                    if local[s] < 0:
                        local[s] = len(nodes)
                        nodes.append(s)
                    src_l[r].append(local[s])
                    dst_l[r].append(li)
                    eid[r].append(p)
            e_per_hop[r].append(len(eid[r]) - before)
        n_per_hop.append(len(nodes) - hi)
        lo = hi
    return {
        "node": np.asarray(nodes, dtype=np.int64),
        "src": [np.asarray(v, dtype=np.int64) for v in src_l],
        "dst": [np.asarray(v, dtype=np.int64) for v in dst_l],
        "edge": [np.asarray(v, dtype=np.int64) for v in eid],
        "num_sampled_nodes": n_per_hop,
        "num_sampled_edges": e_per_hop,
    }
