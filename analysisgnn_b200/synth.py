"""Seeded synthetic music-score graphs (host-side data generation, numpy only).

There is no network for the reference's corpora, so every test and benchmark
runs on synthetic 4/4 textures whose edge structure follows the reference's
own score-graph definition (``analysisgnn/utils/hgraph.py:214-300``: onset=0,
consecutive=1, during=2, rest=3) and whose sizes follow SURVEY.md §8(d).

The edge construction here is a vectorised O(N log N) formulation (sorted
onsets + ``searchsorted`` ranges); it emits edges in exactly the order the
reference's per-note loop does, which ``tests/test_graph_oracle.py`` checks
against the oracle restatement and the golden edge lists.
"""
from __future__ import annotations

import numpy as np

NOTE_EDGE_TYPES = ("onset", "consecutive", "during", "rest")
REV_EDGE_TYPES = ("consecutive_rev", "during_rev", "rest_rev")
# in-tree naming of the 7 relations (analysisgnn/models/chord.py:513-514)
INTREE_ETYPES = {"onset": 0, "consecutive": 1, "during": 2, "rests": 3,
                 "consecutive_rev": 4, "during_rev": 5, "rests_rev": 6}

DIVS_PER_BEAT = 4
BEATS_PER_MEASURE = 4

NOTE_DTYPE = np.dtype([
    ("onset_div", "<i4"), ("duration_div", "<i4"), ("onset_beat", "<f4"),
    ("duration_beat", "<f4"), ("ts_beats", "<i4"), ("pitch", "<i4"), ("voice", "<i4"),
])


def synth_note_array(n_notes: int, seed: int, voices: int = 4) -> np.ndarray:
    """A structured note array of ``n_notes`` notes sorted by (onset_div, pitch).

    Per voice a monophonic stream: durations from {1,2,2,4,4,8} divs, a rest gap
    from {1,2,4} divs with p=0.05 after a note, pitch = 36+12*v+U[0,12).
    """
    rng = np.random.default_rng(seed)
    dur_choices = np.array([1, 2, 2, 4, 4, 8], dtype=np.int64)
    gap_choices = np.array([1, 2, 4], dtype=np.int64)
    per_voice = n_notes  # over-generate, then keep the earliest n_notes
    onsets, durs, pitches, vids = [], [], [], []
    for v in range(voices):
        d = dur_choices[rng.integers(0, len(dur_choices), per_voice)]
        gap = np.where(rng.random(per_voice) < 0.05,
                       gap_choices[rng.integers(0, len(gap_choices), per_voice)], 0)
        step = d + gap
        on = np.concatenate(([0], np.cumsum(step)[:-1]))
        onsets.append(on)
        durs.append(d)
        pitches.append(36 + 12 * v + rng.integers(0, 12, per_voice))
        vids.append(np.full(per_voice, v))
    onset = np.concatenate(onsets)
    dur = np.concatenate(durs)
    pitch = np.concatenate(pitches)
    voice = np.concatenate(vids)
    order = np.lexsort((pitch, onset))[:n_notes]
    onset, dur = onset[order], dur[order]
    # notes sounding past the last onset are all held to the same final end time
    # (a closing chord); otherwise the reference builder's "no later onset" branch
    # links them to every note of the score (see score_graph_edges).
    end = onset + dur
    dur = np.where(end > onset[-1], end.max() - onset, dur)
    out = np.zeros(n_notes, dtype=NOTE_DTYPE)
    out["onset_div"] = onset
    out["duration_div"] = dur
    out["onset_beat"] = onset / DIVS_PER_BEAT
    out["duration_beat"] = dur / DIVS_PER_BEAT
    out["ts_beats"] = BEATS_PER_MEASURE
    out["pitch"] = pitch[order]
    out["voice"] = voice[order]
    return out


def _ranges_to_pairs(lo, hi):
    """All (i, j) with lo[i] <= j < hi[i], i-major then j ascending."""
    cnt = np.maximum(hi - lo, 0)
    total = int(cnt.sum())
    i = np.repeat(np.arange(len(lo), dtype=np.int64), cnt)
    start = np.cumsum(cnt) - cnt
    j = np.arange(total, dtype=np.int64) - np.repeat(start, cnt) + np.repeat(lo, cnt)
    return i, j


def score_graph_edges(note_array: np.ndarray) -> np.ndarray:
    """int64 [3, E] = (src, dst, type) in the reference's emission order.

    Requires ``note_array`` sorted by ``onset_div`` (the synthetic arrays are).
    Semantics: analysisgnn/utils/hgraph.py:232-285 (no rest_array, pot_edge_dist=0).
    """
    onset = np.asarray(note_array["onset_div"], dtype=np.int64)
    dur = np.asarray(note_array["duration_div"], dtype=np.int64)
    n = len(onset)
    if n == 0:
        return np.zeros((3, 0), dtype=np.int64)
    if np.any(np.diff(onset) < 0):
        raise ValueError("note_array must be sorted by onset_div")
    end = onset + dur
    same_lo = np.searchsorted(onset, onset, "left")
    same_hi = np.searchsorted(onset, onset, "right")
    i0, j0 = _ranges_to_pairs(same_lo, same_hi)
    keep = i0 != j0
    i0, j0 = i0[keep], j0[keep]
    i1, j1 = _ranges_to_pairs(np.searchsorted(onset, end, "left"), np.searchsorted(onset, end, "right"))
    # during: onset_i < onset_j < end_i  (dur==0 notes produce nothing)
    i2, j2 = _ranges_to_pairs(same_hi, np.searchsorted(onset, end, "left"))
    src = np.concatenate((i0, i1, i2))
    dst = np.concatenate((j0, j1, j2))
    typ = np.concatenate((np.zeros_like(i0), np.ones_like(i1), np.full_like(i2, 2)))
    order = np.lexsort((dst, typ, src))  # per source note: type 0, then 1, then 2
    src, dst, typ = src[order], dst[order], typ[order]
    # rest edges: end times (except the last) at which no note starts
    ends = np.unique(end)[:-1]
    ends = ends[~np.isin(ends, onset)]
    if len(ends):
        nxt = np.searchsorted(onset, ends, "right")          # first note with onset > et
        late = nxt >= n
        nxt_hi = np.searchsorted(onset, onset[np.minimum(nxt, n - 1)], "right")
        # reference quirk (hgraph.py:277-279): with no later onset every gap is inf and
        # ``tmp == tmp.min()`` selects ALL notes as destinations
        nxt = np.where(late, 0, nxt)
        nxt_hi = np.where(late, n, nxt_hi)
        end_order = np.argsort(end, kind="stable")
        end_sorted = end[end_order]
        s_lo = np.searchsorted(end_sorted, ends, "left")
        s_hi = np.searchsorted(end_sorted, ends, "right")
        rs, rd = [], []
        for a, b, c, d in zip(s_lo, s_hi, nxt, nxt_hi):
            scr = np.sort(end_order[a:b])
            dstn = np.arange(c, d, dtype=np.int64)
            rs.append(np.repeat(scr, len(dstn)))
            rd.append(np.tile(dstn, len(scr)))
        if rs:
            rs, rd = np.concatenate(rs), np.concatenate(rd)
            src = np.concatenate((src, rs))
            dst = np.concatenate((dst, rd))
            typ = np.concatenate((typ, np.full_like(rs, 3)))
    return np.stack((src, dst, typ)).astype(np.int64)


def beat_edges(note_array: np.ndarray, reference_quirk: bool = True):
    """(n_beats, int64 [2, E]) note->beat edges, ``floor(onset_beat) == b``.

    ``reference_quirk=True`` mirrors analysisgnn/utils/hgraph.py:61-73, where the
    beat nodes are ``arange(int(onset_beat.max()))`` so notes in the last
    (partial) beat get no beat node.  ``False`` adds that last beat.
    """
    ob = np.asarray(note_array["onset_beat"], dtype=np.float64)
    n_beats = int(ob.max()) if reference_quirk else int(np.floor(ob.max())) + 1
    b = np.floor(ob).astype(np.int64)
    idx = np.nonzero(b < n_beats)[0]
    order = np.lexsort((idx, b[idx]))
    idx = idx[order]
    return n_beats, np.stack((idx, b[idx])).astype(np.int64)


def measure_bounds(note_array: np.ndarray) -> np.ndarray:
    """[[start_div, end_div)] rows covering the score in 4/4."""
    span = DIVS_PER_BEAT * BEATS_PER_MEASURE
    last = int(np.asarray(note_array["onset_div"]).max())
    n_meas = last // span + 1
    starts = np.arange(n_meas, dtype=np.int64) * span
    return np.stack((starts, starts + span), axis=1)


def measure_edges(note_array: np.ndarray, measures: np.ndarray):
    """(n_measures, int64 [2, E]) note->measure edges, start <= onset_div < end
    (analysisgnn/utils/hgraph.py:41-59)."""
    onset = np.asarray(note_array["onset_div"], dtype=np.int64)
    src, dst = [], []
    for m, (s, e) in enumerate(np.asarray(measures)):
        idx = np.nonzero((onset >= s) & (onset < e))[0]
        src.append(idx)
        dst.append(np.full(idx.size, m, dtype=np.int64))
    if src:
        e = np.stack((np.concatenate(src), np.concatenate(dst))).astype(np.int64)
    else:
        e = np.zeros((2, 0), dtype=np.int64)
    return len(measures), e


# --------------------------------------------------------------------------
# Batches
# --------------------------------------------------------------------------

def intree_batch(n_graphs: int, notes_per_graph: int, seed: int, voices: int = 4,
                 in_features: int = 64, reverse: bool = True, metrical: bool = True):
    """A batched homogeneous-index score graph in the layout the reference's in-tree
    ``MetricalGNN.forward`` takes (analysisgnn/models/core/hgnn.py:373): one node
    index space, ``edge_index`` [2,E] + ``edge_type`` [E]; beat/measure nodes with
    note->beat / note->measure edges and cumulative ``*_lengths`` pointers.

    Reverse relations (types 4-6) are flipped copies of types 1-3 appended after
    the forward edges (SURVEY.md §8a quirk list: built here, not by the
    reference's ``add_reverse_edges_from_edge_index``).
    """
    import torch

    rng = np.random.default_rng(seed + 7919)
    eis, ets, bes, mes = [], [], [], []
    n_off = b_off = m_off = 0
    b_ptr, m_ptr = [0], [0]
    for g in range(n_graphs):
        na = synth_note_array(notes_per_graph, seed * 100003 + g, voices)
        e = score_graph_edges(na)
        if reverse:
            rev = e[:, e[2] > 0]
            e = np.concatenate((e, np.stack((rev[1], rev[0], rev[2] + 3))), axis=1)
        eis.append(e[:2] + n_off)
        ets.append(e[2])
        if metrical:
            nb, be = beat_edges(na, reference_quirk=True)
            nm, me = measure_edges(na, measure_bounds(na))
            bes.append(be + np.array([[n_off], [b_off]]))
            mes.append(me + np.array([[n_off], [m_off]]))
            b_off += nb
            m_off += nm
            b_ptr.append(b_off)
            m_ptr.append(m_off)
        n_off += len(na)
    out = {
        "x": torch.from_numpy(rng.standard_normal((n_off, in_features), dtype=np.float32)),
        "edge_index": torch.from_numpy(np.concatenate(eis, axis=1)),
        "edge_type": torch.from_numpy(np.concatenate(ets)),
        "etypes": dict(INTREE_ETYPES) if reverse else {k: v for k, v in INTREE_ETYPES.items() if v < 4},
    }
    if metrical:
        out.update(
            beat_nodes=torch.arange(b_off), measure_nodes=torch.arange(m_off),
            beat_edges=torch.from_numpy(np.concatenate(bes, axis=1)),
            measure_edges=torch.from_numpy(np.concatenate(mes, axis=1)),
            beat_lengths=torch.tensor(b_ptr, dtype=torch.long),
            measure_lengths=torch.tensor(m_ptr, dtype=torch.long),
        )
    return out


def hetero_metadata(add_beats: bool = True, add_measures: bool = True, reverse: bool = True):
    """PyG-style ``(node_types, edge_types)`` for a score graph."""
    node_types = ["note"]
    rels = list(NOTE_EDGE_TYPES) + (list(REV_EDGE_TYPES) if reverse else [])
    edge_types = [("note", r, "note") for r in rels]
    if add_beats:
        node_types.append("beat")
        edge_types += [("note", "connects", "beat"), ("beat", "connects_rev", "note"), ("beat", "next", "beat")]
    if add_measures:
        node_types.append("measure")
        edge_types += [("note", "connects", "measure"), ("measure", "connects_rev", "note"),
                       ("measure", "next", "measure")]
    return node_types, edge_types


def hetero_batch(n_graphs: int, notes_per_graph: int, seed: int, voices: int = 4,
                 in_features: int = 25, add_beats: bool = True, add_measures: bool = True,
                 reverse: bool = True, task_dict=None):
    """A collated PyG-style hetero batch of score subgraphs as plain dicts:
    ``x_dict, edge_index_dict`` (row 0 = source, row 1 = target), ``batch_dict``,
    ``batch_size``, ``pitch_spelling``, ``key_signature`` and per-task labels --
    the fields ``TorchAnalysisGNN.encode`` consumes (analysisgnn/models/analysis.py:571-591).
    Every note is a target (no extra sampled hops), as in BASELINE.json config 1/2.
    """
    import torch

    rng = np.random.default_rng(seed + 104729)
    node_types, edge_types = hetero_metadata(add_beats, add_measures, reverse)
    ei = {et: [] for et in edge_types}
    batch = {nt: [] for nt in node_types}
    off = {nt: 0 for nt in node_types}
    for g in range(n_graphs):
        na = synth_note_array(notes_per_graph, seed * 100003 + g, voices)
        e = score_graph_edges(na)
        n = len(na)
        for t, name in enumerate(NOTE_EDGE_TYPES):
            sel = e[:2, e[2] == t]
            ei[("note", name, "note")].append(sel + off["note"])
            if reverse and t > 0:
                ei[("note", name + "_rev", "note")].append(sel[::-1] + off["note"])
        counts = {"note": n}
        if add_beats:
            nb, be = beat_edges(na, reference_quirk=False)
            shift = np.array([[off["note"]], [off["beat"]]])
            ei[("note", "connects", "beat")].append(be + shift)
            ei[("beat", "connects_rev", "note")].append((be + shift)[::-1])
            nxt = np.stack((np.arange(nb - 1), np.arange(1, nb))).astype(np.int64) + off["beat"]
            ei[("beat", "next", "beat")].append(nxt)
            counts["beat"] = nb
        if add_measures:
            nm, me = measure_edges(na, measure_bounds(na))
            shift = np.array([[off["note"]], [off["measure"]]])
            ei[("note", "connects", "measure")].append(me + shift)
            ei[("measure", "connects_rev", "note")].append((me + shift)[::-1])
            nxt = np.stack((np.arange(nm - 1), np.arange(1, nm))).astype(np.int64) + off["measure"]
            ei[("measure", "next", "measure")].append(nxt)
            counts["measure"] = nm
        for nt, c in counts.items():
            batch[nt].append(np.full(c, g, dtype=np.int64))
            off[nt] += c
    n_note = off["note"]
    x_dict = {"note": torch.from_numpy(rng.standard_normal((n_note, in_features), dtype=np.float32))}
    for nt in node_types[1:]:
        x_dict[nt] = torch.zeros((off[nt], in_features), dtype=torch.float32)
    out = {
        "metadata": (node_types, edge_types),
        "x_dict": x_dict,
        "edge_index_dict": {et: torch.from_numpy(np.ascontiguousarray(np.concatenate(v, axis=1)))
                            for et, v in ei.items()},
        "batch_dict": {nt: torch.from_numpy(np.concatenate(v)) for nt, v in batch.items()},
        "batch_size": n_note,
        "pitch_spelling": torch.from_numpy(rng.integers(0, 35, n_note)),
        "key_signature": torch.from_numpy(rng.integers(0, 15, n_note)),
    }
    task_dict = task_dict if task_dict is not None else {"cadence": 4, "localkey": 50, "romanNumeral": 185}
    out["labels"] = {t: torch.from_numpy(rng.integers(0, c, n_note)) for t, c in task_dict.items()}
    return out


def corpus(n_scores: int, notes_per_score, seed: int, voices: int = 4, in_features: int = 25, task_dict=None):
    """A synthetic corpus in the layout ``sampler.Corpus`` holds on the device: per-note features, typed note -> note
    edges with GLOBAL node ids (the four forward types of analysisgnn/utils/hgraph.py:214-300), per-score node ranges,
    per-note beat / measure ids (local to the score: ``floor(onset_beat)``, 4/4 measures -- hgraph.py:41-73) and the
    per-note arrays the training step reads (spellings, key signatures, labels).  ``notes_per_score``: int or a
    callable ``score -> int``."""
    import torch

    rng = np.random.default_rng(seed + 15485863)
    task_dict = task_dict if task_dict is not None else {"cadence": 4, "localkey": 50, "romanNumeral": 185}
    edges, beat_of, meas_of, node_ptr, n_beats, n_meas = [], [], [], [0], [], []
    for g in range(n_scores):
        n = notes_per_score(g) if callable(notes_per_score) else int(notes_per_score)
        na = synth_note_array(n, seed * 100003 + g, voices)
        e = score_graph_edges(na)
        edges.append(np.stack((e[0] + node_ptr[-1], e[1] + node_ptr[-1], e[2])))
        nb, be = beat_edges(na, reference_quirk=False)
        nm, me = measure_edges(na, measure_bounds(na))
        b = np.zeros(n, dtype=np.int64)
        b[be[0]] = be[1]
        m = np.zeros(n, dtype=np.int64)
        m[me[0]] = me[1]
        beat_of.append(b)
        meas_of.append(m)
        n_beats.append(nb)
        n_meas.append(nm)
        node_ptr.append(node_ptr[-1] + n)
    total = node_ptr[-1]
    extras = {"pitch_spelling": torch.from_numpy(rng.integers(0, 35, total)),
              "key_signature": torch.from_numpy(rng.integers(0, 15, total))}
    for t, c in task_dict.items():
        extras[t] = torch.from_numpy(rng.integers(0, c, total))
    return {"x": torch.from_numpy(rng.standard_normal((total, in_features), dtype=np.float32)),
            "edges": torch.from_numpy(np.ascontiguousarray(np.concatenate(edges, axis=1))),
            "node_ptr": node_ptr, "beat_of": torch.from_numpy(np.concatenate(beat_of)),
            "measure_of": torch.from_numpy(np.concatenate(meas_of)), "n_beats": n_beats, "n_measures": n_meas,
            "extras": extras, "tasks": dict(task_dict)}


DECODE_TASKS = {"quality": 15, "inversion": 4, "degree1": 22, "degree2": 22, "localkey": 50}


def decode_case(n_notes: int, seed: int, n_scores: int = 1, extra_nodes: int = 0, with_tpc: bool = False,
                valid_fraction: float = 1.0, smooth: int = 6):
    """Inputs of the reference's ``onsetwise_logit_aggregation`` (analysisgnn/models/analysis.py:44-101) for a
    synthetic prediction: softmaxed logits per task, notes sorted by onset with chords (equal onsets), the
    ``("note", "onset", "note")`` edges between chord members (both directions, plus some that touch the
    ``extra_nodes`` sampled beyond ``batch_size``), graph ids and an optional valid-label mask.  ``smooth``
    makes neighbouring onsets share their arg-max so that change points are sparse, as in real predictions."""
    import torch

    g = torch.Generator().manual_seed(seed)
    n_total = n_notes + extra_nodes
    steps = torch.randint(0, 3, (n_total,), generator=g)          # 0 = same onset as the previous note (chord)
    steps[0] = 0
    onset = torch.cumsum(steps, 0).to(torch.int64) + 7
    batch = torch.sort(torch.randint(0, n_scores, (n_total,), generator=g)).values if n_scores > 1 else \
        torch.zeros(n_total, dtype=torch.int64)
    src, dst = [], []
    start = 0
    onset_list = onset.tolist()
    for i in range(1, n_total + 1):
        if i == n_total or onset_list[i] != onset_list[start]:
            for a in range(start, i):
                for b in range(start, i):
                    if a != b:
                        src.append(a)
                        dst.append(b)
            start = i
    loops = torch.randint(0, n_total, (max(n_total // 50, 1),), generator=g)    # self loops the reference drops
    e = torch.tensor([src + loops.tolist(), dst + loops.tolist()], dtype=torch.int64).reshape(2, -1)
    e = e[:, torch.randperm(e.shape[1], generator=g)]
    logits = {}
    for k, c in DECODE_TASKS.items():
        base = torch.randn(n_total // smooth + 2, c, generator=g) * 3.0
        rows = base[torch.arange(n_total) // smooth] + 0.5 * torch.randn(n_total, c, generator=g)
        logits[k] = torch.softmax(rows[:n_notes], dim=-1)
    if with_tpc:
        logits["tpc_in_label"] = torch.softmax(torch.randn(n_notes, 2, generator=g) + torch.tensor([0.0, 1.5]), dim=-1)
    valid = None
    if valid_fraction < 1.0:
        valid = torch.rand(n_notes, generator=g) < valid_fraction
        valid[0] = True
    return dict(logits=logits, onset_div=onset, batch=batch, x=torch.zeros(n_total, 1),
                edge_index_dict={("note", "onset", "note"): e}, batch_size=n_notes, valid_label_mask=valid)
