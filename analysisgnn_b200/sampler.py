"""Subgraph sampling and batch collation on the GPU (agnn_window_subgraph, agnn_sample_hop_*).

Replaces the CPU loaders of the reference (analysisgnn/data/datamodules/analysis.py:270-323: graphmuse
``MuseNeighborLoader`` -> PyG ``NeighborSampler`` -> pyg-lib, in DataLoader worker processes) for
corpora that live on the device: a ``Corpus`` holds every score's notes and typed edges (global node
ids, as ``scoregraph.score_graph_edges`` builds them); a batch is ``batch_size`` scores, one contiguous
``subgraph_size``-note window each (analysisgnn/data/datasets/chord.py:217-229), optionally grown by
``num_neighbors`` sampled hops, collated with PyG's target-first node order and per-hop counts.

Random choices are functions of ``(seed, ...)`` only (splitmix64 counter RNG, oracle/graph.py), so a
batch is reproducible bit for bit on any device count.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib, graph

MASK64 = (1 << 64) - 1


def _mix64(x: int) -> int:
    x &= MASK64
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & MASK64
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & MASK64
    x ^= x >> 31
    return x


def rng_u64(seed: int, a: int, b: int, c: int, d: int) -> int:
    """The counter RNG shared with the kernels: one 64-bit draw keyed on (seed, a, b, c, d)."""
    h = _mix64(seed + 0x9E3779B97F4A7C15)
    for v in (a, b, c, d):
        h = _mix64(h ^ ((v + 0x9E3779B97F4A7C15) & MASK64))
    return h


def window_start(seed: int, graph_id: int, n_nodes: int, size: int, draw: int = 0) -> int:
    """``start = randint(0, n - size)`` (chord.py:219) with the counter RNG.  ``draw`` tells apart several windows of
    one score in the same batch (``subgraph_sample_ratio``); 0 for the first."""
    if n_nodes <= size:
        return 0
    return rng_u64(seed, 0x57494E, graph_id, draw, 0) % (n_nodes - size + 1)


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def window_subgraphs(src, dst, etype, node_ptr: Sequence[int], edge_ptr: Sequence[int], graph_ids: Sequence[int],
                     starts: Sequence[int], size: int):
    """Node-induced subgraphs of contiguous windows of several scores, collated.

    ``src/dst/etype``: corpus COO with global node ids (device int64); ``node_ptr/edge_ptr`` host lists.
    Returns ``(edges int64 [3, E], edge_id int64 [E], node_index int64 [N_batch], slot_ptr list)``:
    ``node_index`` = corpus node of every batch node, ``slot_ptr`` = first batch node of every slot."""
    dev = src.device
    if dev.type != "cuda":
        raise _lib.AgnnError("analysisgnn_b200 has no CPU path: the sampler needs CUDA tensors")
    lib = _lib.lib()
    cand_ptr, edge_lo, node_lo, win, out_off = [0], [], [], [], []
    off = 0
    for g, st in zip(graph_ids, starts):
        n = node_ptr[g + 1] - node_ptr[g]
        w = min(size, n - st)
        edge_lo.append(edge_ptr[g])
        cand_ptr.append(cand_ptr[-1] + edge_ptr[g + 1] - edge_ptr[g])
        node_lo.append(node_ptr[g] + st)
        win.append(w)
        out_off.append(off)
        off += w
    n_cand = cand_ptr[-1]
    t64 = lambda a: torch.tensor(a, dtype=torch.int64, device=dev)
    d_cand, d_elo, d_nlo, d_off = t64(cand_ptr), t64(edge_lo), t64(node_lo), t64(out_off)
    d_win = torch.tensor(win, dtype=torch.int32, device=dev)
    ws_bytes = lib.agnn_window_workspace(n_cand)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    n_out = torch.zeros(1, dtype=torch.int32, device=dev)
    capacity = max(n_cand, 1)
    edges = torch.empty((3, capacity), dtype=torch.int64, device=dev)
    eid = torch.empty(capacity, dtype=torch.int64, device=dev)
    _lib.check(lib.agnn_window_subgraph(len(win), n_cand, d_cand.data_ptr(), d_elo.data_ptr(), d_nlo.data_ptr(),
                                        d_win.data_ptr(), d_off.data_ptr(), src.data_ptr(), dst.data_ptr(),
                                        etype.data_ptr() if etype is not None else None, edges.data_ptr(),
                                        eid.data_ptr(), capacity, n_out.data_ptr(), ws.data_ptr(), ws_bytes,
                                        _stream(dev)), "agnn_window_subgraph")
    _lib.count_launches(5)
    total = int(n_out.item())
    node_index = torch.cat([torch.arange(lo, lo + w, device=dev) for lo, w in zip(node_lo, win)]) if win else \
        torch.zeros(0, dtype=torch.int64, device=dev)
    return edges[:, :total], eid[:total], node_index, out_off + [off]


def neighbor_sample(csr: graph.CSR, n_nodes: int, seeds: torch.Tensor, fanouts: Sequence[int], seed: int):
    """k-hop sampling on a destination-keyed relation-major CSR (``graph.build_csr`` of (dst, src, etype)).

    Returns a dict like oracle/graph.py::neighbor_sample: ``node`` (global ids, discovery order),
    per relation ``src`` / ``dst`` (batch-local) and ``edge`` (CSR slot), ``num_sampled_nodes``,
    ``num_sampled_edges[r]`` per hop."""
    dev = csr.rowptr.device
    lib = _lib.lib()
    r = csr.n_rel
    rowptr = csr.rowptr.reshape(-1)
    seeds32 = seeds.to(torch.int32).contiguous()
    n_seeds = int(seeds32.numel())
    local = torch.empty(max(n_nodes, 1), dtype=torch.int32, device=dev)
    first_pos = torch.empty(max(n_nodes, 1), dtype=torch.int32, device=dev)
    nodes = torch.empty(max(n_nodes, 1), dtype=torch.int32, device=dev)
    st = _stream(dev)
    _lib.check(lib.agnn_sample_init(n_nodes, seeds32.data_ptr(), n_seeds, local.data_ptr(), nodes.data_ptr(), st),
               "agnn_sample_init")
    n_known, lo = n_seeds, 0
    nodes_per_hop = [n_seeds]
    per_rel = {k: [[] for _ in range(r)] for k in ("src", "dst", "edge")}
    edges_per_hop = [[] for _ in range(r)]
    base = int(rowptr[0].item()) if rowptr.numel() else 0      # slots are relative to the segment's col block
    for hop, k in enumerate(fanouts):
        f = n_known - lo
        counts = torch.empty(r * f + 1, dtype=torch.int32, device=dev)
        ws_bytes = lib.agnn_sample_hop_workspace(r, f)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.agnn_sample_hop_count(r, n_nodes, rowptr.data_ptr(), nodes.data_ptr(), lo, f, int(k),
                                             counts.data_ptr(), ws.data_ptr(), ws_bytes, st), "agnn_sample_hop_count")
        offs = counts.cpu()
        n_cand = int(offs[-1])
        cand = torch.empty((4, max(n_cand, 1)), dtype=torch.int32, device=dev)
        flag = torch.empty(n_cand + 1, dtype=torch.int32, device=dev)
        ws2 = torch.empty(max((n_cand // 4096 + 2) * 4, 16), dtype=torch.uint8, device=dev)
        _lib.check(lib.agnn_sample_hop_draw(r, n_nodes, rowptr.data_ptr(), csr.col.data_ptr(), nodes.data_ptr(), lo, f,
                                            n_known, hop, int(k), int(seed) & MASK64, counts.data_ptr(), n_cand,
                                            cand[0].data_ptr(), cand[1].data_ptr(), cand[2].data_ptr(),
                                            cand[3].data_ptr(), local.data_ptr(), first_pos.data_ptr(),
                                            flag.data_ptr(), ws2.data_ptr(), ws2.numel(), st), "agnn_sample_hop_draw")
        _lib.count_launches(12)
        n_new = int(flag[-1].item())
        for rel in range(r):
            a, b = int(offs[rel * f]), int(offs[(rel + 1) * f])
            per_rel["src"][rel].append(cand[3, a:b].long())
            per_rel["dst"][rel].append(cand[2, a:b].long())
            per_rel["edge"][rel].append(cand[0, a:b].long() - base)
            edges_per_hop[rel].append(b - a)
        lo, n_known = n_known, n_known + n_new
        nodes_per_hop.append(n_new)
    cat = lambda parts: torch.cat(parts) if parts else torch.zeros(0, dtype=torch.int64, device=dev)
    return {"node": nodes[:n_known].long(), "src": [cat(v) for v in per_rel["src"]],
            "dst": [cat(v) for v in per_rel["dst"]], "edge": [cat(v) for v in per_rel["edge"]],
            "num_sampled_nodes": nodes_per_hop, "num_sampled_edges": edges_per_hop}


class Corpus:
    """Scores resident on the device: note features, typed edges with global node ids, per-score ranges."""

    def __init__(self, x: torch.Tensor, edges: torch.Tensor, node_ptr: Sequence[int], n_rel: int = 4,
                 extras: Optional[Dict[str, torch.Tensor]] = None):
        self.x, self.edges, self.n_rel = x, edges, n_rel
        self.node_ptr = [int(v) for v in node_ptr]
        self.extras = extras or {}
        score_of_edge = torch.bucketize(edges[0], torch.tensor(self.node_ptr[1:], device=edges.device), right=True)
        counts = torch.bincount(score_of_edge, minlength=len(self.node_ptr) - 1).cpu().tolist()
        self.edge_ptr = [0]
        for c in counts:
            self.edge_ptr.append(self.edge_ptr[-1] + c)
        self._csr = None

    @property
    def n_scores(self):
        return len(self.node_ptr) - 1

    def csr_by_destination(self) -> graph.CSR:
        if self._csr is None:
            n = self.node_ptr[-1]
            self._csr = graph.build_csr([graph.Segment(self.edges[1], self.edges[0], n, n, self.edges[2], self.n_rel)])[0]
        return self._csr


class NodeStore(dict):
    """The part of a PyG node storage the reference's training step reads (analysisgnn/models/analysis.py:926-968):
    ``batch["note"][key]``, ``.keys()``, attribute access (``.pitch_spelling``, ``.batch_size``)."""

    def __init__(self, tensors: Dict[str, torch.Tensor], batch_size: int):
        super().__init__(tensors)
        self.batch_size = int(batch_size)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None


class HeteroBatch:
    """What ``MuseNeighborLoader(..., transform=transform_to_pyg)`` yields, as far as ``ContinualAnalysisGNN.common_step``
    uses it (analysis.py:947-961): ``.x_dict``, ``.edge_index_dict``, ``.batch_dict``, ``.num_sampled_nodes_dict``,
    ``.num_sampled_edges_dict`` and ``batch["note"]`` with ``.batch_size``, the features, the graph ids and every
    per-note array of the corpus (labels, spellings, onsets).  Built from ``ScoreGraphLoader.batch``'s dict."""

    def __init__(self, out: dict):
        self.x_dict = out["x_dict"]
        self.edge_index_dict = out["edge_index_dict"]
        self.batch_dict = out["batch_dict"]
        self.num_sampled_nodes_dict = out.get("num_sampled_nodes_dict")
        self.num_sampled_edges_dict = out.get("num_sampled_edges_dict")
        self.graph_ids = out.get("graph_ids")
        self.node_index = out.get("node_index")
        self._stores = {}
        for t, x in self.x_dict.items():
            fields = {"x": x, "batch": self.batch_dict[t]}
            if t == "note":
                fields.update(out.get("extras", {}))
            self._stores[t] = NodeStore(fields, out["batch_size"] if t == "note" else x.shape[0])

    def __getitem__(self, node_type: str) -> NodeStore:
        return self._stores[node_type]

    @property
    def node_types(self):
        return list(self._stores)

    @property
    def edge_types(self):
        return list(self.edge_index_dict)


class ScoreGraphLoader:
    """Batches of ``batch_size`` score windows (``subgraph_size`` notes each), optionally extended by
    ``num_neighbors`` sampled hops, in the dict layout ``TorchAnalysisGNN.encode`` consumes
    (analysisgnn/models/analysis.py:948-961): ``x_dict, edge_index_dict, batch_dict, batch_size,
    num_sampled_nodes_dict, num_sampled_edges_dict`` plus gathered ``extras`` (labels, spellings)."""

    REL_NAMES = ("onset", "consecutive", "during", "rest")

    def __init__(self, corpus: Corpus, subgraph_size: int, batch_size: int, num_neighbors: Sequence[int] = (),
                 seed: int = 0, shuffle: bool = True, rank: int = 0, world_size: int = 1,
                 subgraph_sample_ratio: Optional[float] = None):
        self.corpus, self.subgraph_size, self.batch_size = corpus, subgraph_size, batch_size
        self.num_neighbors, self.seed, self.shuffle = list(num_neighbors), seed, shuffle
        self.rank, self.world_size = rank, world_size
        # The reference passes ``subgraph_sample_ratio=0.5`` to every MuseNeighborLoader (datamodules/analysis.py:276,
        # 290, 305, 320).  graphmuse (third party, not in /root/reference) documents it as the number of windows an
        # epoch draws from a score relative to how many windows fit in it: a score of n notes is visited
        # ``max(1, ceil(ratio * n / subgraph_size))`` times per epoch, each visit with its own random window.  That
        # published behaviour is what is restated here (parity unpinned: the third-party source is absent); ``None``
        # keeps one visit per score.
        self.subgraph_sample_ratio = subgraph_sample_ratio
        if subgraph_sample_ratio is None:
            self.visits = [1] * corpus.n_scores
        else:
            if not subgraph_sample_ratio > 0:
                raise ValueError("subgraph_sample_ratio must be positive")
            ptr = corpus.node_ptr
            # integer arithmetic where the ratio allows it (0.5, 2, ...): no float rounding at exact multiples
            num, den = float(subgraph_sample_ratio).as_integer_ratio()
            self.visits = [max(1, -((-(ptr[g + 1] - ptr[g]) * num) // (den * subgraph_size)))
                           for g in range(corpus.n_scores)]
        self.epoch_size = sum(self.visits)

    def __len__(self):
        return (self.epoch_size + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        """One epoch of ``HeteroBatch``es (what the Lightning loop iterates over); every ``iter()`` starts the next
        epoch, so shuffling differs from epoch to epoch and is the same on every rank."""
        epoch = getattr(self, "_epoch", 0)
        self._epoch = epoch + 1
        for index in range(len(self)):                     # every rank yields len(self) batches (see batch_ids)
            yield HeteroBatch(self.batch(epoch, index))

    def batches(self, epoch: int) -> List[List[int]]:
        """The epoch's global batches (score ids).  One visit per score: the shuffled score order cut into consecutive
        batches.  With ``subgraph_sample_ratio``: the shuffled scores are laid out visit by visit (a score's visits
        next to each other) and position ``j`` goes to batch ``j mod len(self)``, so the visits of one score land in
        DIFFERENT batches (as long as it has no more visits than the epoch has batches) and batch sizes differ by at
        most one."""
        cached = getattr(self, "_batches", None)
        if cached is not None and cached[0] == epoch:
            return cached[1]
        ids = list(range(self.corpus.n_scores))
        if self.shuffle:
            ids.sort(key=lambda g: rng_u64(self.seed, 0x5348, epoch, g, 0))
        n = len(self)
        if self.subgraph_sample_ratio is None:
            out = [ids[i * self.batch_size:(i + 1) * self.batch_size] for i in range(n)]
        else:
            layout = [g for g in ids for _ in range(self.visits[g])]
            out = [layout[i::n] for i in range(n)]
        self._batches = (epoch, out)
        return out

    def order(self, epoch: int) -> List[int]:
        """The epoch's visits, batch after batch."""
        return [g for b in self.batches(epoch) for g in b]

    def _global_batch(self, epoch: int, index: int) -> List[int]:
        ids = list(self.batches(epoch)[index])
        if 0 < len(ids) < self.world_size:
            # a short last batch with fewer scores than ranks: wrap around the epoch order (DistributedSampler's
            # padding) so that EVERY rank has a share -- ranks must run the same number of steps, or the others
            # block in the gradient allreduce
            order = self.order(epoch)
            ids = ids + [order[i % len(order)] for i in range(self.world_size - len(ids))]
        return ids

    def batch_ids(self, epoch: int, index: int) -> List[int]:
        """Scores of global batch ``index`` that this rank takes (data parallel: positions ``g mod W == rank`` of the
        batch, so the ranks' shares are disjoint and together are the global batch).  Host arithmetic only."""
        return self._global_batch(epoch, index)[self.rank::self.world_size]

    def batch_draws(self, epoch: int, index: int) -> List[int]:
        """For every entry of ``batch_ids``: how many earlier positions of the GLOBAL batch hold the same score.  The
        window start is keyed on it (``window_start(..., draw)``), so two visits of one score that land in the same
        batch get different windows, and a rank's windows do not depend on the number of ranks."""
        seen: Dict[int, int] = {}
        draws = []
        for g in self._global_batch(epoch, index):
            draws.append(seen.get(g, 0))
            seen[g] = draws[-1] + 1
        return draws[self.rank::self.world_size]

    def window_starts(self, epoch: int, index: int) -> List[int]:
        """Window start of every entry of ``batch_ids`` (the counter RNG; host arithmetic only)."""
        c = self.corpus
        step_seed = rng_u64(self.seed, 0x424154, epoch, index, 0)
        return [window_start(step_seed, g, c.node_ptr[g + 1] - c.node_ptr[g], self.subgraph_size, k)
                for g, k in zip(self.batch_ids(epoch, index), self.batch_draws(epoch, index))]

    def batch(self, epoch: int, index: int):
        c = self.corpus
        ids = self.batch_ids(epoch, index)
        step_seed = rng_u64(self.seed, 0x424154, epoch, index, 0)
        starts = self.window_starts(epoch, index)
        edges, _, node_index, slot_ptr = window_subgraphs(c.edges[0], c.edges[1], c.edges[2], c.node_ptr, c.edge_ptr,
                                                          ids, starts, self.subgraph_size)
        n_target = slot_ptr[-1]
        dev = edges.device
        batch_vec = torch.repeat_interleave(torch.arange(len(ids), device=dev),
                                            torch.tensor(np.diff(slot_ptr), device=dev))
        out = {"batch_size": n_target, "graph_ids": ids}
        if not self.num_neighbors:
            ei = {("note", name, "note"): edges[:2, edges[2] == k] for k, name in enumerate(self.REL_NAMES[:c.n_rel])}
            out.update(node_index=node_index, edge_index_dict=ei, num_sampled_nodes_dict=None,
                       num_sampled_edges_dict=None)
        else:
            if len(set(ids)) != len(ids):
                raise ValueError("ScoreGraphLoader: a score appears twice in one batch (more visits per epoch than "
                                 "batches); the k-hop sampler grows ONE node set per score -- lower "
                                 "subgraph_sample_ratio or the batch size")
            s = neighbor_sample(c.csr_by_destination(), c.node_ptr[-1], node_index, self.num_neighbors, step_seed)
            node_index = s["node"]
            ei = {("note", name, "note"): torch.stack((s["src"][k], s["dst"][k]))
                  for k, name in enumerate(self.REL_NAMES[:c.n_rel])}
            score_of = torch.bucketize(node_index, torch.tensor(c.node_ptr[1:], device=dev), right=True)
            remap = torch.full((c.n_scores,), -1, dtype=torch.long, device=dev)
            remap[torch.tensor(ids, device=dev)] = torch.arange(len(ids), device=dev)
            batch_vec = remap[score_of]
            out.update(node_index=node_index, edge_index_dict=ei,
                       num_sampled_nodes_dict={"note": s["num_sampled_nodes"]},
                       num_sampled_edges_dict={("note", name, "note"): s["num_sampled_edges"][k]
                                               for k, name in enumerate(self.REL_NAMES[:c.n_rel])})
        out["x_dict"] = {"note": c.x.index_select(0, node_index)}
        out["batch_dict"] = {"note": batch_vec}
        out["extras"] = {k: v.index_select(0, node_index) for k, v in c.extras.items()}
        return out


class StaticBatcher:
    """Batches of fixed SHAPE built entirely on the device from a resident ``Corpus`` -- the form a CUDA graph can
    replay: the whole loader step (window selection, induced note -> note edges, beat / measure nodes and their
    edges, feature and label gathers) is a fixed sequence of launches on static buffers, fed by ONE small host -> device
    copy per step (``select``: the ``[2, batch_size]`` score ids and window starts, host arithmetic of the counter RNG).

    What varies from batch to batch lives in padding: the typed note -> note COO has ``edge_cap`` slots (the maximum any
    ``batch_size`` windows of this corpus can need), unused ones carry type / index -1 and are dropped by
    agnn_csr_build; beat / measure node arrays have ``beat_cap`` / ``measure_cap`` rows, the unused ones are isolated
    zero-feature nodes that no note ever sees.  Every score must have at least ``subgraph_size`` notes (each window
    then has exactly that many: the sequence layout of the GRU branch is static too); shorter scores belong to the
    eager ``ScoreGraphLoader``.

    Replaces graphmuse ``MuseNeighborLoader`` + ``transform_to_pyg`` with ``add_beats / add_measures``
    (analysisgnn/data/datamodules/analysis.py:217-225, 270-293; node and edge types per analysisgnn/utils/hgraph.py:
    41-73 and data/data_utils.py:194) for full-window batches (``num_neighbors`` hops: ``ScoreGraphLoader``)."""

    REL = ScoreGraphLoader.REL_NAMES

    def __init__(self, corpus: Corpus, subgraph_size: int, batch_size: int, beat_of: Optional[torch.Tensor] = None,
                 measure_of: Optional[torch.Tensor] = None, reverse: bool = True):
        c, s, b = corpus, int(subgraph_size), int(batch_size)
        dev = c.x.device
        if dev.type != "cuda":
            raise _lib.AgnnError("analysisgnn_b200 has no CPU path: the sampler needs CUDA tensors")
        sizes = np.diff(np.asarray(c.node_ptr))
        if sizes.min() < s:
            raise ValueError(f"StaticBatcher: every score needs >= {s} notes (shortest has {int(sizes.min())})")
        self.corpus, self.s, self.b, self.reverse = c, s, b, reverse
        # no CSR row of a batch is longer than a window (graph.set_degree_bound: the caller may pass this on)
        self.degree_bound = s
        i64 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.int64, device=dev)
        self.node_ptr, self.edge_ptr = i64(c.node_ptr), i64(c.edge_ptr)
        n = c.node_ptr[-1]
        # ---- capacities (one pass over the corpus, once): the most edges any window holds, the widest beat span
        src, dst = c.edges[0], c.edges[1]
        lo, hi = torch.minimum(src, dst), torch.maximum(src, dst)
        score = torch.bucketize(lo, self.node_ptr[1:], right=True)
        inside = (hi - lo) < s
        first = torch.maximum(hi - s + 1, self.node_ptr[score])[inside]     # first / last window start holding the edge
        last = lo[inside]
        diff = torch.zeros(n + 2, dtype=torch.int64, device=dev)
        diff.index_add_(0, first, torch.ones_like(first))
        diff.index_add_(0, last + 1, -torch.ones_like(last))
        per_window = int(torch.cumsum(diff, 0).max()) if first.numel() else 0
        self.edge_cap = max(b * per_window, 1)
        self.cand_cap = max(b * int(np.diff(np.asarray(c.edge_ptr)).max()), 1)
        note_score = torch.bucketize(torch.arange(n, device=dev), self.node_ptr[1:], right=True)
        idx = torch.arange(max(n - s + 1, 0), device=dev)
        same = note_score[idx] == note_score[idx + s - 1]
        self.virtual = {}
        for name, of in (("beat", beat_of), ("measure", measure_of)):
            if of is None:
                continue
            of = of.to(dev, torch.int64)
            span = torch.where(same, of[idx + s - 1] - of[idx] + 1, torch.zeros_like(idx))
            self.virtual[name] = (of, max(b * int(span.max()), 1))
        # ---- constants of every batch
        self.arange_s = torch.arange(s, device=dev)
        self.win = torch.full((b,), s, dtype=torch.int32, device=dev)
        self.out_off = torch.arange(b, device=dev) * s
        self.note_batch = torch.arange(b, device=dev).repeat_interleave(s)
        self.note_ids = torch.arange(b * s, device=dev)
        names = [("note", r, "note") for r in self.REL[:c.n_rel]]
        if reverse:
            names += [("note", r + "_rev", "note") for r in self.REL[1:c.n_rel]]
        self.names = names
        self.zeros = {name: torch.zeros((cap, c.x.shape[1]), dtype=c.x.dtype, device=dev)
                      for name, (_, cap) in self.virtual.items()}
        self._host = torch.empty((2, b), dtype=torch.int64).pin_memory()

    @property
    def metadata(self):
        node_types = ["note"] + list(self.virtual)
        edge_types = list(self.names)
        for v in self.virtual:
            edge_types += [("note", "connects", v), (v, "connects_rev", "note"), (v, "next", v)]
        return node_types, edge_types

    def select(self, loader: ScoreGraphLoader, epoch: int, index: int) -> torch.Tensor:
        """Host side of a step: the scores of global batch ``index`` this rank takes and their window starts (the
        counter RNG of ``ScoreGraphLoader``), in a pinned ``[2, batch_size]`` tensor to be copied into the static
        device selector (``non_blocking``)."""
        c = self.corpus
        ids = loader.batch_ids(epoch, index)
        if len(ids) != self.b:
            raise ValueError(f"StaticBatcher: batch {index} has {len(ids)} scores on this rank, the static shape is "
                             f"{self.b} (use a corpus size that is a multiple of batch_size x world_size)")
        if loader.subgraph_size != self.s:
            raise ValueError(f"StaticBatcher: the loader draws {loader.subgraph_size}-note windows, the static shape "
                             f"is {self.s}")
        starts = loader.window_starts(epoch, index)
        self._host[0] = torch.tensor(ids, dtype=torch.int64)
        self._host[1] = torch.tensor(starts, dtype=torch.int64)
        return self._host

    def batch(self, sel: torch.Tensor) -> dict:
        """``sel``: DEVICE int64 ``[2, batch_size]`` (score ids, window starts).  No host read, no data-dependent
        shape: capturable.  Returns the dict layout of ``ScoreGraphLoader.batch``."""
        c, s, b = self.corpus, self.s, self.b
        dev = sel.device
        ids, starts = sel[0], sel[1]
        node_lo = self.node_ptr[ids] + starts
        node_index = (node_lo.unsqueeze(1) + self.arange_s).reshape(-1)
        # note -> note edges induced by the windows, in corpus order per slot
        e_lo = self.edge_ptr[ids]
        cand_ptr = torch.zeros(b + 1, dtype=torch.int64, device=dev)
        cand_ptr[1:] = torch.cumsum(self.edge_ptr[ids + 1] - e_lo, 0)
        lib = _lib.lib()
        ws_bytes = lib.agnn_window_workspace(self.cand_cap)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        n_out = torch.zeros(1, dtype=torch.int32, device=dev)
        edges = torch.full((3, self.edge_cap), -1, dtype=torch.int64, device=dev)
        _lib.check(lib.agnn_window_subgraph(b, self.cand_cap, cand_ptr.data_ptr(), e_lo.data_ptr(), node_lo.data_ptr(),
                                            self.win.data_ptr(), self.out_off.data_ptr(), c.edges[0].data_ptr(),
                                            c.edges[1].data_ptr(), c.edges[2].data_ptr(), edges.data_ptr(), None,
                                            self.edge_cap, n_out.data_ptr(), ws.data_ptr(), ws_bytes, _stream(dev)),
                   "agnn_window_subgraph")
        _lib.count_launches(5)
        ei, et = edges[:2], edges[2]
        if self.reverse:
            rev = torch.where(et > 0, et + (c.n_rel - 1), torch.full_like(et, -1))
            ei, et = torch.cat((ei, ei.flip(0)), dim=1), torch.cat((et, rev))
        plain, x_dict, batch_dict = {}, {"note": c.x.index_select(0, node_index)}, {"note": self.note_batch}
        for name, (of, cap) in self.virtual.items():
            first, last = of[node_lo], of[node_lo + (s - 1)]
            count = last - first + 1
            end = torch.cumsum(count, 0)
            off = end - count
            local = (of[node_index].view(b, s) - first.unsqueeze(1) + off.unsqueeze(1)).reshape(-1)
            up = torch.stack((self.note_ids, local))
            k = torch.arange(cap, device=dev)
            slot = torch.bucketize(k, end, right=True)
            slot_c = slot.clamp(max=b - 1)
            ok = (slot < b) & (k + 1 < end[slot_c])
            nxt = torch.where(ok.unsqueeze(0), torch.stack((k, k + 1)), torch.full((2, cap), -1, dtype=torch.int64,
                                                                                    device=dev))
            plain[("note", "connects", name)] = up
            plain[(name, "connects_rev", "note")] = up.flip(0)
            plain[(name, "next", name)] = nxt
            x_dict[name] = self.zeros[name]
            batch_dict[name] = slot_c
        eid = graph.TypedEdgeDict([graph.TypedEdges(ei, et, self.names)], plain)
        return {"batch_size": b * s, "node_index": node_index, "x_dict": x_dict, "edge_index_dict": eid,
                "batch_dict": batch_dict, "num_sampled_nodes_dict": None, "num_sampled_edges_dict": None,
                "extras": {k: v.index_select(0, node_index) for k, v in c.extras.items()}, "n_edges": n_out}
