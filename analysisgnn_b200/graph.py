"""Device-side graph structure for the message-passing kernels.

Everything the layers need from a batch's ``edge_index`` tensors is one call of
``agnn_csr_build`` (include/agnn.h): per relation a CSR keyed on the REDUCE side
(used by the forward) and a CSR keyed on the GATHERED side (the transposed graph,
used by the backward), both stable in input edge order, int32.

Two layouts mirror the two conventions in the reference:

* ``TypedCSR``  -- one node space, ``edge_index [2,E]`` + ``edge_type [E]``; the
  in-tree layers reduce at ``edge_index[0]`` reading ``edge_index[1]``
  (analysisgnn/models/core/gnn.py:70-74, hgnn.py:480-483).
* ``HeteroCSR`` -- PyG ``edge_index_dict``; reduces at row 1 reading row 0.

Structures are cached per batch (``cached``), keyed on the identity and version
of the index tensors, so a stack of L layers builds them once.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise _lib.AgnnError(f"{what} must be a CUDA tensor: analysisgnn_b200 has no CPU path")


def _index_row(t: torch.Tensor) -> torch.Tensor:
    """A stride-1 int64 view (or copy) of a 1-D index tensor."""
    if t.dtype != torch.int64:
        t = t.long()
    if t.dim() != 1:
        raise ValueError("index rows must be 1-D")
    return t if t.stride(0) == 1 or t.numel() <= 1 else t.contiguous()


class Segment:
    """One COO segment handed to agnn_csr_build."""

    def __init__(self, row, col, n_rows, n_cols, etype=None, n_rel=1):
        self.row, self.col = _index_row(row), _index_row(col)
        self.etype = None if etype is None else _index_row(etype)
        self.n_rows, self.n_cols, self.n_rel = int(n_rows), int(n_cols), int(n_rel)
        self.n_edges = int(self.row.numel())
        if self.col.numel() != self.n_edges or (self.etype is not None and self.etype.numel() != self.n_edges):
            raise ValueError("row / col / etype must have the same length")


class CSR:
    """Result for one segment: ``rowptr`` [n_rel, n_rows+1] (absolute positions in
    ``col`` / ``perm``), ``col`` [E] gathered-side ids, ``perm`` [E] input positions."""

    __slots__ = ("rowptr", "col", "perm", "n_rows", "n_cols", "n_rel", "n_edges", "heavy", "n_heavy", "heavy_cap")

    def __init__(self, rowptr, col, perm, seg: Segment, heavy=None, n_heavy=None, heavy_cap=0):
        self.rowptr, self.col, self.perm = rowptr, col, perm
        self.n_rows, self.n_cols, self.n_rel, self.n_edges = seg.n_rows, seg.n_cols, seg.n_rel, seg.n_edges
        # rows with >= HEAVY_ROW entries, per relation: [n_rel, 2 heavy_cap] (ascending row ids, then the exclusive
        # prefix of their chunk counts) + [n_rel] counts (device)
        self.heavy, self.n_heavy, self.heavy_cap = heavy, n_heavy, heavy_cap


_degree_bound: Optional[int] = None


def set_degree_bound(bound: Optional[int]) -> Optional[int]:
    """Performance hint for the CSRs built from now on (``None`` = unknown, the default): no row of them has more than
    ``bound`` entries -- e.g. a batch of disjoint ``subgraph_size``-note windows, where every neighbour of a node, every
    note of a beat / measure and every node of a pooled graph lies inside one window.  Below the library's hub-row
    threshold (``_lib.HEAVY_ROW``) ``build_csr`` then leaves out the hub-row lists and every aggregation leaves out its two
    hub-row launches (3 launches -> 1; ~40 launches of a headline training step).  A violated hint costs speed, never
    correctness: without a list every row, however long, is reduced by its own warp.  Returns the previous value."""
    global _degree_bound
    old, _degree_bound = _degree_bound, (None if bound is None else int(bound))
    if old != _degree_bound:
        _cache.clear()               # cached CSRs were built under the other promise
    return old


def build_csr(segments: Sequence[Segment], device=None, validate: bool = False) -> List[CSR]:
    """Convert COO segments to CSR on the GPU (one library call per 32 segments)."""
    if not segments:
        return []
    device = device if device is not None else segments[0].row.device
    for s in segments:
        _require_cuda(s.row, "edge_index")
    lib = _lib.lib()
    key_total = sum(s.n_rel * (s.n_rows + 1) for s in segments)
    edge_total = sum(s.n_edges for s in segments)
    rowptr = torch.empty(key_total, dtype=torch.int32, device=device)
    col = torch.empty(max(edge_total, 1), dtype=torch.int32, device=device)
    perm = torch.empty(max(edge_total, 1), dtype=torch.int32, device=device)
    status = torch.zeros(1, dtype=torch.int32, device=device)
    hubs = _degree_bound is None or _degree_bound >= _lib.HEAVY_ROW     # can a row reach the hub-row threshold?
    caps = [s.n_edges // _lib.HEAVY_ROW + 1 if hubs else 0 for s in segments]
    # per relation: [cap] heavy row ids (ascending) + [cap] exclusive prefix of their chunk counts
    heavy = torch.empty(sum(s.n_rel * 2 * c for s, c in zip(segments, caps)), dtype=torch.int32, device=device) \
        if hubs else None
    n_heavy = torch.empty(sum(s.n_rel for s in segments), dtype=torch.int32, device=device) if hubs else None
    stream = torch.cuda.current_stream(device).cuda_stream
    out, k_off, e_off, h_off, c_off = [], 0, 0, 0, 0
    for lo in range(0, len(segments), _lib.MAX_SEG):
        chunk = segments[lo:lo + _lib.MAX_SEG]
        arr = (_lib.Coo * len(chunk))()
        for i, s in enumerate(chunk):
            arr[i].row = s.row.data_ptr() if s.n_edges else None
            arr[i].col = s.col.data_ptr() if s.n_edges else None
            arr[i].etype = s.etype.data_ptr() if (s.etype is not None and s.n_edges) else None
            arr[i].n_edges, arr[i].n_rows, arr[i].n_cols, arr[i].n_rel = s.n_edges, s.n_rows, s.n_cols, s.n_rel
            arr[i].rowptr_off, arr[i].edge_off = k_off, e_off
            cap = caps[lo + i]
            arr[i].heavy_off, arr[i].count_off, arr[i].heavy_cap = h_off, c_off, cap
            keys = s.n_rel * (s.n_rows + 1)
            out.append(CSR(rowptr[k_off:k_off + keys].view(s.n_rel, s.n_rows + 1), col[e_off:e_off + s.n_edges],
                           perm[e_off:e_off + s.n_edges], s,
                           heavy[h_off:h_off + s.n_rel * 2 * cap].view(s.n_rel, 2 * cap) if hubs else None,
                           n_heavy[c_off:c_off + s.n_rel] if hubs else None, cap))
            k_off += keys
            e_off += s.n_edges
            h_off += s.n_rel * 2 * cap
            c_off += s.n_rel
        ws_bytes = lib.agnn_csr_build_workspace(len(chunk), arr)
        if ws_bytes == 0:
            raise _lib.AgnnError("agnn_csr_build_workspace: " + lib.agnn_last_error().decode())
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
        _lib.check(lib.agnn_csr_build(len(chunk), arr, rowptr.data_ptr(), col.data_ptr(), perm.data_ptr(),
                                      status.data_ptr(), heavy.data_ptr() if hubs else None,
                                      n_heavy.data_ptr() if hubs else None, ws.data_ptr(), ws_bytes,
                                      stream), "agnn_csr_build")
        _lib.count_launches((10 if any(s.n_edges for s in chunk) else 6) - (0 if hubs else 1))
    if validate and int(status.item()) != 0:
        raise ValueError("edge_index contains node ids outside [0, num_nodes)")
    return out


class TypedCSR:
    """In-tree convention: ``fwd`` reduces at ``edge_index[0]`` (rows) reading
    ``edge_index[1]``; ``bwd`` is the transposed graph.  ``n_rel`` relations with
    codes ``0..n_rel-1`` in ``edge_type`` (``None`` = a single relation)."""

    def __init__(self, edge_index, edge_type, n_rows, n_cols=None, n_rel=1, reduce_row=0, validate=False):
        n_cols = n_rows if n_cols is None else n_cols
        r, c = edge_index[reduce_row], edge_index[1 - reduce_row]
        self.fwd, self.bwd = build_csr(
            [Segment(r, c, n_rows, n_cols, edge_type, n_rel), Segment(c, r, n_cols, n_rows, edge_type, n_rel)],
            validate=validate)
        self.n_rows, self.n_cols, self.n_rel = int(n_rows), int(n_cols), int(n_rel)
        self.n_edges = int(edge_index.shape[1])
        self._t = None

    def t(self) -> "TypedCSR":
        """The transposed graph (rows <-> cols) as a view on the same arrays."""
        if self._t is None:
            o = object.__new__(TypedCSR)
            o.fwd, o.bwd = self.bwd, self.fwd
            o.n_rows, o.n_cols, o.n_rel, o.n_edges = self.n_cols, self.n_rows, self.n_rel, self.n_edges
            o._t = self
            self._t = o
        return self._t


class TypedEdges:
    """Several edge types between ONE pair of node types as a typed COO of fixed capacity: ``edge_index`` [2, cap]
    (row 0 = source, row 1 = target), ``edge_type`` [cap] with codes ``0..len(names)-1`` and ``-1`` for unused slots
    (dropped by agnn_csr_build).  This is what a static-shape (CUDA-graph replayable) loader emits: no per-type
    compaction, so no shape depends on the data."""

    def __init__(self, edge_index: torch.Tensor, edge_type: torch.Tensor, names: Sequence[Tuple[str, str, str]]):
        self.edge_index, self.edge_type = edge_index, edge_type
        self.names = [tuple(n) for n in names]
        if len({(n[0], n[2]) for n in self.names}) != 1:
            raise ValueError("TypedEdges: all edge types must connect the same pair of node types")


class TypedEdgeDict(dict):
    """An ``edge_index_dict`` backed by ``TypedEdges`` groups (plus ordinary ``[2, E]`` entries).  Looks like the PyG
    dict to generic consumers -- ``d[edge_type]`` is a ``[2, cap]`` tensor whose slots of other types are -1 (made on
    first use) -- while ``hetero_csr`` builds every group in one typed pass."""

    def __init__(self, groups: Sequence[TypedEdges], plain: Optional[dict] = None):
        super().__init__()
        self.groups = list(groups)
        self._where = {}
        for gi, g in enumerate(self.groups):
            for k, name in enumerate(g.names):
                self._where[name] = (gi, k)
                dict.__setitem__(self, name, None)
        for k, v in (plain or {}).items():
            dict.__setitem__(self, tuple(k), v)

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if v is None:
            gi, k = self._where[key]
            g = self.groups[gi]
            v = torch.where((g.edge_type == k).unsqueeze(0), g.edge_index, torch.full_like(g.edge_index, -1))
            dict.__setitem__(self, key, v)
        return v

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]

    def index_tensors(self):
        """The tensors whose identity / version define the structure (cache key)."""
        out = []
        for g in self.groups:
            out += [g.edge_index, g.edge_type]
        return out + [dict.__getitem__(self, k) for k in self.keys() if k not in self._where]


class _CSRView:
    """Relation ``k`` of a typed CSR as a single-relation CSR (what ``ops.rel_of(csr, 0, ...)`` reads)."""

    __slots__ = ("rowptr", "col", "perm", "n_rows", "n_cols", "n_rel", "n_edges", "heavy", "n_heavy", "heavy_cap")

    def __init__(self, parent: CSR, k: int):
        self.rowptr, self.col, self.perm = parent.rowptr[k:k + 1], parent.col, parent.perm
        self.n_rows, self.n_cols, self.n_rel, self.n_edges = parent.n_rows, parent.n_cols, 1, parent.n_edges
        self.heavy = parent.heavy[k:k + 1] if parent.heavy is not None else None
        self.n_heavy = parent.n_heavy[k:k + 1] if parent.n_heavy is not None else None
        self.heavy_cap = parent.heavy_cap


class HeteroCSR:
    """PyG convention for an ``edge_index_dict``: per edge type ``(src, rel, dst)``
    ``fwd[et]`` has rows = dst nodes / cols = src ids, ``bwd[et]`` rows = src nodes /
    cols = dst ids.  A ``TypedEdgeDict`` is built group by group (one typed segment pair per group)."""

    def __init__(self, edge_index_dict, num_nodes: Dict[str, int], validate=False):
        self.edge_types: List[Tuple[str, str, str]] = [tuple(et) for et in edge_index_dict.keys()]
        self.num_nodes = {k: int(v) for k, v in num_nodes.items()}
        groups = edge_index_dict.groups if isinstance(edge_index_dict, TypedEdgeDict) else []
        grouped = {n for g in groups for n in g.names}
        plain = [et for et in self.edge_types if et not in grouped]
        segs = []
        for g in groups:
            ns, nd = self.num_nodes[g.names[0][0]], self.num_nodes[g.names[0][2]]
            ei, r = g.edge_index, len(g.names)
            segs.append(Segment(ei[1], ei[0], nd, ns, g.edge_type, r))
            segs.append(Segment(ei[0], ei[1], ns, nd, g.edge_type, r))
        for et in plain:
            ei = edge_index_dict[et]
            ns, nd = self.num_nodes[et[0]], self.num_nodes[et[2]]
            segs.append(Segment(ei[1], ei[0], nd, ns))
            segs.append(Segment(ei[0], ei[1], ns, nd))
        built = build_csr(segs, validate=validate)
        self.fwd, self.bwd = {}, {}
        for gi, g in enumerate(groups):
            for k, name in enumerate(g.names):
                if name in self.edge_types:
                    self.fwd[name], self.bwd[name] = _CSRView(built[2 * gi], k), _CSRView(built[2 * gi + 1], k)
        base = 2 * len(groups)
        for i, et in enumerate(plain):
            self.fwd[et], self.bwd[et] = built[base + 2 * i], built[base + 2 * i + 1]
        self.n_edges = {et: self.fwd[et].n_edges for et in self.edge_types}


# ---------------------------------------------------------------------- cache

class _StructureCache:
    """Small LRU keyed on tensor identity; an entry is stale when a tensor's version counter moved
    (in-place update = new batch).  Entries hold the index tensors, so an address can never be
    recycled under a live key.  ``frozen`` (set while a CUDA graph is captured) accepts entries whose
    tensors were overwritten in place: a captured step promises static layouts."""

    def __init__(self, capacity=8):
        self.capacity = capacity
        self.frozen = False
        self.entries: "OrderedDict[tuple, tuple]" = OrderedDict()

    def get(self, tensors: Sequence[torch.Tensor], extra: tuple, make):
        key = tuple((id(t), tuple(t.shape)) for t in tensors) + extra
        versions = tuple(t._version for t in tensors)
        hit = self.entries.get(key)
        if hit is not None and (self.frozen or hit[1] == versions):
            self.entries.move_to_end(key)
            return hit[2]
        value = make()
        self.entries[key] = (list(tensors), versions, value)
        self.entries.move_to_end(key)
        while len(self.entries) > self.capacity:
            self.entries.popitem(last=False)
        return value

    def clear(self):
        self.entries.clear()


_cache = _StructureCache(capacity=64)        # device structures (CSR, derived index tensors)
_host_cache = _StructureCache(capacity=64)   # layouts that needed a device -> host read


def freeze_host_layouts(frozen: bool) -> None:
    """While True, sequence layouts cached from earlier (eager) steps are reused even if their index
    tensors were overwritten in place -- required inside a CUDA-graph capture, where the device ->
    host read a fresh layout needs is illegal."""
    _host_cache.frozen = bool(frozen)


def clear_cache(host_layouts: bool = False):
    """Forget the per-batch device structures (and, on request, the host-side sequence layouts)."""
    _cache.clear()
    if host_layouts:
        _host_cache.clear()


def typed_csr(edge_index, edge_type, n_rows, n_rel, reduce_row=0, n_cols=None) -> TypedCSR:
    tensors = [edge_index] + ([edge_type] if edge_type is not None else [])
    return _cache.get(tensors, ("typed", int(n_rows), int(n_cols or n_rows), int(n_rel), reduce_row),
                      lambda: TypedCSR(edge_index, edge_type, n_rows, n_cols, n_rel, reduce_row))


def hetero_csr(edge_index_dict, num_nodes: Dict[str, int]) -> HeteroCSR:
    if isinstance(edge_index_dict, HeteroCSR):
        return edge_index_dict
    ets = list(edge_index_dict.keys())
    if isinstance(edge_index_dict, TypedEdgeDict):
        tensors = edge_index_dict.index_tensors()
    else:
        tensors = [edge_index_dict[et] for et in ets]
    extra = ("hetero", tuple(ets), tuple(sorted(num_nodes.items())))
    return _cache.get(tensors, extra, lambda: HeteroCSR(edge_index_dict, num_nodes))


def edge_csr(rows: torch.Tensor, n_rows: int) -> TypedCSR:
    """CSR whose gathered side is the EDGE list itself (col = input edge position):
    reduces per-edge messages ``[E, F]`` at ``rows`` without atomics."""
    def make():
        ids = torch.arange(rows.numel(), dtype=torch.long, device=rows.device)
        return TypedCSR(torch.stack((rows, ids)), None, n_rows, n_cols=int(rows.numel()))
    return _cache.get([rows], ("edge", int(n_rows)), make)


def derived(t: torch.Tensor, key: tuple, make):
    """Cache a tensor derived from index tensor ``t`` (per batch, like the CSR)."""
    return _cache.get([t], ("derived",) + key, make)


class SequenceLayout:
    """How a flat ``[n, C]`` node matrix maps to padded ``[B, T, C]`` sequences
    (uniform lengths = a view; ragged = zero padding)."""

    def __init__(self, sizes: Sequence[int], device):
        self.sizes = [int(v) for v in sizes]
        self.n = sum(self.sizes)
        self.t = max(self.sizes) if self.sizes else 0
        self.uniform = all(v == self.t for v in self.sizes)
        if not self.uniform:
            sz = torch.tensor(self.sizes, dtype=torch.long, device=device)
            valid = torch.arange(self.t, device=device).unsqueeze(0) < sz.unsqueeze(1)     # [B, T]
            self.flat_index = valid.reshape(-1).nonzero(as_tuple=False).squeeze(1)          # padded slot of node k
            self.n_slots = len(self.sizes) * self.t

    @staticmethod
    def from_lengths(lengths: Optional[torch.Tensor], n: int) -> "SequenceLayout":
        """``lengths`` as in analysisgnn/models/core/gnn.py:506-521: ``None`` = one sequence;
        all-equal entries = that many nodes per sequence; otherwise a cumulative pointer
        whose differences are the sequence lengths.  One device->host copy (cached)."""
        if lengths is None:
            return SequenceLayout([n], None)
        host = lengths.detach().cpu().tolist()
        if all(v == host[0] for v in host):
            t = int(host[0])
            return SequenceLayout([t] * (n // t if t else 0), lengths.device)
        return SequenceLayout([b - a for a, b in zip(host[:-1], host[1:])], lengths.device)

    def pad(self, x: torch.Tensor) -> torch.Tensor:
        if self.uniform:
            return x.view(-1, self.t, x.shape[1])
        out = x.new_zeros((self.n_slots, x.shape[1]))
        out.index_copy_(0, self.flat_index, x)
        return out.view(-1, self.t, x.shape[1])

    def unpad(self, h: torch.Tensor) -> torch.Tensor:
        flat = h.reshape(-1, h.shape[-1])
        return flat if self.uniform else flat.index_select(0, self.flat_index)


def sequence_layout(lengths: Optional[torch.Tensor], n: int) -> SequenceLayout:
    if lengths is None:
        return SequenceLayout.from_lengths(None, n)
    return _host_cache.get([lengths], ("seq", int(n)), lambda: SequenceLayout.from_lengths(lengths, n))


def batch_layout(batch: torch.Tensor) -> SequenceLayout:
    """Sequences = runs of equal graph id in a sorted PyG ``batch`` vector
    (``x.split(bincount(batch))``, analysisgnn/models/cadence.py:276-279)."""
    def make():
        counts = torch.bincount(batch).cpu().tolist() if batch.numel() else []
        return SequenceLayout([c for c in counts], batch.device)
    base = batch._base if batch._base is not None else batch
    return _host_cache.get([base], ("batch", int(batch.numel()), int(batch.storage_offset())), make)


def hetero_csr_trimmed(edge_index_dict, n_edges: Dict[Tuple[str, str, str], int], num_nodes: Dict[str, int]) -> "HeteroCSR":
    """``HeteroCSR`` of the first ``n_edges[et]`` edges of every type (what
    ``trim_to_layer`` leaves for a deeper layer), cached on the untrimmed tensors."""
    ets = list(edge_index_dict.keys())
    tensors = [edge_index_dict[et] for et in ets]
    extra = ("trim", tuple(ets), tuple(int(n_edges[et]) for et in ets), tuple(sorted(num_nodes.items())))
    return _cache.get(tensors, extra, lambda: HeteroCSR(
        {et: edge_index_dict[et][:, : int(n_edges[et])] for et in ets}, num_nodes))
