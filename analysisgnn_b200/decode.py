"""Onset-wise logit aggregation + decode on the GPU: drop-in for the reference's
``onsetwise_logit_aggregation(logits_softmax_dict, graph, edge_index_dict=None, batch_size=None,
valid_label_mask=None, rna_keys=[...])`` (analysisgnn/models/analysis.py:44-101; called from ``predict``, :1588).

Same arguments, same return value, same side effects (the onset mean is written into the caller's logit tensors,
:66; the entries of the Roman-numeral tasks are replaced in the dict, :69) and the same quirks (two softmaxes, the
last segment keeps its per-note distributions).  ``graph`` is anything that offers ``graph["note"].x / .batch /
.onset_div`` (attributes or keys) and ``graph.edge_index_dict`` -- a PyG ``HeteroData`` or a plain namespace.

Device work: ONE aggregation launch for all tasks (their logits are packed side by side, the onset CSR is shared;
the edge filters of :50-59 become relation id -1 in ``agnn_csr_build``), then per task ``agnn_softmax2_rows``,
``agnn_row_argmax``, ``agnn_run_heads``, ``agnn_decode_assign`` (include/agnn.h).  The Python loop over change
points with one mask over all notes per segment (:96-99) is a binary search per note.  Host synchronisations: the
ones the reference has as well (boolean-mask selections, ``torch.all(batch_id == batch_id[0])``) plus one flag read.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, graph as _graph, ops

RNA_KEYS = ("quality", "inversion", "degree1", "degree2")
ONSET = ("note", "onset", "note")


def _field(store, name):
    return store[name] if isinstance(store, dict) else getattr(store, name)


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _run_heads(keys: torch.Tensor, n: int, n_dev: Optional[torch.Tensor], unsorted: Optional[torch.Tensor]):
    lib = _lib.lib()
    dev = keys.device
    heads = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    n_runs = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.agnn_run_heads_workspace(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.agnn_run_heads(keys.data_ptr(), n, n_dev.data_ptr() if n_dev is not None else None, heads.data_ptr(),
                                  n_runs.data_ptr(), unsorted.data_ptr() if unsorted is not None else None,
                                  ws.data_ptr(), ws_bytes, _stream(keys)), "agnn_run_heads")
    _lib.count_launches(5)
    return heads, n_runs


@torch.no_grad()
def onsetwise_logit_aggregation(logits_softmax_dict, graph, edge_index_dict=None, batch_size=None,
                                valid_label_mask=None, rna_keys=RNA_KEYS):
    rna_keys = list(rna_keys)
    if not (rna_keys and all(k in logits_softmax_dict for k in rna_keys)):          # :45
        return logits_softmax_dict
    note = graph["note"]
    x = _field(note, "x")
    batch_size = len(x) if batch_size is None else int(batch_size)
    if edge_index_dict is None:
        edge_index_dict = _field(graph, "edge_index_dict") if isinstance(graph, dict) and "edge_index_dict" in graph \
            else graph.edge_index_dict
    first = logits_softmax_dict[rna_keys[0]]
    if not first.is_cuda:
        raise _lib.AgnnError("analysisgnn_b200 has no CPU path: tensors must live on a CUDA device")
    dev = first.device
    lib = _lib.lib()
    n_rows = first.shape[0]
    for k in rna_keys:
        v = logits_softmax_dict[k]
        if v.dim() != 2 or v.shape[0] != n_rows or v.dtype != torch.float32 or not v.is_cuda:
            raise ValueError("the Roman-numeral task logits must be fp32 CUDA matrices with one row per note")
    if valid_label_mask is None:
        mask, n_valid, rows_idx = None, batch_size, None
        if n_rows != batch_size:      # v[ones(batch_size)] in the reference: boolean index of the wrong length
            raise IndexError(f"logits have {n_rows} rows but the valid-label mask has {batch_size} entries")
    else:
        mask = valid_label_mask.to(dev)
        if mask.numel() != n_rows:
            raise IndexError(f"logits have {n_rows} rows but the valid-label mask has {mask.numel()} entries")
        rows_idx = mask.nonzero(as_tuple=True)[0].to(torch.int32)                   # sync, as boolean indexing is
        n_valid = int(rows_idx.numel())

    # ---- onset mean with the self term (:50-66): edge filters folded into the CSR build
    e = edge_index_dict[ONSET]
    tpc = None
    if "tpc_in_label" in logits_softmax_dict:
        tpc = logits_softmax_dict["tpc_in_label"].argmax(-1).bool()

    def make_etype():
        keep = (e[0] < batch_size) & (e[1] < batch_size) & (e[0] != e[1])
        if tpc is not None:
            last = tpc.numel() - 1
            keep = keep & tpc[e[0].clamp(max=last)] & tpc[e[1].clamp(max=last)]
        return keep.long() - 1                                                       # -1: dropped by agnn_csr_build

    etype = make_etype() if tpc is not None else _graph.derived(e, ("onset_decode", batch_size), make_etype)
    csr = _graph.typed_csr(e, etype, n_rows, 1, reduce_row=1)
    widths = [logits_softmax_dict[k].shape[1] for k in rna_keys]
    offs, total = [], 0
    for w in widths:
        offs.append(total)
        total += (w + 3) // 4 * 4
    packed = torch.zeros((n_rows, total), dtype=torch.float32, device=dev)
    for k, o, w in zip(rna_keys, offs, widths):
        packed[:, o:o + w].copy_(logits_softmax_dict[k])
    agg = ops.segment_mean_self(packed, packed, csr)
    out = {}
    for k, o, w in zip(rna_keys, offs, widths):
        logits_softmax_dict[k].copy_(agg[:, o:o + w])                               # scatter_mean(..., out=v)
        y = torch.empty((n_valid, w), dtype=torch.float32, device=dev)
        _lib.check(lib.agnn_softmax2_rows(agg.data_ptr() + 4 * o, agg.stride(0),
                                          rows_idx.data_ptr() if rows_idx is not None else None, n_valid, w,
                                          y.data_ptr(), y.stride(0) if n_valid else w, _stream(agg)), "agnn_softmax2_rows")
        _lib.count_launches(1)
        out[k] = y
    logits_softmax_dict.update(out)                                                  # :69

    # ---- single score: distributions held constant between arg-max change points (:70-99)
    batch_id = _field(note, "batch")[:batch_size]
    onsets = _field(note, "onset_div")[:batch_size]
    if mask is not None:
        batch_id, onsets = batch_id[mask], onsets[mask]
    if batch_id.numel() == 0:
        raise IndexError("index 0 is out of bounds for dimension 0 with size 0")     # batch_id[0] in the reference
    if not bool((batch_id == batch_id[0]).all()):
        return logits_softmax_dict
    onsets = (onsets - onsets.min()).to(torch.int64).contiguous()
    if tpc is not None:
        if tpc.numel() != onsets.numel():
            raise IndexError(f"the tpc_in_label mask has {tpc.numel()} entries for {onsets.numel()} notes")
        fidx = tpc.nonzero(as_tuple=True)[0].to(torch.int32)
        onsets_f = onsets[tpc].contiguous()
    else:
        fidx, onsets_f = None, onsets
    n_f = int(onsets_f.numel())
    if n_f == 0:
        raise IndexError("index 0 is out of bounds for dimension 0 with size 0")     # v[unique_logit_map] on no rows
    unsorted = torch.zeros(1, dtype=torch.int32, device=dev)
    onset_heads, n_unique = _run_heads(onsets_f, n_f, None, unsorted)
    for k in rna_keys:
        y = out[k]
        w = y.shape[1]
        pred = torch.empty(n_f, dtype=torch.int64, device=dev)
        _lib.check(lib.agnn_row_argmax(y.data_ptr(), y.stride(0), onset_heads.data_ptr(),
                                       fidx.data_ptr() if fidx is not None else None, n_unique.data_ptr(), n_f, w,
                                       pred.data_ptr(), _stream(y)), "agnn_row_argmax")
        cp_heads, n_cp = _run_heads(pred, n_f, n_unique, None)
        _lib.check(lib.agnn_decode_assign(y.data_ptr(), y.stride(0), w, onsets.data_ptr(), n_valid,
                                          onsets_f.data_ptr(), onset_heads.data_ptr(),
                                          fidx.data_ptr() if fidx is not None else None, cp_heads.data_ptr(),
                                          n_cp.data_ptr(), _stream(y)), "agnn_decode_assign")
        _lib.count_launches(2)
    if int(unsorted.item()):
        raise ValueError("onsetwise_logit_aggregation: note onsets must not decrease (the reference's unique / "
                         "change-point bookkeeping, analysis.py:79-95, indexes out of range otherwise)")
    return logits_softmax_dict
