"""Score-graph construction on the GPU (agnn_score_graph_build).

Drop-in for the reference's ``hetero_graph_from_note_array`` (analysisgnn/utils/hgraph.py:214-300):
same edge set, same edge order, same integer codes (onset=0, consecutive=1, during=2, rest=3), for one
score or a whole batch of scores at once (node ids offset per score, as ``batch_graphs`` /
``Batch.from_data_list`` would collate them).
"""
from __future__ import annotations

from typing import Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib


def _host_plan(note_arrays: Sequence[np.ndarray]):
    score_ptr, key_base = [0], [0]
    onset, dur = [], []
    for na in note_arrays:
        o = np.asarray(na["onset_div"], dtype=np.int64)
        d = np.asarray(na["duration_div"], dtype=np.int64)
        if len(o) and np.any(np.diff(o) < 0):
            raise ValueError("note_array must be sorted by onset_div")
        span = int((o + d).max() - o[0]) + 1 if len(o) else 1
        score_ptr.append(score_ptr[-1] + len(o))
        key_base.append(key_base[-1] + span)
        onset.append(o)
        dur.append(d)
    onset = np.concatenate(onset) if onset else np.zeros(0, np.int64)
    dur = np.concatenate(dur) if dur else np.zeros(0, np.int64)
    if len(onset) and (np.abs(onset).max() >= 2 ** 30 or dur.max() >= 2 ** 30 or key_base[-1] >= 2 ** 31 - 1):
        raise ValueError("onset / duration values exceed the int32 range of the builder")
    return (np.asarray(score_ptr, np.int32), np.asarray(key_base, np.int32), onset.astype(np.int32),
            dur.astype(np.int32))


def score_graph_edges(note_arrays: Union[np.ndarray, Sequence[np.ndarray]], device="cuda") -> Tuple[torch.Tensor, torch.Tensor]:
    """(edges int64 [3, E] on ``device``, score_ptr int32 [S+1]) for one note array or a list of them."""
    if isinstance(note_arrays, np.ndarray):
        note_arrays = [note_arrays]
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.AgnnError("analysisgnn_b200 has no CPU path: the score-graph builder needs a CUDA device")
    score_ptr, key_base, onset, dur = _host_plan(note_arrays)
    n, s, slots = len(onset), len(note_arrays), int(key_base[-1])
    lib = _lib.lib()
    to = lambda a: torch.from_numpy(a).to(device, non_blocking=True)
    d_ptr, d_base, d_on, d_dur = to(score_ptr), to(key_base), to(onset), to(dur)
    ws_bytes = lib.agnn_score_graph_workspace(n, s, slots)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=device)
    n_edges = torch.zeros(1, dtype=torch.int32, device=device)
    stream = torch.cuda.current_stream(device).cuda_stream
    capacity = 16 * n + 1024
    while True:
        edges = torch.empty((3, capacity), dtype=torch.int64, device=device)
        _lib.check(lib.agnn_score_graph_build(s, d_ptr.data_ptr(), d_base.data_ptr(), d_on.data_ptr(), d_dur.data_ptr(),
                                              n, slots, edges.data_ptr(), capacity, n_edges.data_ptr(), ws.data_ptr(),
                                              ws_bytes, stream), "agnn_score_graph_build")
        _lib.count_launches(12)
        total = int(n_edges.item())
        if total <= capacity:
            return edges[:, :total], d_ptr
        capacity = total


def hetero_graph_from_note_array(note_array, rest_array=None, norm2bar=False, pot_edge_dist=0, device="cuda"):
    """Signature of the reference builder (hgraph.py:214); returns ``(nodes, edges)`` with ``edges`` a
    numpy int64 ``[3, E]`` array exactly as the reference produces it."""
    if rest_array is not None or pot_edge_dist:
        raise NotImplementedError("rest_array / pot_edge_dist are not part of the accelerated path")
    if norm2bar:
        note_array = note_array.copy()
        note_array["onset_beat"] = np.mod(note_array["onset_beat"], note_array["ts_beats"])
    edges, _ = score_graph_edges(note_array, device)
    return note_array, edges.cpu().numpy()
