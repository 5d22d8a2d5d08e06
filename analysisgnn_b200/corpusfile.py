"""On-disk corpus format for the GPU loader.

The reference keeps a dataset as ONE pickled PyG ``InMemoryDataset`` file, ``processed/data.pt`` =
``(collated HeteroData.to_dict(), slices, HeteroData)`` (analysisgnn/data/data_utils.py:53, 79-80, 115;
analysisgnn/data/datasets/dlc.py:341-342, 434): int64 indices, ``torch.load`` unpickles every tensor into
pageable memory, and a score is cut out of the collation in Python on every access.  The device-resident
``sampler.Corpus`` needs the same content as a few flat arrays, so this module stores exactly those:

    offset 0      magic  b"AGNNCORP"
    8             u32 version (= 1), u32 length of the JSON table
    16            JSON table: {"n_rel": R, "arrays": [{"name", "dtype", "shape", "offset", "nbytes", "crc32"}, ...]}
    4096-aligned  the arrays, little endian, C order, each starting on a 4096-byte boundary

    node_ptr  int64 [S+1]   first note of every score (global note ids)
    x         float32 [N,F] note features
    edges     int32 [3,E]   source, destination (global note ids), relation id; the edges of a score are
                            contiguous and scores appear in node_ptr order
    extra.<k> any [N,...]   per-note labels / spellings / onsets (int64 on disk stays int64; int32 is widened)

Indices are int32 on disk (half the bytes of the reference's int64) and widened on the device, where
``agnn_window_subgraph`` / ``agnn_csr_build`` take int64.  Reading is one ``np.memmap`` per file: every array is
copied page-aligned into a pinned staging buffer and sent with one asynchronous host->device copy; nothing is
unpickled, and no per-score Python runs.  ``from_pyg_collated`` converts the reference's collated dict (run it where
PyG is installed: ``data, slices, _ = torch.load("processed/data.pt")``); it needs only ``dict``/``Tensor`` access.

Host-side code only: no kernel is involved, so this module works without a GPU (``device="cpu"``) and is covered by
the CPU test suite.
"""
from __future__ import annotations

import json
import struct
import zlib
from typing import Dict, Mapping, Optional, Sequence

import numpy as np
import torch

from .sampler import Corpus

MAGIC = b"AGNNCORP"
VERSION = 1
ALIGN = 4096
REL_NAMES = ("onset", "consecutive", "during", "rest")     # analysisgnn/utils/hgraph.py:229-285

_DTYPES = {"int32": np.int32, "int64": np.int64, "float32": np.float32, "uint8": np.uint8, "float64": np.float64}


class CorpusFormatError(ValueError):
    pass


def _align(n: int) -> int:
    return (n + ALIGN - 1) // ALIGN * ALIGN


def _as_numpy(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().contiguous().numpy()
    return np.ascontiguousarray(t)


def save_arrays(path: str, arrays: Mapping[str, np.ndarray], n_rel: int) -> None:
    """Write the container: header, JSON table, page-aligned arrays with a CRC-32 each."""
    entries, blobs = [], []
    for name, a in arrays.items():
        a = _as_numpy(a)
        if a.dtype.name not in _DTYPES:
            raise CorpusFormatError(f"{name}: dtype {a.dtype} is not storable ({sorted(_DTYPES)})")
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        entries.append({"name": name, "dtype": a.dtype.name, "shape": list(a.shape), "offset": 0, "nbytes": int(a.nbytes),
                        "crc32": zlib.crc32(a.tobytes()) & 0xFFFFFFFF})
        blobs.append(a)

    def table(es):
        return json.dumps({"n_rel": int(n_rel), "arrays": es}, separators=(",", ":")).encode()

    # offsets depend on the table length, the table length on the offsets' digits: fixed point in <= 3 passes
    for _ in range(4):
        pos = _align(16 + len(table(entries)))
        changed = False
        for e in entries:
            if e["offset"] != pos:
                e["offset"], changed = pos, True
            pos = _align(pos + e["nbytes"])
        if not changed:
            break
    tab = table(entries)
    with open(path, "wb") as fh:
        fh.write(MAGIC)
        fh.write(struct.pack("<II", VERSION, len(tab)))
        fh.write(tab)
        for e, a in zip(entries, blobs):
            fh.write(b"\0" * (e["offset"] - fh.tell()))
            fh.write(a.tobytes())
        fh.write(b"\0" * (_align(fh.tell()) - fh.tell()))


def read_table(path: str) -> dict:
    with open(path, "rb") as fh:
        head = fh.read(16)
        if len(head) < 16 or head[:8] != MAGIC:
            raise CorpusFormatError(f"{path}: not an analysisgnn_b200 corpus file (bad magic)")
        version, n = struct.unpack("<II", head[8:])
        if version != VERSION:
            raise CorpusFormatError(f"{path}: format version {version}, this reader understands {VERSION}")
        raw = fh.read(n)
        if len(raw) != n:
            raise CorpusFormatError(f"{path}: truncated header")
        fh.seek(0, 2)
        size = fh.tell()
    try:
        tab = json.loads(raw.decode())
    except (UnicodeDecodeError, json.JSONDecodeError) as exc:
        raise CorpusFormatError(f"{path}: unreadable array table ({exc})") from None
    for e in tab["arrays"]:
        if e["dtype"] not in _DTYPES or e["offset"] % ALIGN or e["offset"] + e["nbytes"] > size:
            raise CorpusFormatError(f"{path}: array {e['name']!r} lies outside the file or is misaligned")
        if int(np.prod(e["shape"], dtype=np.int64)) * np.dtype(_DTYPES[e["dtype"]]).itemsize != e["nbytes"]:
            raise CorpusFormatError(f"{path}: array {e['name']!r}: shape and byte count disagree")
    return tab


def load_arrays(path: str, verify: bool = True) -> (Dict[str, np.ndarray], int):
    """Memory-map every array (read only, zero copy).  ``verify`` checks the CRC-32 of each (one pass over the file)."""
    tab = read_table(path)
    out = {}
    for e in tab["arrays"]:
        a = np.memmap(path, dtype=_DTYPES[e["dtype"]], mode="r", offset=e["offset"], shape=tuple(e["shape"])) \
            if e["nbytes"] else np.zeros(tuple(e["shape"]), dtype=_DTYPES[e["dtype"]])
        if verify and (zlib.crc32(a.tobytes()) & 0xFFFFFFFF) != e["crc32"]:
            raise CorpusFormatError(f"{path}: array {e['name']!r} fails its checksum")
        out[e["name"]] = a
    return out, int(tab["n_rel"])


# ------------------------------------------------------------------------------------------ Corpus <-> file

def check_corpus_arrays(node_ptr: np.ndarray, edges: np.ndarray, n_notes: int, n_rel: int) -> None:
    """The invariants ``agnn_window_subgraph`` relies on (sampler.window_subgraphs): scores tile the notes, every
    edge stays inside one score, edges are grouped by score in score order, relation ids are in range."""
    if node_ptr.ndim != 1 or node_ptr.size < 1 or node_ptr[0] != 0 or node_ptr[-1] != n_notes or \
            (np.diff(node_ptr) < 0).any():
        raise CorpusFormatError("node_ptr must rise from 0 to the number of notes")
    if edges.ndim != 2 or edges.shape[0] != 3:
        raise CorpusFormatError("edges must be [3, E] (source, destination, relation)")
    if edges.shape[1] == 0:
        return
    if edges[:2].min() < 0 or edges[:2].max() >= n_notes:
        raise CorpusFormatError("edge endpoints outside [0, notes)")
    if edges[2].min() < 0 or edges[2].max() >= n_rel:
        raise CorpusFormatError("relation ids outside [0, n_rel)")
    s_src = np.searchsorted(node_ptr[1:], edges[0], side="right")
    s_dst = np.searchsorted(node_ptr[1:], edges[1], side="right")
    if (s_src != s_dst).any():
        raise CorpusFormatError("an edge joins two different scores")
    if (np.diff(s_src) < 0).any():
        raise CorpusFormatError("edges must be grouped by score, in score order")


def save_corpus(path: str, corpus: Corpus) -> None:
    """``sampler.Corpus`` (device or host tensors) -> file."""
    n = corpus.node_ptr[-1]
    if n >= 2 ** 31:
        raise CorpusFormatError("more than 2^31 - 1 notes do not fit the int32 indices of format version 1")
    node_ptr = np.asarray(corpus.node_ptr, dtype=np.int64)
    edges = _as_numpy(corpus.edges).astype(np.int32)
    check_corpus_arrays(node_ptr, edges, n, corpus.n_rel)
    x = _as_numpy(corpus.x)
    if x.dtype != np.float32 or x.ndim != 2 or x.shape[0] != n:
        raise CorpusFormatError("x must be float32 [notes, features]")
    arrays = {"node_ptr": node_ptr, "x": x, "edges": edges}
    for k, v in corpus.extras.items():
        v = _as_numpy(v)
        if v.shape[0] != n:
            raise CorpusFormatError(f"extra {k!r} must have one row per note")
        arrays["extra." + k] = v
    save_arrays(path, arrays, corpus.n_rel)


def _to_device(a: np.ndarray, device, pin: bool, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """memmap -> (pinned) host tensor -> device, widening int32 indices on the device."""
    device = torch.device(device)
    host = torch.empty(a.shape, dtype=torch.from_numpy(np.zeros(0, dtype=a.dtype)).dtype,
                       pin_memory=pin and device.type == "cuda")
    np.copyto(host.numpy(), a)                       # the one pass over the mapped pages
    t = host.to(device, non_blocking=True) if device.type == "cuda" else host
    return t.to(dtype) if dtype is not None and t.dtype != dtype else t


def load_corpus(path: str, device="cuda", verify: bool = True, pin: bool = True) -> Corpus:
    arrays, n_rel = load_arrays(path, verify=verify)
    for need in ("node_ptr", "x", "edges"):
        if need not in arrays:
            raise CorpusFormatError(f"{path}: array {need!r} is missing")
    node_ptr = np.asarray(arrays["node_ptr"], dtype=np.int64)
    if verify:
        check_corpus_arrays(node_ptr, np.asarray(arrays["edges"]), int(arrays["x"].shape[0]), n_rel)
    x = _to_device(arrays["x"], device, pin)
    edges = _to_device(arrays["edges"], device, pin, torch.int64)
    extras = {k[6:]: _to_device(v, device, pin, torch.int64 if v.dtype == np.int32 else None)
              for k, v in arrays.items() if k.startswith("extra.")}
    if torch.device(device).type == "cuda":
        torch.cuda.current_stream(torch.device(device)).synchronize()    # the pinned staging buffers die here
    return Corpus(x, edges, node_ptr.tolist(), n_rel=n_rel, extras=extras)


# ------------------------------------------------------------------------------------------ converters

def corpus_from_scores(xs: Sequence, edge_lists: Sequence, n_rel: int = 4,
                       extras: Optional[Mapping[str, Sequence]] = None, device="cpu") -> Corpus:
    """Per-score arrays -> Corpus: ``xs[s]`` float [N_s, F]; ``edge_lists[s]`` int [3, E_s] with LOCAL note ids, as
    ``hetero_graph_from_note_array`` returns them (analysisgnn/utils/hgraph.py:300); ``extras[k][s]`` [N_s, ...]."""
    if len(xs) != len(edge_lists):
        raise ValueError("one edge list per score")
    node_ptr = np.concatenate(([0], np.cumsum([len(x) for x in xs]))).astype(np.int64)
    parts = []
    for s, e in enumerate(edge_lists):
        e = torch.as_tensor(_as_numpy(e)).to(torch.int64).reshape(3, -1).clone()
        if e.numel() and (int(e[:2].min()) < 0 or int(e[:2].max()) >= len(xs[s])):
            raise ValueError(f"score {s}: edge endpoints outside the score")
        e[:2] += int(node_ptr[s])
        parts.append(e)
    edges = torch.cat(parts, dim=1) if parts else torch.zeros((3, 0), dtype=torch.int64)
    x = torch.cat([torch.as_tensor(_as_numpy(v), dtype=torch.float32) for v in xs]) if len(xs) else torch.zeros((0, 0))
    ex = {k: torch.cat([torch.as_tensor(_as_numpy(v)) for v in vs]).to(device) for k, vs in (extras or {}).items()}
    return Corpus(x.to(device), edges.to(device), node_ptr.tolist(), n_rel=n_rel, extras=ex)


def from_pyg_collated(data: Mapping, slices: Mapping, rel_names: Sequence[str] = REL_NAMES, node_type: str = "note",
                      feature_key: str = "x", extra_keys: Optional[Sequence[str]] = None, device="cpu") -> Corpus:
    """The reference's ``processed/data.pt`` content -> Corpus.

    ``data`` / ``slices`` are the first two items of the saved tuple: ``data[node_type][key]`` is the concatenation
    over scores and ``slices[node_type][key]`` its cumulative row pointer; ``data[(node_type, rel, node_type)]
    ["edge_index"]`` is ``[2, sum E]`` with LOCAL note ids per score (``InMemoryDataset.collate`` does not increment
    indices) and ``slices[...]["edge_index"]`` the cumulative edge pointer.  Relations the file does not hold are
    empty; other node / edge types (beats, measures) are rebuilt per batch from the note fields and are not stored."""
    store, sl = data[node_type], slices[node_type]
    node_ptr = torch.as_tensor(sl[feature_key]).to(torch.int64)
    n_scores = node_ptr.numel() - 1
    x = torch.as_tensor(store[feature_key]).to(torch.float32)
    per_score = [[] for _ in range(n_scores)]
    for r, name in enumerate(rel_names):
        key = (node_type, name, node_type)
        if key not in data:
            continue
        ei = torch.as_tensor(data[key]["edge_index"]).to(torch.int64)
        ptr = torch.as_tensor(slices[key]["edge_index"]).to(torch.int64).tolist()
        for s in range(n_scores):
            e = ei[:, ptr[s]:ptr[s + 1]]
            per_score[s].append(torch.cat((e + node_ptr[s], torch.full((1, e.shape[1]), r, dtype=torch.int64))))
    flat = [torch.cat(p, dim=1) if p else torch.zeros((3, 0), dtype=torch.int64) for p in per_score]
    edges = torch.cat(flat, dim=1) if flat else torch.zeros((3, 0), dtype=torch.int64)
    if extra_keys is None:
        extra_keys = [k for k, v in store.items() if k != feature_key and isinstance(v, torch.Tensor)
                      and v.dim() >= 1 and v.shape[0] == x.shape[0]]
    extras = {k: torch.as_tensor(store[k]).to(device) for k in extra_keys}
    return Corpus(x.to(device), edges.to(device), node_ptr.tolist(), n_rel=len(rel_names), extras=extras)
