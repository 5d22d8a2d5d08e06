"""On-disk corpus format (analysisgnn_b200/corpusfile.py): host-side code, so these run without a GPU.
Round trips are exact (integers and fp32 bits), damaged files are refused, and the converter from the reference's
collated ``processed/data.pt`` layout (PyG ``InMemoryDataset.collate``: concatenated stores + cumulative slices,
edge indices NOT incremented; analysisgnn/data/data_utils.py:53,115) reproduces the per-score graphs."""
import os

import numpy as np
import pytest
import torch

from analysisgnn_b200 import corpusfile as cf
from analysisgnn_b200 import synth


def make_scores(n_scores=5, seed=0):
    rng = np.random.default_rng(seed)
    xs, edges, spell, onset = [], [], [], []
    for s in range(n_scores):
        na = synth.synth_note_array(int(rng.integers(40, 160)), seed * 100 + s, voices=int(rng.integers(2, 6)))
        e = synth.score_graph_edges(na)                      # [3, E] local ids, the reference builder's order
        xs.append(rng.standard_normal((len(na), 7)).astype(np.float32))
        edges.append(e)
        spell.append(rng.integers(0, 35, len(na)))
        onset.append(na["onset_div"].astype(np.int64))
    return xs, edges, {"pitch_spelling": spell, "onset_div": onset}


def same_corpus(a, b):
    assert a.node_ptr == b.node_ptr and a.edge_ptr == b.edge_ptr and a.n_rel == b.n_rel
    assert torch.equal(a.x.cpu(), b.x.cpu()) and a.x.dtype == b.x.dtype
    assert torch.equal(a.edges.cpu(), b.edges.cpu()) and b.edges.dtype == torch.int64
    assert set(a.extras) == set(b.extras)
    for k in a.extras:
        assert torch.equal(a.extras[k].cpu(), b.extras[k].cpu()), k


def test_round_trip_is_exact(tmp_path):
    xs, edges, extras = make_scores()
    c = cf.corpus_from_scores(xs, edges, extras=extras)
    assert c.n_scores == 5 and c.edge_ptr[-1] == c.edges.shape[1]
    path = str(tmp_path / "corpus.agc")
    cf.save_corpus(path, c)
    same_corpus(c, cf.load_corpus(path, device="cpu"))
    tab = cf.read_table(path)
    assert all(e["offset"] % cf.ALIGN == 0 for e in tab["arrays"])
    assert os.path.getsize(path) % cf.ALIGN == 0
    by_name = {e["name"]: e for e in tab["arrays"]}
    assert by_name["edges"]["dtype"] == "int32" and by_name["node_ptr"]["dtype"] == "int64"
    # int32 indices: the edge array takes half the bytes of the reference's int64 edge_index + edge type
    assert by_name["edges"]["nbytes"] == 3 * 4 * c.edges.shape[1]


def test_empty_corpus_and_score_without_edges(tmp_path):
    path = str(tmp_path / "e.agc")
    c = cf.corpus_from_scores([np.zeros((3, 2), np.float32), np.zeros((0, 2), np.float32)],
                              [np.zeros((3, 0), np.int64), np.zeros((3, 0), np.int64)])
    cf.save_corpus(path, c)
    d = cf.load_corpus(path, device="cpu")
    assert d.node_ptr == [0, 3, 3] and d.edge_ptr == [0, 0, 0] and d.edges.shape == (3, 0)


@pytest.mark.parametrize("damage", ["magic", "version", "payload", "truncate", "table"])
def test_damaged_files_are_refused(tmp_path, damage):
    xs, edges, extras = make_scores(3, seed=1)
    path = str(tmp_path / "c.agc")
    cf.save_corpus(path, cf.corpus_from_scores(xs, edges, extras=extras))
    raw = bytearray(open(path, "rb").read())
    tab = cf.read_table(path)
    if damage == "magic":
        raw[0] ^= 0xFF
    elif damage == "version":
        raw[8] = 9
    elif damage == "payload":
        e = next(e for e in tab["arrays"] if e["name"] == "x")
        raw[e["offset"] + 5] ^= 0x01
    elif damage == "truncate":
        raw = raw[: len(raw) - cf.ALIGN]
    elif damage == "table":
        raw[20] = 0xFF
    open(path, "wb").write(bytes(raw))
    with pytest.raises(cf.CorpusFormatError):
        cf.load_corpus(path, device="cpu")


def test_invariants_are_checked_on_save(tmp_path):
    xs, edges, _ = make_scores(2, seed=2)
    c = cf.corpus_from_scores(xs, edges)
    path = str(tmp_path / "c.agc")
    bad = c.edges.clone()
    bad[1, 0] = c.node_ptr[1]                   # destination in the next score
    c.edges = bad
    with pytest.raises(cf.CorpusFormatError, match="two different scores"):
        cf.save_corpus(path, c)
    c = cf.corpus_from_scores(xs, edges)
    c.edges = c.edges.flip(1)                   # scores out of order
    with pytest.raises(cf.CorpusFormatError, match="grouped by score"):
        cf.save_corpus(path, c)
    c = cf.corpus_from_scores(xs, edges)
    c.edges[2, 0] = 9
    with pytest.raises(cf.CorpusFormatError, match="relation ids"):
        cf.save_corpus(path, c)


def collate_like_pyg(xs, edges, extras, rel_names):
    """What ``InMemoryDataset.collate`` stores: concatenations + cumulative slices, local edge indices."""
    n_ptr = torch.tensor(np.concatenate(([0], np.cumsum([len(x) for x in xs]))))
    data = {"note": {"x": torch.cat([torch.from_numpy(x) for x in xs])}}
    slices = {"note": {"x": n_ptr}}
    for k, vs in extras.items():
        data["note"][k] = torch.cat([torch.as_tensor(v) for v in vs])
        slices["note"][k] = n_ptr
    data["note"]["name"] = ["a"] * len(xs)        # non-tensor attributes exist in the real files
    for r, name in enumerate(rel_names):
        per = [torch.as_tensor(e[:2, e[2] == r]) for e in edges]
        data[("note", name, "note")] = {"edge_index": torch.cat(per, dim=1)}
        slices[("note", name, "note")] = {"edge_index": torch.tensor(np.concatenate(([0], np.cumsum([p.shape[1] for p in per]))))}
    return data, slices


def test_converter_from_the_reference_collation(tmp_path):
    xs, edges, extras = make_scores(4, seed=3)
    data, slices = collate_like_pyg(xs, edges, extras, cf.REL_NAMES)
    c = cf.from_pyg_collated(data, slices)
    assert set(c.extras) == {"pitch_spelling", "onset_div"}
    assert c.node_ptr == [0] + list(np.cumsum([len(x) for x in xs]))
    for s in range(4):
        lo, hi = c.edge_ptr[s], c.edge_ptr[s + 1]
        got = c.edges[:, lo:hi].clone()
        got[:2] -= c.node_ptr[s]
        want = torch.as_tensor(edges[s])
        # same edge multiset per relation, grouped by relation in the file (the collation groups by relation too)
        key = lambda e: sorted(map(tuple, e.t().tolist()))
        assert key(got) == key(want), s
    path = str(tmp_path / "ref.agc")
    cf.save_corpus(path, c)
    same_corpus(c, cf.load_corpus(path, device="cpu"))


def test_loader_reads_a_loaded_corpus_like_the_original(tmp_path):
    """The windows the loader cuts depend only on (seed, score sizes): a corpus that went through the file gives the
    same window starts and score order."""
    from analysisgnn_b200 import sampler
    xs, edges, extras = make_scores(6, seed=4)
    c = cf.corpus_from_scores(xs, edges, extras=extras)
    path = str(tmp_path / "c.agc")
    cf.save_corpus(path, c)
    d = cf.load_corpus(path, device="cpu")
    la = sampler.ScoreGraphLoader(c, 32, 3, seed=7)
    lb = sampler.ScoreGraphLoader(d, 32, 3, seed=7)
    assert la.order(2) == lb.order(2) and len(la) == len(lb)


def test_convert_cli(tmp_path):
    """tools/convert_corpus.py on a file laid out like the reference's processed/data.pt (plain dicts pickle without PyG)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    xs, edges, extras = make_scores(3, seed=9)
    data, slices = collate_like_pyg(xs, edges, extras, cf.REL_NAMES)
    src, dst = str(tmp_path / "data.pt"), str(tmp_path / "c.agc")
    torch.save((data, slices, dict), src)
    run = subprocess.run([sys.executable, os.path.join(root, "tools", "convert_corpus.py"), src, dst],
                         capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    same_corpus(cf.from_pyg_collated(data, slices), cf.load_corpus(dst, device="cpu"))
    info = subprocess.run([sys.executable, os.path.join(root, "tools", "convert_corpus.py"), "--info", dst],
                          capture_output=True, text=True)
    assert info.returncode == 0 and "checksums ok" in info.stdout and "extra.onset_div" in info.stdout


def test_container_round_trips_arbitrary_arrays(tmp_path):
    """The container itself (save_arrays / load_arrays): every storable dtype, empty and odd shapes, names with dots --
    bytes come back exactly, every array on a page boundary (hypothesis-driven)."""
    from hypothesis import given, settings, strategies as st
    from hypothesis.extra import numpy as hnp

    dtypes = st.sampled_from([np.int32, np.int64, np.float32, np.float64, np.uint8])
    arrays = st.dictionaries(
        st.text(alphabet="abcxyz._0189", min_size=1, max_size=12),
        dtypes.flatmap(lambda d: hnp.arrays(d, hnp.array_shapes(min_dims=1, max_dims=3, min_side=0, max_side=9))),
        min_size=1, max_size=5)
    counter = {"n": 0}

    @settings(max_examples=40, deadline=None)
    @given(arrays, st.integers(1, 9))
    def check(arrs, n_rel):
        counter["n"] += 1
        path = str(tmp_path / f"h{counter['n']}.agc")
        cf.save_arrays(path, arrs, n_rel)
        back, rel = cf.load_arrays(path, verify=True)
        assert rel == n_rel and set(back) == set(arrs)
        for k, a in arrs.items():
            assert back[k].dtype == a.dtype and back[k].shape == a.shape
            assert np.asarray(back[k]).tobytes() == np.ascontiguousarray(a).tobytes()
        assert all(e["offset"] % cf.ALIGN == 0 for e in cf.read_table(path)["arrays"])

    check()
