"""Pins oracle/decode.py (onset-wise logit aggregation + decode) to the reference's own function
(analysisgnn/models/analysis.py:44-101): against the committed golden vectors
(tests/golden/decode_*.pt, made by tests/golden/make_golden.py) and, in the build container, against the
function executed live on further random cases."""
import glob
import os

import pytest
import torch

from analysisgnn_b200 import synth
from oracle import decode as odecode
from oracle import ref_loader

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "decode_*.pt")))


def run(fn, case):
    logits = {k: v.clone() for k, v in case["logits"].items()}
    originals = dict(logits)
    graph = odecode.note_store(case["x"], case["batch"], case["onset_div"], case["edge_index_dict"])
    out = fn(logits, graph, batch_size=case["batch_size"], valid_label_mask=case["valid_label_mask"])
    return out, originals


def same(a, b, what):
    assert set(a) == set(b), what
    for k in a:
        assert a[k].shape == b[k].shape, (what, k, a[k].shape, b[k].shape)
        assert torch.allclose(a[k], b[k], rtol=0, atol=2e-7), (what, k, float((a[k] - b[k]).abs().max()))


def test_golden_files_exist():
    assert len(GOLDEN) >= 6


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[7:-3] for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    rec = torch.load(path)
    out, originals = run(odecode.onsetwise_logit_aggregation, synth.decode_case(**rec["kwargs"]))
    same(out, rec["out"], "returned dict")
    same(originals, rec["mutated_inputs"], "caller's tensors after the in-place onset mean")


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("seed", range(8))
def test_oracle_matches_live_reference(seed):
    kw = dict(n_notes=50 + 37 * seed, seed=100 + seed, extra_nodes=(seed % 3) * 11, with_tpc=seed % 4 == 1,
              valid_fraction=0.7 if seed % 4 == 2 else 1.0, n_scores=2 if seed % 4 == 3 else 1, smooth=1 + seed)
    ref, ref_in = run(ref_loader.load_onsetwise_decode(), synth.decode_case(**kw))
    got, got_in = run(odecode.onsetwise_logit_aggregation, synth.decode_case(**kw))
    same(got, ref, "returned dict")
    same(got_in, ref_in, "mutated inputs")


def test_missing_rna_key_returns_the_dict_untouched():
    case = synth.decode_case(40, 9)
    case["logits"].pop("degree2")
    before = {k: v.clone() for k, v in case["logits"].items()}
    out, _ = run(odecode.onsetwise_logit_aggregation, case)
    same(out, before, "untouched")
