"""Parity of analysisgnn_b200.decode.onsetwise_logit_aggregation (CUDA, through libagnn's C ABI) with the
reference's function (analysisgnn/models/analysis.py:44-101): against the golden vectors the reference itself
produced (tests/golden/decode_*.pt), against the pinned CPU restatement (oracle/decode.py) on further seeded
cases, and through size-independent properties on a 200 k-note score.

Bars: the integer bookkeeping (onset runs, arg-max change points, which row every note takes) is exact, so a
decoded row is either a bit-for-bit copy of its segment's representative or its own distribution; the
distributions themselves are fp32 (mean over <= a chord's notes, two softmaxes): 1e-5 relative to the
largest probability (tests/util.py::FP32_REL), in practice ~1e-7."""
import glob
import os

import pytest
import torch

from analysisgnn_b200 import _lib, decode, synth
from oracle import decode as odecode
from tests.util import DEV, FP32_REL, assert_close

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "decode_*.pt")))


def to_dev(case):
    mv = lambda t: None if t is None else t.to(DEV)
    logits = {k: v.clone().to(DEV) for k, v in case["logits"].items()}
    graph = odecode.note_store(mv(case["x"]), mv(case["batch"]), mv(case["onset_div"]),
                               {k: mv(v) for k, v in case["edge_index_dict"].items()})
    return logits, graph, mv(case["valid_label_mask"])


def run_gpu(case):
    logits, graph, mask = to_dev(case)
    originals = dict(logits)
    out = decode.onsetwise_logit_aggregation(logits, graph, batch_size=case["batch_size"], valid_label_mask=mask)
    torch.cuda.synchronize()
    return out, originals


def run_cpu(case):
    logits = {k: v.clone() for k, v in case["logits"].items()}
    originals = dict(logits)
    graph = odecode.note_store(case["x"], case["batch"], case["onset_div"], case["edge_index_dict"])
    out = odecode.onsetwise_logit_aggregation(logits, graph, batch_size=case["batch_size"],
                                              valid_label_mask=case["valid_label_mask"])
    return out, originals


def same(got, want, what):
    assert set(got) == set(want), what
    for k in want:
        assert got[k].is_cuda, (what, k)
        assert_close(got[k], want[k], FP32_REL, f"{what}[{k}]")


def row_pattern(t):
    """Which rows are bit-for-bit copies of which: id of the first row equal to each row."""
    t = t.detach().cpu()
    _, inv = torch.unique(t, dim=0, return_inverse=True)
    first = torch.full((int(inv.max()) + 1,), t.shape[0], dtype=torch.int64)
    first.scatter_reduce_(0, inv, torch.arange(t.shape[0]), reduce="amin")
    return first[inv]


def run_starts(t, tol=1e-6):
    """Start index of the run of (nearly) equal consecutive rows each row belongs to.  Notes are in onset order, so a
    decoded segment is one run; so is a chord in the part that keeps per-note rows (the onset mean gives chord mates
    the same distribution up to the order of the fp32 sum, which is why the comparison is not bit-wise: WHICH chord
    mates come out bit-identical depends on that order).  Rows of different chords / segments differ by >= 1e-4."""
    t = t.detach().cpu()
    new = torch.ones(t.shape[0], dtype=torch.bool)
    if t.shape[0] > 1:
        new[1:] = (t[1:] - t[:-1]).abs().amax(-1) > tol
    idx = torch.arange(t.shape[0])
    return torch.cummax(torch.where(new, idx, torch.zeros_like(idx)), 0).values


def check_segments(got, want, onsets=None):
    """Same segmentation as the reference, and the rows of a decoded segment (a run spanning several onsets) are
    bit-for-bit copies of its first row."""
    for k in odecode.RNA_KEYS:
        g = got[k].detach().cpu()
        starts = run_starts(g)
        assert torch.equal(starts, run_starts(want[k])), f"{k}: segments differ from the reference's"
        if onsets is not None and g.shape[0] == onsets.numel():
            spans = onsets != onsets[starts]              # rows whose run began at another onset: decoded rows
            assert torch.equal(g[spans], g[starts[spans]]), f"{k}: a decoded row is not a copy of its segment's row"


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[7:-3] for p in GOLDEN])
def test_matches_reference_golden(path):
    rec = torch.load(path)
    out, originals = run_gpu(synth.decode_case(**rec["kwargs"]))
    same(out, rec["out"], "returned dict")
    same(originals, rec["mutated_inputs"], "caller's tensors after the in-place onset mean")
    kw = rec["kwargs"]
    full = kw.get("valid_fraction", 1.0) >= 1.0
    check_segments(out, rec["out"], synth.decode_case(**kw)["onset_div"][:kw["n_notes"]] if full else None)


@pytest.mark.parametrize("seed", range(10))
def test_matches_oracle_seeded(seed):
    kw = dict(n_notes=60 + 173 * seed, seed=200 + seed, extra_nodes=(seed % 3) * 17, with_tpc=seed % 4 == 1,
              valid_fraction=0.6 if seed % 4 == 2 else 1.0, n_scores=3 if seed % 4 == 3 else 1, smooth=1 + seed % 7)
    want, want_in = run_cpu(synth.decode_case(**kw))
    got, got_in = run_gpu(synth.decode_case(**kw))
    same(got, want, "returned dict")
    same(got_in, want_in, "mutated inputs")
    full = kw["valid_fraction"] >= 1.0
    check_segments(got, want, synth.decode_case(**kw)["onset_div"][:kw["n_notes"]] if full else None)


def test_valid_mask_together_with_several_scores_returns_before_the_decode():
    kw = dict(n_notes=400, seed=5, n_scores=2, valid_fraction=0.5)
    want, _ = run_cpu(synth.decode_case(**kw))
    got, _ = run_gpu(synth.decode_case(**kw))
    same(got, want, "returned dict")


def two_stage(kw):
    """The arg-max of a twice-softmaxed distribution is ill-conditioned (the second softmax squeezes a 1e-6 gap
    between the two best classes into one fp32 ulp), so on a long score two correct fp32 implementations of the mean +
    softmaxes (libm exp on the CPU, expf on the GPU) pick different change points somewhere.  Hence two stages:
    (A) the distributions BEFORE the hold stage -- obtained from the same call on a batch the reference does not
    treat as a single score -- match the oracle to fp32 tolerance; (B) given exactly those distributions, the hold
    stage (comparisons, arg-max, row copies: analysis.py:72-99) matches the oracle's BIT FOR BIT."""
    case = synth.decode_case(**kw)
    n = case["batch_size"]
    got, _ = run_gpu(case)
    split = synth.decode_case(**kw)
    split["batch"] = split["batch"].clone()
    split["batch"][0] += 1                  # two graph ids (note 0 always has a valid label): the reference returns at :71
    before_gpu, _ = run_gpu(split)
    before_cpu, _ = run_cpu(split)
    same(before_gpu, before_cpu, "distributions before the hold stage")
    mask = case["valid_label_mask"]
    onsets = case["onset_div"][:n] if mask is None else case["onset_div"][:n][mask]
    tpc = case["logits"]["tpc_in_label"].argmax(-1).bool() if "tpc_in_label" in case["logits"] else None
    want = odecode.hold_between_change_points({k: v.cpu().clone() for k, v in before_gpu.items()}, onsets, tpc)
    for k in odecode.RNA_KEYS:
        assert torch.equal(got[k].cpu(), want[k]), f"{k}: hold stage differs from the reference's on equal inputs"
    return got


def test_20k_note_score_two_stage():
    two_stage(dict(n_notes=20000, seed=11, smooth=9))


@pytest.mark.parametrize("seed", range(6))
def test_two_stage_seeded(seed):
    two_stage(dict(n_notes=300 + 411 * seed, seed=300 + seed, extra_nodes=(seed % 2) * 23, with_tpc=seed % 3 == 1,
                   valid_fraction=0.7 if seed % 3 == 2 else 1.0, smooth=2 + seed))


def test_missing_rna_key_returns_the_dict_untouched():
    case = synth.decode_case(40, 9)
    case["logits"].pop("degree2")
    before = {k: v.clone() for k, v in case["logits"].items()}
    out, _ = run_gpu(case)
    assert set(out) == set(before)
    for k in before:
        assert torch.equal(out[k].cpu(), before[k])


def test_one_note():
    want, _ = run_cpu(synth.decode_case(1, 3))
    got, _ = run_gpu(synth.decode_case(1, 3))
    same(got, want, "one note")


def test_decreasing_onsets_are_refused():
    case = synth.decode_case(50, 4)
    case["onset_div"] = case["onset_div"].flip(0).contiguous()
    with pytest.raises(ValueError):
        run_gpu(case)


def test_cpu_tensors_are_refused():
    case = synth.decode_case(30, 2)
    graph = odecode.note_store(case["x"], case["batch"], case["onset_div"], case["edge_index_dict"])
    with pytest.raises(_lib.AgnnError):
        decode.onsetwise_logit_aggregation(case["logits"], graph, batch_size=case["batch_size"])


def test_200k_note_score_properties():
    """Full-score size (BASELINE.json config 5): no oracle run (its loop is segments x notes); instead
    * every row sums to 1 and the decode is constant on each run of equal onsets outside the last segment;
    * a row either keeps softmax(softmax(onset mean)) or is a bit-for-bit copy of an EARLIER-or-equal onset's row;
    * running the call again on the un-decoded means gives bit-identical results (no atomics, no races)."""
    case = synth.decode_case(200_000, 21, smooth=12)
    out1, in1 = run_gpu(case)
    out2, _ = run_gpu(case)
    onsets = case["onset_div"]
    for k in odecode.RNA_KEYS:
        y = out1[k]
        assert torch.equal(y, out2[k]), f"{k}: not deterministic"
        assert float((y.sum(-1) - 1).abs().max()) < 1e-5
        own = in1[k].softmax(-1).softmax(-1)            # the caller's tensor now holds the onset mean
        yc, ownc = y.cpu(), own.cpu()
        kept = (yc - ownc).abs().amax(-1) < 1e-6      # own row (or a chord mate's: the onset mean makes them equal)
        pred = yc.argmax(-1)
        # the last segment keeps per-note rows; before it, every note of a run of equal onsets holds the same row
        changed = (~kept).nonzero(as_tuple=True)[0]
        assert changed.numel() > 0
        last_changed = int(changed.max())
        same_onset = onsets[1:last_changed + 1] == onsets[:last_changed]
        assert bool((yc[1:last_changed + 1][same_onset] == yc[:last_changed][same_onset]).all())
        # a segment is constant in its arg-max and the arg-max changes between consecutive segments
        first = row_pattern(yc[:last_changed + 1])
        assert bool((first[1:] >= first[:-1]).all()), "segments must be contiguous in onset order"
        heads = torch.cat([torch.zeros(1, dtype=torch.int64), (first[1:] != first[:-1]).nonzero(as_tuple=True)[0] + 1])
        assert bool((pred[heads][1:] != pred[heads][:-1]).all())
        assert bool((first <= torch.arange(first.numel())).all())
