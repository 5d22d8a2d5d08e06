"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

DEV = os.environ.get("AGNN_TEST_DEVICE", "cuda:0")   # the parity tests proper run on the GPU
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FP32_REL = 1e-5   # BASELINE.json north_star: "within 1e-5 relative for fp32 forward outputs and gradients"
BF16_REL = 2e-2   # "... and 2e-2 for a stated bf16 mode"


def rel_err(got: torch.Tensor, want: torch.Tensor) -> float:
    """max |got - want| / max |want|  (relative to the tensor's scale; a per-element
    ratio is meaningless where the reference value itself is a rounding residue)."""
    got = got.detach().double().cpu()
    want = want.detach().double().cpu()
    assert got.shape == want.shape, (got.shape, want.shape)
    if want.numel() == 0:
        return 0.0
    scale = max(float(want.abs().max()), 1e-30)
    return float((got - want).abs().max()) / scale


def assert_close(got, want, tol, what=""):
    err = rel_err(got, want)
    assert err <= tol, f"{what}: relative error {err:.3e} > {tol:.1e}"
    return err


def golden_edges(name):
    return np.load(os.path.join(GOLDEN, f"edges_{name}.npz"))


def golden_intree(seed):
    return torch.load(os.path.join(GOLDEN, f"intree_seed{seed}.pt"), weights_only=False)


EDGE_CASES = ["hand12", "synth_s0_n60_v4", "synth_s1_n97_v2", "synth_s2_n120_v8", "synth_s3_n500_v4"]


def probe_loss(out: torch.Tensor) -> torch.Tensor:
    """The fixed scalar the goldens differentiate: sum(out * linspace(0.25, 1.25)) -- one-signed weights, so bias gradients (plain column sums) do not cancel to rounding noise."""
    w = torch.linspace(0.25, 1.25, out.numel(), dtype=torch.float32).view_as(out).to(out.device, out.dtype)
    return (out * w).sum()


def grads_of(module, out, inputs):
    loss = probe_loss(out)
    named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
    got = torch.autograd.grad(loss, [p for _, p in named] + list(inputs), allow_unused=True)
    pg = {n: g for (n, _), g in zip(named, got[:len(named)]) if g is not None}
    return pg, list(got[len(named):])


# ---------------------------------------------------------------------------------------------
# ReLU-pattern bookkeeping.  A gradient is a discontinuous function of the inputs wherever a
# pre-activation crosses zero: two correct fp32 implementations that sum in a different order can
# put a |z| ~ 1e-7 pre-activation on opposite sides, and on a few-hundred-node test graph that one
# flipped unit moves a weight gradient by ~1e-3 of its scale.  The parity claim "gradients within
# 1e-5" is therefore stated -- and tested -- on inputs where both implementations produce the same
# activation pattern; the helpers below record the patterns so a test can tell the two cases apart
# and move to the next seed instead of asserting on a flipped unit.
# ---------------------------------------------------------------------------------------------

class ActivationPatterns:
    """Records ``output > 0`` of every module selected by ``pick(name, module)`` (tensor or dict
    of tensors), keyed by module name (the two implementations may run branches in a different
    order) and call count."""

    def __init__(self, model, pick):
        self.masks = {}
        self.handles = [m.register_forward_hook(self._hook(n)) for n, m in model.named_modules() if pick(n, m)]

    def _hook(self, name):
        def hook(module, args, out):
            outs = out if isinstance(out, dict) else {"": out[0] if isinstance(out, tuple) else out}
            self.masks.setdefault(name, []).append({k: (o.detach() > 0).cpu() for k, o in outs.items()})
        return hook

    def close(self):
        for h in self.handles:
            h.remove()

    def mismatches(self, other: "ActivationPatterns") -> int:
        assert set(self.masks) == set(other.masks), set(self.masks) ^ set(other.masks)
        total = 0
        for name, calls in self.masks.items():
            assert len(calls) == len(other.masks[name]), name
            for a, b in zip(calls, other.masks[name]):
                # dict outputs: a destination type one side skipped as unused (last layer) is not compared
                total += sum(int((a[k] != b[k]).sum()) for k in a if k in b)
        return total


def feeds_relu(name, module):
    """Modules whose output goes straight into a ReLU in the encoders under test."""
    cls = type(module).__name__
    if cls in ("ReLU", "HeteroSAGELayer", "HGTConv", "HeteroConv"):
        return True
    return cls == "Linear" and (name.endswith("conv_out") or ".project_metrical." in name
                                or name.startswith("project_metrical."))


MAX_SKIPPED_SEEDS = 2


def first_seed_with_equal_patterns(run, seeds=(0, 1, 2, 3, 4, 5)):
    """``run(seed)`` -> (mismatches, payload).  Returns the payload of the first seed on which the CPU oracle and the
    CUDA path agree on every activation sign (``run`` asserts the forward outputs on EVERY seed it sees, and the
    gradients on that one).  Seeds skipped because a |z| ~ 1e-7 unit landed on the other side are counted, printed,
    and more than ``MAX_SKIPPED_SEEDS`` of them fail the test: a real bug that flips activations on most seeds must
    not pass because one seed happens to agree.  (The full-size tests, tests/test_fullsize_gpu.py, do not skip at
    all: they evaluate the oracle's gradient with the CUDA side's sign choice at the flipped units.)"""
    seen = []
    for seed in seeds[:MAX_SKIPPED_SEEDS + 1]:
        mism, payload = run(seed)
        seen.append(mism)
        if mism == 0:
            print(f"[relu patterns] gradients asserted on seed {seed}; skipped {len(seen) - 1} seed(s) with flipped "
                  f"units: {seen[:-1]}")
            return payload
    raise AssertionError(f"activation patterns differed on more than {MAX_SKIPPED_SEEDS} seeds in a row "
                         f"(flipped units per seed: {seen})")
