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
