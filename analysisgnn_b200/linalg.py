"""Dense contractions used by the fused layers.

The per-relation / per-node-type projections are the only GEMM-shaped work on
the hot path.  ``backend()`` selects who runs them:

* ``"cublas"``  -- ``torch.addmm`` / ``torch.mm`` (library GEMM, fp32 without TF32).
* ``"tcgen05"`` -- this repo's sm_100a tcgen05/TMEM grouped GEMM (csrc/gemm.cu).

There is no CPU path in either case: inputs must be CUDA tensors.
"""
from __future__ import annotations

import os

import torch

_BACKEND = os.environ.get("AGNN_GEMM", "cublas")


def backend() -> str:
    return _BACKEND


def set_backend(name: str) -> None:
    global _BACKEND
    if name not in ("cublas", "tcgen05"):
        raise ValueError(name)
    _BACKEND = name


def _need_cuda(t):
    if not t.is_cuda:
        from ._lib import AgnnError
        raise AgnnError("analysisgnn_b200 has no CPU path: tensors must live on a CUDA device")


def linear(x, weight, bias=None, relu: bool = False):
    """``x @ weight.T + bias`` (optionally ReLU'd), ``weight`` [out, in]."""
    _need_cuda(x)
    if bias is not None:
        y = torch.addmm(bias, x, weight.t())
    else:
        y = torch.mm(x, weight.t())
    return y.relu_() if relu else y


def mm(a, b, out=None, accumulate: bool = False):
    """``a @ b``; with ``out`` and ``accumulate`` adds into ``out`` in place."""
    _need_cuda(a)
    if out is None:
        return torch.mm(a, b)
    if accumulate:
        return out.addmm_(a, b)
    return torch.mm(a, b, out=out)


def mm_tn(a, b):
    """``a.T @ b`` (weight gradients)."""
    _need_cuda(a)
    return torch.mm(a.t(), b)


def relu_backward(grad, out):
    """Gradient through ``out = relu(.)`` given the saved output."""
    return torch.where(out > 0, grad, torch.zeros((), dtype=grad.dtype, device=grad.device))
