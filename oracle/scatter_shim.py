"""Pure-torch stand-in for the ``torch_scatter`` entry points the reference uses.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Semantics restated from torch_scatter 2.1 (third-party, not installable here;
pinned by the reference only as ``torch-scatter>=2.1.0``, requirements.txt:5):

* ``reduce='sum'/'add'`` accumulates into ``out`` (prior contents are kept).
* ``reduce='mean'`` sums into ``out`` and then divides every row by
  ``clamp(count(index), min=1)`` -- the count ignores what ``out`` held, which
  is what gives the reference its "self term, divisor excludes self" quirk at
  ``analysisgnn/models/core/gnn.py:74`` and ``analysisgnn/models/analysis.py:586``.
"""
import torch


def _rows(index, like):
    view = index.view(-1, *([1] * (like.dim() - 1)))
    return view.expand_as(like)


def scatter_sum(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0, "the reference only scatters along dim 0"
    if out is None:
        if dim_size is None:
            dim_size = int(index.max()) + 1 if index.numel() else 0
        out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    return out.scatter_add_(0, _rows(index, src), src)


scatter_add = scatter_sum


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    total = scatter_sum(src, index, dim, out, dim_size)
    count = torch.zeros(total.shape[0], dtype=src.dtype, device=src.device)
    count.scatter_add_(0, index, torch.ones_like(index, dtype=src.dtype))
    count = count.clamp_(min=1).view(-1, *([1] * (total.dim() - 1)))
    if total.is_floating_point():
        total.div_(count)
    else:
        total.div_(count, rounding_mode="floor")
    return total


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    raise NotImplementedError(f"reduce={reduce!r} is not used on the hot path")
