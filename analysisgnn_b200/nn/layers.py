"""``nn.Linear`` / ``nn.LayerNorm`` with the same parameters and ``state_dict`` keys, running on
libagnn's tensor-core GEMM (agnn_gemm) and row-normalisation kernels."""
from __future__ import annotations

import torch.nn as nn

from .. import ops


class Linear(nn.Linear):
    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


class LayerNorm(nn.LayerNorm):
    def forward(self, x):
        if len(self.normalized_shape) != 1 or not self.elementwise_affine or self.bias is None:
            return super().forward(x)
        return ops.layer_norm(x, self.weight, self.bias, self.eps)
