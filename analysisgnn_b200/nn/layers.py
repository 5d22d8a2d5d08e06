"""``nn.Linear`` / ``nn.LayerNorm`` with the same parameters and ``state_dict`` keys, running on
libagnn's tensor-core GEMM (agnn_gemm) and row-normalisation kernels."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


class Linear(nn.Linear):
    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


class LayerNorm(nn.LayerNorm):
    def forward(self, x):
        if len(self.normalized_shape) != 1 or not self.elementwise_affine or self.bias is None:
            return super().forward(x)
        return ops.layer_norm(x, self.weight, self.bias, self.eps)


class GRU(nn.GRU):
    """``nn.GRU`` (same parameters / ``state_dict``): batch-first fp32 CUDA inputs with hidden size 32, 64 or
    128 run on libagnn (tensor-core GEMMs for every projection + agnn_gru_fwd / _bwd for the recurrence);
    anything else (other sizes, given ``h0``, packed sequences, projections) goes to cuDNN."""

    def forward(self, x, hx=None):
        if (hx is not None or not isinstance(x, torch.Tensor) or not self.batch_first or self.proj_size != 0
                or not ops.gru_supported(x, self.hidden_size)):
            if isinstance(x, torch.Tensor) and x.is_cuda:
                from .. import _lib
                _lib.library_route("GRU outside agnn_gru's shapes (hidden size, given h0, packed input): cuDNN")
            return super().forward(x, hx)
        n_dir = 2 if self.bidirectional else 1
        h_n = []
        for layer in range(self.num_layers):
            params = []
            for d in range(n_dir):
                sfx = f"_l{layer}" + ("_reverse" if d else "")
                w_ih, w_hh = getattr(self, "weight_ih" + sfx), getattr(self, "weight_hh" + sfx)
                if self.bias:
                    b_ih, b_hh = getattr(self, "bias_ih" + sfx), getattr(self, "bias_hh" + sfx)
                else:
                    b_ih = b_hh = torch.zeros(3 * self.hidden_size, dtype=x.dtype, device=x.device)
                params += [w_ih, w_hh, b_ih, b_hh]
            x = ops.gru_layer(x, params)
            h_n.append(x[:, -1, :self.hidden_size])
            if n_dir == 2:
                h_n.append(x[:, 0, self.hidden_size:])
            if self.dropout > 0 and self.training and layer < self.num_layers - 1:
                x = F.dropout(x, self.dropout, True)
        return x, torch.stack(h_n, dim=0)
