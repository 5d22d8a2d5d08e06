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


class MLP(nn.Sequential):
    """``nn.Sequential`` of Linear / ReLU / LayerNorm / Dropout modules (same children, same ``state_dict`` keys) whose
    forward runs as fused projection stages ``[LayerNorm ->] [Dropout ->] Linear [-> ReLU]`` (fused.stage_group): the
    reference's project_dict / project_enc / clf_dict / sequence-MLP blocks (analysisgnn/models/analysis.py:429-443,
    474-496; models/cadence.py:252-260).  ``forward_group`` runs several structurally identical MLPs (one per node
    type, one per task head) stage by stage with ONE grouped GEMM launch per stage and direction.  Sequences that are
    not made of such stages run module by module."""

    def stages(self, pre_norm=None):
        """[(LayerNorm or None, dropout p, Linear, ReLU module or None)] or None if the children do not parse into
        stages."""
        out, norm, p = [], pre_norm, 0.0
        mods = list(self.children())
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.LayerNorm) and norm is None and p == 0.0 and len(m.normalized_shape) == 1 \
                    and m.elementwise_affine and m.bias is not None:
                norm = m
            elif isinstance(m, nn.Dropout) and p == 0.0:
                p = m.p
            elif isinstance(m, nn.Linear):
                relu = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU) else None
                out.append((norm, p, m, relu))
                norm, p = None, 0.0
                i += 1 if relu is not None else 0
            else:
                return None
            i += 1
        return out if norm is None and p == 0.0 else None

    def forward(self, x, pre_norm=None):
        return MLP.forward_group([self], [x], pre_norms=[pre_norm])[0]

    @staticmethod
    def forward_group(mlps, xs, pre_norms=None, share_amax=False):
        from .. import fused
        pre_norms = pre_norms if pre_norms is not None else [None] * len(mlps)
        plans = [m.stages(pn) for m, pn in zip(mlps, pre_norms)]
        same = all(pl is not None and len(pl) == len(plans[0]) and
                   [(s[0] is None, s[1], s[3] is None) for s in pl] == [(s[0] is None, s[1], s[3] is None) for s in plans[0]]
                   for pl in plans) if plans and plans[0] is not None else False
        if not same or not all(x.is_cuda for x in xs):
            outs = []
            for m, x, pn in zip(mlps, xs, pre_norms):
                x = pn(x) if pn is not None else x
                outs.append(nn.Sequential.forward(m, x))
            return outs
        xs = list(xs)
        training = mlps[0].training
        for k in range(len(plans[0])):
            st = [pl[k] for pl in plans]
            norms = [None if s[0] is None else (s[0].weight, s[0].bias, s[0].eps) for s in st]
            xs = fused.stage_group(xs, [s[2].weight for s in st], [s[2].bias for s in st], norms,
                                   relu=st[0][3] is not None, dropout=st[0][1], training=training,
                                   share_amax=share_amax and k == len(plans[0]) - 1)
            for s, x in zip(st, xs):        # the ReLU ran in the GEMM epilogue: its module's forward hooks still see it
                if s[3] is not None and s[3]._forward_hooks:
                    for hook in list(s[3]._forward_hooks.values()):
                        hook(s[3], (x,), x)
        return xs


class GRU(nn.GRU):
    """``nn.GRU`` (same parameters / ``state_dict``): batch-first fp32 CUDA inputs run on libagnn (tensor-core GEMMs
    for every projection + agnn_gru_fwd / _bwd for the recurrence): hidden size 32, 64 or 128 with W_hh resident in
    registers, multiples of 64 from 192 to 2048 (MetricalConvLayer's 512) as one launch per time step.  Anything else
    (other sizes, given ``h0``, packed sequences, projections) goes to cuDNN and is counted as a library route."""

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
            drop = None
            if layer > 0 and self.dropout > 0 and self.training:
                from .. import linalg
                drop = (float(self.dropout), linalg.dropout_site())   # nn.GRU: dropout on the outputs of every layer
            x = ops.gru_layer(x, params, dropout=drop)                # but the last = on the inputs of layers >= 1
            h_n.append(x[:, -1, :self.hidden_size])
            if n_dir == 2:
                h_n.append(x[:, 0, self.hidden_size:])
        return x, torch.stack(h_n, dim=0)
