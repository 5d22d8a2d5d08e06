"""Pure-torch CPU restatement of the PyG / graphmuse shaped encoders the
reference's production model uses (``analysisgnn/models/analysis.py:444-473``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

**Parity unpinned.**  torch_geometric (pinned only as ``>=2.3.0``,
requirements.txt:4), pyg-lib and graphmuse (un-pinned, requirements.txt:21) are
third-party, absent from /root/reference and not installable here; the
reference's tests hold no vectors for them.  The operators below restate the
published PyG >= 2.3 semantics and the wiring of the only
in-tree statement of the stack (``analysisgnn/models/cadence.py:142-176,
229-332``), per SURVEY.md Appendix A.  Reports say "oracle = this repo's
restatement".

Convention (PyG): ``edge_index[0]`` = source j, ``edge_index[1]`` = target i.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def rel_key(edge_type) -> str:
    return "__".join(edge_type)


# --------------------------------------------------------------- ReLU with a test hook
# Every ReLU of the restated encoders goes through ``_relu``.  ``relu_override`` is None in normal use (plain
# ``z.relu()``).  The full-size parity tests (tests/test_fullsize_gpu.py) set it to evaluate the oracle's GRADIENT with
# the CUDA side's choice of subgradient at the handful of |z| ~ 1e-7 units that two correct fp32 implementations put
# on opposite sides of zero (the forward value is untouched); tags are module names (``assign_tags``).
relu_override = None


def _relu(z, tag, key=""):
    return z.relu() if relu_override is None else relu_override(tag, key, z)


class ReLU(nn.ReLU):
    """``nn.ReLU`` (no parameters, same ``state_dict``) routed through ``_relu``."""

    def forward(self, z):
        return _relu(z, getattr(self, "_tag", None))


class MaskedReLU(torch.autograd.Function):
    """forward ``relu(z)``; backward ``g * mask`` with a GIVEN mask instead of ``z > 0``."""

    @staticmethod
    def forward(ctx, z, mask):
        ctx.save_for_backward(mask)
        return z.relu()

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * mask.to(g.dtype), None


def assign_tags(model):
    """``module._tag`` = its name inside ``model`` (what ``_relu`` reports to the override)."""
    for name, m in model.named_modules():
        m._tag = name
    return model


# ------------------------------------------------------------------ primitives

def scatter_mean_rows(values, rows, n_rows):
    """PyG ``aggr='mean'``: mean over incoming messages, 0 for isolated targets."""
    out = values.new_zeros((n_rows, values.shape[1])).index_add_(0, rows, values)
    cnt = torch.bincount(rows, minlength=n_rows).clamp_(min=1).to(values.dtype)
    return out / cnt.unsqueeze(-1)


def segment_softmax(scores, rows, n_rows):
    """``torch_geometric.utils.softmax``: max-shifted, denominator + 1e-16."""
    shape = (n_rows,) + tuple(scores.shape[1:])
    idx = rows.view(-1, *([1] * (scores.dim() - 1))).expand_as(scores)
    top = scores.new_full(shape, float("-inf")).scatter_reduce_(0, idx, scores.detach(), "amax", include_self=True)
    top = torch.where(torch.isinf(top), torch.zeros_like(top), top)
    ex = (scores - top[rows]).exp()
    den = scores.new_zeros(shape).index_add_(0, rows, ex) + 1e-16
    return ex / den[rows]


def trim_to_layer(layer, nodes_per_hop, edges_per_hop, x_dict, ei_dict):
    """``torch_geometric.utils.trim_to_layer`` for dict inputs (cadence.py:166-173):
    for layer > 0 drop the nodes / edges of the outermost remaining hop."""
    if layer <= 0:
        return x_dict, ei_dict
    x_dict = {k: v[: v.size(0) - nodes_per_hop[k][-layer]] for k, v in x_dict.items()}
    ei_dict = {k: v[:, : v.size(1) - edges_per_hop[k][-layer]] for k, v in ei_dict.items()}
    return x_dict, ei_dict


# ----------------------------------------------------------------------- SAGE

class SAGEConv(nn.Module):
    """``SAGEConv(in, out, aggr='mean', root_weight=True, bias=True)``:
    ``lin_l(mean_j x_j) + lin_r(x_i)`` (``lin_r`` has no bias)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin_l = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x_src, x_dst, edge_index):
        agg = scatter_mean_rows(x_src[edge_index[0]], edge_index[1], x_dst.size(0))
        return self.lin_l(agg) + self.lin_r(x_dst)


class HeteroSAGELayer(nn.Module):
    """PyG ``HeteroConv({et: SAGEConv}, aggr)``: per relation present in
    ``edge_index_dict``; results grouped per destination type by ``aggr``;
    destination types that receive nothing are absent from the output."""

    def __init__(self, edge_types, in_channels, out_channels, aggr="sum"):
        super().__init__()
        self.edge_types = [tuple(et) for et in edge_types]
        self.aggr = aggr
        self.convs = nn.ModuleDict({rel_key(et): SAGEConv(in_channels, out_channels) for et in self.edge_types})

    def forward(self, x_dict, ei_dict):
        outs = {}
        for et in self.edge_types:
            src, _, dst = et
            if et not in ei_dict or src not in x_dict or dst not in x_dict:
                continue
            outs.setdefault(dst, []).append(self.convs[rel_key(et)](x_dict[src], x_dict[dst], ei_dict[et]))
        red = {"sum": lambda t: t.sum(0), "mean": lambda t: t.mean(0)}[self.aggr]
        return {k: red(torch.stack(v, dim=0)) for k, v in outs.items()}


class HeteroSAGEStack(nn.Module):
    """cadence.py:142-176: trim_to_layer -> HeteroConv{SAGEConv}(aggr='sum') -> relu, L times."""

    def __init__(self, edge_types, in_channels, hidden_channels, num_layers, aggr="sum"):
        super().__init__()
        self.convs = nn.ModuleList(
            HeteroSAGELayer(edge_types, in_channels if i == 0 else hidden_channels, hidden_channels, aggr)
            for i in range(num_layers))

    def forward(self, x_dict, ei_dict, nodes_per_hop=None, edges_per_hop=None, collect=None):
        for i, conv in enumerate(self.convs):
            if edges_per_hop is not None:
                x_dict, ei_dict = trim_to_layer(i, nodes_per_hop, edges_per_hop, x_dict, ei_dict)
            x_dict = {k: _relu(v, getattr(conv, "_tag", None), k) for k, v in conv(x_dict, ei_dict).items()}
            if collect is not None:
                collect.append(x_dict)
        return x_dict


# ------------------------------------------------------------------------ HGT

class HGTConv(nn.Module):
    """PyG >= 2.3 ``HGTConv(in, out, metadata, heads)``: per-node-type fused KQV
    projection; per-(relation, head) bias-free DxD ``k_rel`` / ``v_rel`` applied
    to the SOURCE type's k, v; ``p_rel`` per relation and head; softmax over ALL
    incoming edges of a target across relations; ``out_lin(gelu(.))`` and a gated
    skip when dims match.  ``joint_softmax=False`` gives the pre-2.3 per-relation
    softmax (BASELINE.json's wording), results then summed over relations."""

    def __init__(self, in_channels, out_channels, metadata, heads=1, joint_softmax=True):
        super().__init__()
        assert out_channels % heads == 0
        self.node_types = list(metadata[0])
        self.edge_types = [tuple(et) for et in metadata[1]]
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.joint_softmax = joint_softmax
        d = out_channels // heads
        self.kqv_lin = nn.ModuleDict({t: nn.Linear(in_channels, 3 * out_channels) for t in self.node_types})
        self.out_lin = nn.ModuleDict({t: nn.Linear(out_channels, out_channels) for t in self.node_types})
        n_rel = len(self.edge_types)
        # index = head * n_rel + relation (PyG HeteroLinear type vector layout); x @ W[type]
        self.k_rel = nn.Parameter(torch.empty(heads * n_rel, d, d))
        self.v_rel = nn.Parameter(torch.empty(heads * n_rel, d, d))
        self.skip = nn.ParameterDict({t: nn.Parameter(torch.ones(1)) for t in self.node_types})
        self.p_rel = nn.ParameterDict({rel_key(et): nn.Parameter(torch.ones(1, heads)) for et in self.edge_types})
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / math.sqrt(self.k_rel.shape[1])
        nn.init.uniform_(self.k_rel, -bound, bound)
        nn.init.uniform_(self.v_rel, -bound, bound)
        for lin in list(self.kqv_lin.values()) + list(self.out_lin.values()):
            lin.reset_parameters()
        for p in self.skip.values():
            nn.init.ones_(p)
        for p in self.p_rel.values():
            nn.init.ones_(p)

    def forward(self, x_dict, ei_dict):
        H, D = self.heads, self.out_channels // self.heads
        n_rel = len(self.edge_types)
        k_d, q_d, v_d = {}, {}, {}
        for t, x in x_dict.items():
            k, q, v = torch.tensor_split(self.kqv_lin[t](x), 3, dim=1)
            k_d[t], q_d[t], v_d[t] = k.reshape(-1, H, D), q.reshape(-1, H, D), v.reshape(-1, H, D)
        per_dst = {}
        for r, et in enumerate(self.edge_types):
            if et not in ei_dict:
                continue
            src, _, dst = et
            wk = self.k_rel.view(H, n_rel, D, D)[:, r]
            wv = self.v_rel.view(H, n_rel, D, D)[:, r]
            k = torch.einsum("nhd,hde->nhe", k_d[src], wk)
            v = torch.einsum("nhd,hde->nhe", v_d[src], wv)
            ei = ei_dict[et]
            score = (q_d[dst][ei[1]] * k[ei[0]]).sum(-1) * self.p_rel[rel_key(et)] / math.sqrt(D)
            per_dst.setdefault(dst, []).append((score, v[ei[0]], ei[1]))
        out_dict = {}
        for dst, parts in per_dst.items():
            n = x_dict[dst].size(0)
            if self.joint_softmax:
                score = torch.cat([p[0] for p in parts]); val = torch.cat([p[1] for p in parts])
                rows = torch.cat([p[2] for p in parts])
                alpha = segment_softmax(score, rows, n)
                agg = val.new_zeros((n, H, D)).index_add_(0, rows, val * alpha.unsqueeze(-1))
            else:
                agg = x_dict[dst].new_zeros((n, H, D))
                for score, val, rows in parts:
                    alpha = segment_softmax(score, rows, n)
                    agg = agg.index_add(0, rows, val * alpha.unsqueeze(-1))
            o = self.out_lin[dst](F.gelu(agg.reshape(n, H * D)))
            if o.size(-1) == x_dict[dst].size(-1):
                a = self.skip[dst].sigmoid()
                o = a * o + (1 - a) * x_dict[dst]
            out_dict[dst] = o
        return out_dict


class HeteroHGTStack(nn.Module):
    def __init__(self, metadata, in_channels, hidden_channels, num_layers, heads, dropout=0.0, joint_softmax=True):
        super().__init__()
        self.dropout = dropout
        self.convs = nn.ModuleList(
            HGTConv(in_channels if i == 0 else hidden_channels, hidden_channels, metadata, heads, joint_softmax)
            for i in range(num_layers))

    def forward(self, x_dict, ei_dict, nodes_per_hop=None, edges_per_hop=None, collect=None):
        for i, conv in enumerate(self.convs):
            if edges_per_hop is not None:
                x_dict, ei_dict = trim_to_layer(i, nodes_per_hop, edges_per_hop, x_dict, ei_dict)
            x_dict = {k: F.dropout(_relu(v, getattr(conv, "_tag", None), k), self.dropout, self.training)
                      for k, v in conv(x_dict, ei_dict).items()}
            if collect is not None:
                collect.append(x_dict)
        return x_dict


# ------------------------------------------------------------ hybrid encoders

class SequenceBranch(nn.Module):
    """cadence.py:249-260, 276-285: split by graph -> pad -> 2-layer biGRU ->
    LayerNorm -> MLP -> unpad."""

    def __init__(self, in_channels, hidden_channels, dropout):
        super().__init__()
        self.rnn = nn.GRU(input_size=in_channels, hidden_size=hidden_channels // 2, num_layers=2,
                          batch_first=True, bidirectional=True, dropout=dropout)
        self.rnn_norm = nn.LayerNorm(hidden_channels)
        self.rnn_mlp = nn.Sequential(
            nn.Linear(hidden_channels, hidden_channels), ReLU(), nn.LayerNorm(hidden_channels),
            nn.Dropout(dropout), nn.Linear(hidden_channels, hidden_channels))

    def forward(self, x, batch):
        lengths = torch.bincount(batch)
        seq = nn.utils.rnn.pad_sequence(x.split(lengths.tolist()), batch_first=True, padding_value=0.0)
        seq, _ = self.rnn(seq)
        seq = self.rnn_mlp(self.rnn_norm(seq))
        return torch.cat(nn.utils.rnn.unpad_sequence(seq, batch_first=True, lengths=lengths.cpu()), dim=0)


class JumpingKnowledge(nn.Module):
    """analysisgnn/models/core/gnn.py:345-365 (LSTM attention over layer outputs)."""

    def __init__(self, n_hidden, n_layers):
        super().__init__()
        self.lstm = nn.LSTM(n_hidden, (n_layers * n_hidden) // 2, bidirectional=True, batch_first=True)
        self.att = nn.Linear(2 * ((n_layers * n_hidden) // 2), 1)

    def forward(self, xs):
        x = torch.stack(xs, dim=1)
        alpha, _ = self.lstm(x)
        alpha = torch.softmax(self.att(alpha).squeeze(-1), dim=-1)
        return (x * alpha.unsqueeze(-1)).sum(dim=1)


class _HybridBase(nn.Module):
    def _finish(self, x_dict, collect, x_in, batch_dict, batch_size):
        if self.use_jk:
            x_gnn = self.jk([c["note"][:batch_size] for c in collect])
        else:
            x_gnn = x_dict["note"][:batch_size]
        batch = batch_dict["note"][:batch_size] if batch_dict is not None else \
            torch.zeros(batch_size, dtype=torch.long, device=x_gnn.device)
        x_seq = self.seq(x_in[:batch_size], batch)
        return self.cat_proj(torch.cat((x_gnn, x_seq), dim=-1))

    def forward(self, x_dict, edge_index_dict, batch_dict=None, batch_size=None, neighbor_mask_node=None,
                neighbor_mask_edge=None, return_edge_index=False, edge_attr_dict=None):
        batch_size = x_dict["note"].size(0) if batch_size is None else batch_size
        collect = [] if self.use_jk else None
        out = self.gnn(x_dict, edge_index_dict, neighbor_mask_node, neighbor_mask_edge, collect)
        return self._finish(out, collect, x_dict["note"], batch_dict, batch_size)


class HybridGNN(_HybridBase):
    """HeteroSAGEStack on the graph + GRU branch over the target
    notes, joined by ``Linear(2H, H)`` (cadence.py:301-303)."""

    def __init__(self, metadata, input_channels, hidden_channels, num_layers, dropout=0.5, use_jk=False):
        super().__init__()
        self.use_jk = use_jk
        self.gnn = HeteroSAGEStack(metadata[1], input_channels, hidden_channels, num_layers)
        self.seq = SequenceBranch(input_channels, hidden_channels, dropout)
        self.cat_proj = nn.Linear(2 * hidden_channels, hidden_channels)
        if use_jk:
            self.jk = JumpingKnowledge(hidden_channels, num_layers)


class HybridHGT(_HybridBase):
    def __init__(self, metadata, input_channels, hidden_channels, num_layers, heads=4, dropout=0.5, use_jk=False,
                 joint_softmax=True):
        super().__init__()
        self.use_jk = use_jk
        self.gnn = HeteroHGTStack(metadata, input_channels, hidden_channels, num_layers, heads, dropout, joint_softmax)
        self.seq = SequenceBranch(input_channels, hidden_channels, dropout)
        self.cat_proj = nn.Linear(2 * hidden_channels, hidden_channels)
        if use_jk:
            self.jk = JumpingKnowledge(hidden_channels, num_layers)


class MetricalGNN(nn.Module):
    """graphmuse-flavoured ``MetricalGNN`` (call shape cadence.py:232-234, 298-300):
    HeteroSAGEStack -> note rows -> MLP; returns all (trimmed) note rows."""

    def __init__(self, metadata, input_channels, hidden_channels, output_channels, num_layers, dropout=0.5,
                 use_jk=False, fast=True):
        super().__init__()
        self.gnn = HeteroSAGEStack(metadata[1], input_channels, hidden_channels, num_layers)
        self.mlp = nn.Sequential(
            nn.Linear(hidden_channels, hidden_channels), ReLU(), nn.LayerNorm(hidden_channels),
            nn.Dropout(dropout), nn.Linear(hidden_channels, output_channels))

    def forward(self, x_dict, edge_index_dict, neighbor_mask_node=None, neighbor_mask_edge=None, **kwargs):
        out = self.gnn(x_dict, edge_index_dict, neighbor_mask_node, neighbor_mask_edge)
        return self.mlp(out["note"])


# ----------------------------------------------------- shell around the encoder

class AnalysisEncoderShell(nn.Module):
    """The hot-path part of ``TorchAnalysisGNN`` (analysisgnn/models/analysis.py:
    421-485, 571-591): embeddings -> per-node-type ``project_dict`` -> encoder ->
    onset pooling -> ``project_enc`` -> per-task MLP heads (``clf_dict``, no
    logit fusion)."""

    def __init__(self, metadata, in_channels, hidden_channels, out_channels, task_dict, num_layers, dropout=0.5,
                 use_jk=False, encoder_type="hybridgnn"):
        super().__init__()
        self.pitch_embedding = nn.Embedding(35, 64)
        self.key_embedding = nn.Embedding(15, 64)
        self.hidden_channels = hidden_channels

        def mlp(cin):
            return nn.Sequential(nn.Linear(cin, hidden_channels), ReLU(), nn.LayerNorm(hidden_channels),
                                 nn.Dropout(dropout), nn.Linear(hidden_channels, hidden_channels))

        self.project_dict = nn.ModuleDict({k: mlp(in_channels + 128 if k == "note" else in_channels)
                                           for k in metadata[0]})
        if encoder_type == "hgt":
            self.encoder = HybridHGT(metadata, hidden_channels, hidden_channels, num_layers, heads=4,
                                     dropout=dropout, use_jk=use_jk)
        elif encoder_type == "hybridgnn":
            self.encoder = HybridGNN(metadata, hidden_channels, hidden_channels, num_layers, dropout=dropout,
                                     use_jk=use_jk)
        else:
            raise ValueError(encoder_type)
        self.project_enc = nn.Sequential(
            nn.LayerNorm(2 * hidden_channels), nn.Linear(2 * hidden_channels, hidden_channels), ReLU(),
            nn.LayerNorm(hidden_channels), nn.Dropout(dropout), nn.Linear(hidden_channels, out_channels), ReLU(),
            nn.LayerNorm(out_channels), nn.Dropout(dropout), nn.Linear(out_channels, out_channels))
        self.clf_dict = nn.ModuleDict({
            t: nn.Sequential(nn.Linear(out_channels, out_channels // 2), ReLU(),
                             nn.LayerNorm(out_channels // 2), nn.Linear(out_channels // 2, c))
            for t, c in task_dict.items()})

    def encode(self, pitch_spelling, key_signature, x_dict, edge_index_dict, batch_dict, batch_size,
               neighbor_mask_node=None, neighbor_mask_edge=None):
        from .intree import onset_pool
        z = {k: v.clone() for k, v in x_dict.items()}
        z["note"] = torch.cat((z["note"], self.pitch_embedding(pitch_spelling), self.key_embedding(key_signature)), -1)
        h = {k: self.project_dict[k](z[k]) for k in self.project_dict.keys()}
        x = self.encoder(x_dict=h, edge_index_dict=edge_index_dict, batch_dict=batch_dict, batch_size=batch_size,
                         neighbor_mask_node=neighbor_mask_node, neighbor_mask_edge=neighbor_mask_edge,
                         return_edge_index=False, edge_attr_dict=None)
        pooled = onset_pool(x, edge_index_dict[("note", "onset", "note")], batch_size)
        return self.project_enc(torch.cat((x, pooled), dim=-1))

    def forward(self, pitch_spelling, key_signature, x_dict, edge_index_dict, batch_dict, batch_size,
                neighbor_mask_node=None, neighbor_mask_edge=None):
        x = self.encode(pitch_spelling, key_signature, x_dict, edge_index_dict, batch_dict, batch_size,
                        neighbor_mask_node, neighbor_mask_edge)
        return {t: clf(x) for t, clf in self.clf_dict.items()}


def multitask_ce(logits, labels):
    """Default multi-task objective of ``ContinualAnalysisGNN`` (analysisgnn/models/
    analysis.py:881-908, 1035-1037): ``MultiTaskLoss(requires_grad=False)`` = plain
    sum of per-task ``CrossEntropyLoss(ignore_index=-1, label_smoothing=0.1)``,
    divided by the number of tasks."""
    total = sum(F.cross_entropy(logits[t], labels[t], ignore_index=-1, label_smoothing=0.1) for t in labels)
    return total / len(labels)
