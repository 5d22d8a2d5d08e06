"""The hot-path part of the reference's production model ``TorchAnalysisGNN``
(analysisgnn/models/analysis.py:421-591): embeddings -> per-node-type
``project_dict`` -> encoder -> onset pooling -> ``project_enc`` -> per-task heads.

Attribute names (and therefore ``state_dict`` keys) are the reference's:
``pitch_embedding, key_embedding, project_dict.<type>.{0,2,4}, encoder.*,
project_enc.{0,1,3,5,7,9}, clf_dict.<task>.{0,2,3}``.  Logit fusion
(``logit_fusion=True``, off in the reference's training default,
analysisgnn/models/analysis.py:852) and the optional output RNN are outside the
message-passing path and are not provided.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import graph, ops
from .layers import MLP, LayerNorm, Linear
from .hetero import HybridGNN, HybridHGT, MetricalGNN

ONSET = ("note", "onset", "note")


def onset_pool(x, onset_edges, batch_size):
    """analysisgnn/models/analysis.py:580-586: keep onset edges with both ends
    ``< batch_size``, drop self loops, ``scatter_mean(x[e1], e0, out=x.clone())``
    = ``(x_i + sum_j x_j) / max(deg_i, 1)``.  The filter is folded into the CSR build
    (filtered edges get relation id -1 and are dropped there), so nothing is
    compacted on the host and no shape depends on the data."""
    def make():
        e0, e1 = onset_edges[0], onset_edges[1]
        keep = (e0 < batch_size) & (e1 < batch_size) & (e0 != e1)
        return keep.long() - 1                                   # 0 = keep, -1 = dropped by agnn_csr_build
    etype = graph.derived(onset_edges, ("onset_pool", int(batch_size)), make)
    csr = graph.typed_csr(onset_edges, etype, int(batch_size), 1)
    return ops.segment_mean_self(x, x, csr)


class AnalysisEncoder(nn.Module):
    """Constructor / ``encode`` / ``forward`` arguments of ``TorchAnalysisGNN``
    (analysisgnn/models/analysis.py:422, 546, 571)."""

    def __init__(self, metadata, in_channels, hidden_channels, out_channels, task_dict, num_layers, dropout=0.5,
                 use_jk=False, logit_fusion=False, use_rnn=False, encoder_type="hybridgnn"):
        super().__init__()
        if logit_fusion or use_rnn:
            raise NotImplementedError("logit_fusion / use_rnn are outside the message-passing hot path")
        self.pitch_embedding = nn.Embedding(35, 64)
        self.key_embedding = nn.Embedding(15, 64)
        self.logit_fusion = False
        self.use_rnn = False
        self.hidden_channels = hidden_channels

        def mlp(cin):
            return MLP(Linear(cin, hidden_channels), nn.ReLU(), LayerNorm(hidden_channels),
                       nn.Dropout(dropout), Linear(hidden_channels, hidden_channels))

        self.project_dict = nn.ModuleDict({k: mlp(in_channels + 128 if k == "note" else in_channels)
                                           for k in metadata[0]})
        if encoder_type == "hgt":
            self.encoder = HybridHGT(metadata=metadata, input_channels=hidden_channels,
                                     hidden_channels=hidden_channels, num_layers=num_layers, heads=4,
                                     dropout=dropout, use_jk=use_jk)
        elif encoder_type == "hybridgnn":
            self.encoder = HybridGNN(metadata=metadata, input_channels=hidden_channels,
                                     hidden_channels=hidden_channels, num_layers=num_layers, dropout=dropout,
                                     use_jk=use_jk)
        elif encoder_type == "metricalgnn":
            self.encoder = MetricalGNN(metadata=metadata, input_channels=hidden_channels,
                                       hidden_channels=hidden_channels, output_channels=hidden_channels,
                                       num_layers=num_layers, dropout=dropout, use_jk=use_jk, fast=True)
        else:
            raise ValueError(f"unknown encoder_type {encoder_type!r}")
        self.encoder_type = encoder_type
        self.project_enc = MLP(
            LayerNorm(2 * hidden_channels), Linear(2 * hidden_channels, hidden_channels), nn.ReLU(),
            LayerNorm(hidden_channels), nn.Dropout(dropout), Linear(hidden_channels, out_channels), nn.ReLU(),
            LayerNorm(out_channels), nn.Dropout(dropout), Linear(out_channels, out_channels))
        self.clf_dict = nn.ModuleDict({
            task: MLP(Linear(out_channels, out_channels // 2), nn.ReLU(),
                      LayerNorm(out_channels // 2), Linear(out_channels // 2, n_cls))
            for task, n_cls in task_dict.items()})

    def encode(self, pitch_spelling, key_signature, x_dict, edge_index_dict, batch_dict, batch_size,
               neighbor_mask_node=None, neighbor_mask_edge=None):
        z = dict(x_dict)
        z["note"] = torch.cat((x_dict["note"], ops.embedding(pitch_spelling, self.pitch_embedding.weight),
                               ops.embedding(key_signature, self.key_embedding.weight)), dim=-1)
        # the per-node-type projections run stage by stage as ONE grouped launch each; their outputs share one
        # operand scale (the inputs of the first message-passing layer)
        keys = list(self.project_dict.keys())
        h = dict(zip(keys, MLP.forward_group([self.project_dict[k] for k in keys], [z[k] for k in keys],
                                             share_amax=True)))
        x = self.encoder(x_dict=h, edge_index_dict=edge_index_dict, batch_dict=batch_dict, batch_size=batch_size,
                         neighbor_mask_node=neighbor_mask_node, neighbor_mask_edge=neighbor_mask_edge,
                         return_edge_index=False, edge_attr_dict=None)
        if self.encoder_type == "metricalgnn":
            x = x if x.shape[0] == batch_size else x[:batch_size]
        pooled = onset_pool(x, edge_index_dict[ONSET], batch_size)
        from .. import fused
        return self.project_enc(fused.concat_cols((x, pooled)))

    def forward_clf(self, x, tasks=None):
        tasks = list(self.clf_dict.keys() if tasks is None else tasks)
        # all task heads stage by stage: one grouped launch per stage and direction instead of one chain per task
        return dict(zip(tasks, MLP.forward_group([self.clf_dict[t] for t in tasks], [x] * len(tasks))))

    def clf_task(self, x, task_name):
        return self.clf_dict[task_name](x)

    def forward(self, pitch_spelling, key_signature, x_dict, edge_index_dict, batch_dict, batch_size,
                neighbor_mask_node=None, neighbor_mask_edge=None):
        x = self.encode(pitch_spelling, key_signature, x_dict, edge_index_dict, batch_dict, batch_size,
                        neighbor_mask_node, neighbor_mask_edge)
        return self.forward_clf(x)


def multitask_ce(logits, labels):
    """Default multi-task objective of ``ContinualAnalysisGNN`` (analysisgnn/models/analysis.py:
    881-908, 1035-1037): ``MultiTaskLoss(requires_grad=False)`` = plain sum of per-task
    ``CrossEntropyLoss(ignore_index=-1, label_smoothing=0.1)`` divided by the number of tasks."""
    total = sum(ops.cross_entropy(logits[t], labels[t], ignore_index=-1, label_smoothing=0.1) for t in labels)
    return total / len(labels)


class CrossEntropyLoss(nn.Module):
    """``nn.CrossEntropyLoss(ignore_index, label_smoothing)`` with mean reduction on ``agnn_softmax_ce_fwd/_bwd`` -- the
    per-task criterion ``ContinualAnalysisGNN`` builds (analysisgnn/models/analysis.py:881-888)."""

    def __init__(self, ignore_index: int = -100, label_smoothing: float = 0.0):
        super().__init__()
        self.ignore_index, self.label_smoothing = ignore_index, label_smoothing

    def forward(self, logits, target):
        return ops.cross_entropy(logits, target, ignore_index=self.ignore_index, label_smoothing=self.label_smoothing)


class MultiTaskLoss(nn.Module):
    """Drop-in for the reference's ``MultiTaskLoss(tasks, loss_ft, loss_weights=None, requires_grad=True)``
    (analysisgnn/models/chord.py:16-49): per-task criteria from ``loss_ft``; with ``requires_grad`` the learned
    weighting of Liebel & Koerner, ``sum_i 0.5 / p_i^2 * L_i + log(1 + p_i^2)`` with ``params`` initialised to one
    (same ``state_dict`` key), else the plain sum.  Returns the per-task losses plus ``"total"``; ``p_i`` belongs to
    the i-th task PRESENT in ``gt`` (the reference enumerates ``gt``'s keys, :41-44).  The caller divides ``total`` by
    the number of tasks (analysis.py:1035-1037)."""

    def __init__(self, tasks, loss_ft, loss_weights=None, requires_grad=True):
        super().__init__()
        if set(tasks) != set(loss_ft.keys()):
            raise AssertionError("tasks and loss_ft must name the same tasks")
        if loss_weights is not None and set(tasks) != set(loss_weights.keys()):
            raise AssertionError("tasks and loss_weights must name the same tasks")
        self.loss_weights = loss_weights if loss_weights is not None else {task: 1 for task in tasks}
        self.tasks, self.loss_ft, self.requires_grad = tasks, loss_ft, requires_grad
        if requires_grad:
            self.params = nn.Parameter(torch.ones(len(tasks)))
        else:
            self.params = torch.ones(len(tasks), requires_grad=False)

    def forward(self, pred, gt):
        out = {task: self.loss_ft[task](pred[task], gt[task]) for task in gt.keys()}
        total = 0
        for i, loss in enumerate(list(out.values())):
            if self.requires_grad:
                total = total + (0.5 / (self.params[i] ** 2) * loss + torch.log(1 + self.params[i] ** 2))
            else:
                total = total + loss
        out["total"] = total
        return out

