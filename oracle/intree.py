"""Pure-torch CPU restatement of the reference's in-tree message-passing layers.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PINNED: checked against
the reference's own files executed through ``oracle/ref_loader.py`` and against
``tests/golden/intree_*.pt`` (``tests/test_oracle_pinned.py``).

Each module keeps the reference's constructor arguments, forward arguments and
``state_dict`` keys, so a reference ``state_dict`` loads unchanged.  The
arithmetic is written edge-list style with ``index_add_`` -- no fusion, no
reordering -- because this is the statement the CUDA path is compared with.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.rnn import pad_sequence


def _xavier_relu_(linear: nn.Linear):
    nn.init.xavier_uniform_(linear.weight, gain=nn.init.calculate_gain("relu"))
    if linear.bias is not None:
        nn.init.zeros_(linear.bias)


def mean_into_copy(values, rows, base):
    """torch_scatter ``scatter(values, rows, 0, out=base.clone(), reduce='mean')``:
    ``(base_i + sum of values landing on i) / max(count_i, 1)``; the count
    excludes the base term (analysisgnn/models/core/gnn.py:74)."""
    acc = base.clone().index_add_(0, rows, values)
    cnt = torch.bincount(rows, minlength=base.shape[0]).clamp_(min=1).to(values.dtype)
    return acc / cnt.unsqueeze(-1)


def sum_into(values, rows, n_rows):
    out = values.new_zeros((n_rows, values.shape[1]))
    return out.index_add_(0, rows, values)


class SageConvScatter(nn.Module):
    """analysisgnn/models/core/gnn.py:39-76.  Reduces at ``edge_index[0]`` reading
    ``edge_index[1]``."""

    def __init__(self, in_features, out_features, bias=True, in_edge_features=None):
        super().__init__()
        self.neigh_linear = nn.Linear(in_features, in_features, bias=bias)
        self.linear = nn.Linear(2 * in_features, out_features, bias=bias)
        self.in_edge_features = in_edge_features
        if in_edge_features is not None:
            self.edge_linear = nn.Linear(in_edge_features, in_features, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        _xavier_relu_(self.linear)
        _xavier_relu_(self.neigh_linear)
        if self.in_edge_features is not None:
            _xavier_relu_(self.edge_linear)

    def forward(self, features, edge_index, edge_features=None, neigh_feats=None):
        h = self.neigh_linear(features if neigh_feats is None else neigh_feats)
        if edge_index is None or edge_index.shape[1] == 0:      # gnn.py:67-69: [x || h]
            return self.linear(torch.cat((features, h), dim=-1))
        msg = h[edge_index[1]]
        if self.in_edge_features is not None and edge_features is not None:
            msg = msg + self.edge_linear(edge_features)
        s = mean_into_copy(msg, edge_index[0], features)
        return self.linear(torch.cat((features, s), dim=-1))


class ResGatedGraphConv(nn.Module):
    """analysisgnn/models/core/gnn.py:212-258 (``h1`` enters twice, :256-257)."""

    def __init__(self, in_features, out_features, bias=True, in_edge_features=None):
        super().__init__()
        self.W1 = nn.Linear(in_features, out_features, bias=bias)
        self.W2 = nn.Linear(in_features, out_features, bias=bias)
        self.W3 = nn.Linear(in_features, out_features, bias=bias)
        self.W4 = nn.Linear(in_features, out_features, bias=bias)
        self.in_edge_features = in_edge_features
        if in_edge_features is not None:
            self.W5 = nn.Linear(in_edge_features, out_features, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        for name in ("W1", "W2", "W3", "W4", "W5"):
            if hasattr(self, name):
                _xavier_relu_(getattr(self, name))

    def forward(self, features, edge_index, edge_features=None, neigh_feats=None):
        h1 = self.W1(features)
        h2 = self.W2(features if neigh_feats is None else neigh_feats)
        gate = self.W3(features)[edge_index[0]] + self.W4(features)[edge_index[1]]
        if edge_features is not None and self.in_edge_features is not None:
            gate = gate + self.W5(edge_features)
        msg = torch.sigmoid(gate) * h2[edge_index[1]]
        s = h1.clone().index_add_(0, edge_index[0], msg)
        return h1 + s


class RelEdgeConv(nn.Module):
    """analysisgnn/models/core/gnn.py:79-106."""

    def __init__(self, in_node_features, out_features, bias=True, in_edge_features=None):
        super().__init__()
        self.neigh_linear = nn.Linear(in_node_features, in_node_features, bias=bias)
        self.edge_linear = nn.Linear(in_node_features * 2 if in_edge_features is None
                                     else in_node_features + in_edge_features, in_node_features, bias=bias)
        self.linear = nn.Linear(in_node_features * 2, out_features, bias=bias)
        for lin in (self.linear, self.neigh_linear, self.edge_linear):
            _xavier_relu_(lin)

    def forward(self, features, edge_index, edge_features=None):
        h = self.neigh_linear(features)
        if edge_features is None:
            edge_features = torch.abs(h[edge_index[0]] - h[edge_index[1]])
        new_h = self.edge_linear(torch.cat((h[edge_index[1]], edge_features), dim=-1))
        s = mean_into_copy(new_h, edge_index[0], h)
        return self.linear(torch.cat((features, s), dim=-1))


class GATConvLayer(nn.Module):
    """analysisgnn/models/core/gnn.py:154-209.  The attention weight is ``softmax`` over the HEADS (dim=1) followed by
    the mean over the heads (:206) -- a constant 1 / num_heads for every edge, whatever the scores and the attention
    dropout are -- so the layer computes ``h_i + (1/H) sum_j h_j`` with ``h = linear(x)`` and ``el`` / ``er`` /
    ``attnl`` / ``attnr`` receive (numerically almost) zero gradients.  Restated literally."""

    def __init__(self, in_features, out_features, num_heads=3, bias=True, dropout=0.3, negative_slope=0.2,
                 in_edge_features=None):
        super().__init__()
        self.num_heads, self.in_features, self.out_features = num_heads, in_features, out_features
        self.linear = nn.Linear(in_features, out_features, bias=bias)
        self.el = nn.Linear(in_features, in_features * num_heads, bias=bias)
        self.er = nn.Linear(in_features, in_features * num_heads, bias=bias)
        self.attnl = nn.Parameter(torch.empty(1, num_heads, in_features))
        self.attnr = nn.Parameter(torch.empty(1, num_heads, in_features))
        if in_edge_features is not None:
            self.attne = nn.Parameter(torch.empty(1, num_heads, in_features))
            self.fc_fij = nn.Linear(in_edge_features, in_features * num_heads, bias=bias)
        self.in_edge_feats = in_edge_features
        self.leaky_relu = nn.LeakyReLU(negative_slope)
        self.attndrop = nn.Dropout(dropout)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        for lin in (self.linear, self.el, self.er) + ((self.fc_fij,) if self.in_edge_feats is not None else ()):
            nn.init.xavier_normal_(lin.weight, gain=gain)
            if lin.bias is not None:
                nn.init.constant_(lin.bias, 0.0)
        nn.init.xavier_normal_(self.attnl, gain=gain)
        nn.init.xavier_normal_(self.attnr, gain=gain)
        if self.in_edge_feats is not None:       # the reference leaves attne uninitialised (gnn.py:168); any value works
            nn.init.xavier_normal_(self.attne, gain=gain)

    def forward(self, features, edge_index, edge_features=None):
        prefix = features.shape[:-1]
        fc_src = self.el(features).view(*prefix, self.num_heads, self.in_features)
        fc_dst = self.er(features).view(*prefix, self.num_heads, self.in_features)
        el = (fc_src[edge_index[0]] * self.attnl).sum(dim=-1).unsqueeze(-1)
        er = (fc_dst[edge_index[1]] * self.attnr).sum(dim=-1).unsqueeze(-1)
        if edge_features is not None and self.in_edge_feats is not None:
            fc_eij = self.fc_fij(edge_features).view(*edge_features.shape[:-1], self.num_heads, self.in_features)
            e = self.leaky_relu(el + er + (fc_eij * self.attne).sum(dim=-1).unsqueeze(-1))
        else:
            e = self.leaky_relu(el + er)
        a = torch.softmax(self.attndrop(e), dim=1).mean(dim=1)
        h = self.linear(features)
        return h.clone().index_add_(0, edge_index[0], a * h[edge_index[1]])


class OnsetEmbedding(nn.Module):
    """analysisgnn/models/core/gnn.py:294-311: ``W((x_i + sum_j |x_i - x_j|) / max(deg_i, 1))`` with self loops
    appended first (they add nothing to the sum and one to the divisor)."""

    def __init__(self, in_feats, out_feats, bias=True, add_self_loops=True):
        super().__init__()
        self.W = nn.Linear(in_feats, out_feats, bias=bias)
        self.add_self_loops = add_self_loops

    def forward(self, x, edge_index):
        if self.add_self_loops:
            loops = torch.arange(0, x.size(0), dtype=torch.long, device=x.device).unsqueeze(0).repeat(2, 1)
            edge_index = torch.cat([edge_index, loops], dim=1)
        msg = torch.abs(x[edge_index[0]] - x[edge_index[1]])
        return self.W(mean_into_copy(msg, edge_index[0], x))


_REDUCTIONS = {
    "mean": lambda t: t.mean(dim=0),
    "sum": lambda t: t.sum(dim=0),
}


class HeteroConv(nn.Module):
    """analysisgnn/models/core/hgnn.py:435-484 (the relation buffer is kept on the
    input's device here; the reference allocates it on the CPU)."""

    def __init__(self, in_features, out_features, etypes, in_edge_features=None,
                 module=SageConvScatter, bias=True, reduction="mean"):
        super().__init__()
        self.out_features = out_features
        self.etypes = etypes
        if reduction not in _REDUCTIONS:
            raise NotImplementedError(reduction)
        self.reduction = _REDUCTIONS[reduction]
        self.conv = nn.ModuleDict({
            name: module(in_features, out_features, bias=bias, in_edge_features=in_edge_features)
            for name in etypes
        })

    def reset_parameters(self):
        for conv in self.conv.values():
            conv.reset_parameters()

    def forward(self, x, edge_index, edge_type, edge_features=None):
        per_rel = []
        for name, code in self.etypes.items():
            pick = edge_type == code
            ef = edge_features[pick, :] if edge_features is not None else None
            per_rel.append(self.conv[name](x, edge_index[:, pick], ef))
        return self.reduction(torch.stack(per_rel, dim=0))


class MetricalConvLayer(nn.Module):
    """analysisgnn/models/core/gnn.py:488-540.  ``lengths``: None (one sequence),
    a cumulative pointer when ragged, or equal per-graph counts when uniform."""

    def __init__(self, in_dim, out_dim, activation=None, dropout=0.2, bias=True):
        super().__init__()
        self.input_dim = in_dim
        self.output_dim = out_dim
        self.activation = nn.Identity() if activation is None else activation
        self.dropout = nn.Dropout(dropout)
        self.normalize = nn.BatchNorm1d(out_dim)
        self.neigh = nn.Linear(in_dim, in_dim, bias=bias)
        self.conv_out = nn.Linear(4 * in_dim, out_dim, bias=bias)
        self.seq = nn.GRU(in_dim, in_dim, batch_first=True, bias=bias, bidirectional=True)

    def reset_parameters(self):
        self.neigh.reset_parameters()
        self.conv_out.reset_parameters()
        self.seq.reset_parameters()

    def forward(self, x_metrical, x, edge_index, lengths):
        n_m = x_metrical.size(0)
        if lengths is None:
            lengths = torch.tensor([n_m], dtype=torch.long, device=x_metrical.device)
        ragged = not bool(torch.all(lengths == lengths[0]))
        gathered = sum_into(self.neigh(x)[edge_index[0]], edge_index[1], n_m)     # gnn.py:510-511
        both = torch.cat((gathered, x_metrical), dim=-1)
        if ragged:
            sizes = torch.diff(lengths).tolist()
            both_seq = pad_sequence(torch.split(both, sizes), batch_first=True)
            gath_seq = pad_sequence(torch.split(gathered, sizes), batch_first=True)
        else:
            t = int(lengths[0])
            gath_seq = gathered.view(-1, t, gathered.shape[1])
            both_seq = both.view(-1, t, both.shape[1])
        rec = self.seq(gath_seq)[0]
        h = self.activation(self.conv_out(torch.cat((both_seq, rec), dim=-1)))
        h = self.normalize(h.transpose(1, 2))        # BatchNorm over [B, C, T], pads included (gnn.py:526-531)
        h = self.dropout(h).transpose(1, 2)
        if ragged:
            steps = torch.arange(h.shape[1], device=h.device).unsqueeze(0)
            valid = steps < torch.diff(lengths).unsqueeze(1)
            h = h[valid].view(-1, h.shape[-1])
        else:
            h = h.reshape(-1, h.shape[-1])
        out = sum_into(h[edge_index[1]], edge_index[0], x.size(0))                # gnn.py:539
        return out, h


class MetricalGNN(nn.Module):
    """analysisgnn/models/core/hgnn.py:323-433 (``jk`` unsupported: the reference
    builds it with ``n_layers=hidden_features`` and it cannot run)."""

    def __init__(self, input_features, hidden_features, output_features, etypes, num_layers=2, dropout=0.5,
                 use_reledge=False, jk=False, in_edge_features=None, metrical=False, conv_block=SageConvScatter):
        super().__init__()
        if jk:
            raise NotImplementedError("jk=True is unusable in the reference (hgnn.py:340)")
        self.dropout = dropout
        self.num_layers = num_layers
        self.num_hidden = hidden_features
        self.use_reledge = use_reledge
        self.use_metrical = metrical
        self.convs = nn.ModuleList()
        self.emb_beats = nn.Linear(input_features, hidden_features)
        self.emb_measures = nn.Linear(input_features, hidden_features)
        self.beat_convs = nn.ModuleList()
        self.measure_convs = nn.ModuleList()
        self.project_metrical = nn.ModuleList()
        first_edge = in_edge_features if use_reledge else None
        self.convs.append(HeteroConv(input_features, hidden_features, etypes=etypes,
                                     in_edge_features=first_edge, module=conv_block))
        for _ in range(max(num_layers - 2, 0)):
            self.convs.append(HeteroConv(hidden_features, hidden_features, etypes=etypes, module=conv_block))
            if metrical:
                self._add_metrical(hidden_features, hidden_features, dropout)
        self.convs.append(HeteroConv(hidden_features, hidden_features, etypes=etypes, module=conv_block))
        if metrical:
            self._add_metrical(hidden_features, output_features, dropout)

    def _add_metrical(self, h_in, h_out, dropout):
        self.beat_convs.append(MetricalConvLayer(h_in, h_out, activation=F.relu, dropout=dropout))
        self.measure_convs.append(MetricalConvLayer(h_in, h_out, activation=F.relu, dropout=dropout))
        self.project_metrical.append(nn.Linear(h_out * 3, h_out))

    def _metrical_step(self, k, h, h_beat, h_measure, beat_edges, measure_edges, beat_lengths, measure_lengths):
        from_beats, h_beat = self.beat_convs[k](h_beat, h, beat_edges, beat_lengths)
        from_measures, h_measure = self.measure_convs[k](h_measure, h, measure_edges, measure_lengths)
        h = self.project_metrical[k](torch.cat((h, from_beats, from_measures), dim=-1))
        return F.normalize(F.relu(h), p=2, dim=-1), h_beat, h_measure

    def forward(self, x, edge_index, edge_type, beat_nodes=None, measure_nodes=None, beat_edges=None,
                measure_edges=None, rel_edge=None, beat_lengths=None, measure_lengths=None, **kwargs):
        h_beat = h_measure = None
        if self.use_metrical:                                                   # hgnn.py:405-407
            h_beat = sum_into(self.emb_beats(x)[beat_edges[0]], beat_edges[1], beat_nodes.size(0))
            h_measure = sum_into(self.emb_measures(x)[measure_edges[0]], measure_edges[1], measure_nodes.size(0))
        h = x
        for i in range(len(self.convs) - 1):
            if i != 0 and self.use_metrical:
                h, h_beat, h_measure = self._metrical_step(i - 1, h, h_beat, h_measure, beat_edges, measure_edges,
                                                           beat_lengths, measure_lengths)
            if i == 0 and self.use_reledge:
                h = self.convs[i](h, edge_index, edge_type, edge_features=rel_edge)
            else:
                h = self.convs[i](h, edge_index, edge_type)
            h = F.relu(F.normalize(h, p=2, dim=-1))                              # normalize, then relu (hgnn.py:421-422)
            h = F.dropout(h, p=self.dropout, training=self.training)
        if self.use_metrical:
            h, h_beat, h_measure = self._metrical_step(-1, h, h_beat, h_measure, beat_edges, measure_edges,
                                                       beat_lengths, measure_lengths)
        return self.convs[-1](h, edge_index, edge_type)


def onset_pool(x, onset_edges, batch_size):
    """analysisgnn/models/analysis.py:580-586."""
    inside = (onset_edges[0] < batch_size) & (onset_edges[1] < batch_size)
    e = onset_edges[:, inside]
    e = e[:, e[0] != e[1]]
    return mean_into_copy(x[e[1]], e[0], x)
