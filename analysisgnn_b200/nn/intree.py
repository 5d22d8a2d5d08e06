"""CUDA-backed mirrors of the reference's in-tree message-passing layers.

Replaces (same names / ctor / forward / state_dict keys):

* ``SageConvScatter``    analysisgnn/models/core/gnn.py:39-76
* ``HeteroConv``         analysisgnn/models/core/hgnn.py:435-484
* ``MetricalConvLayer``  analysisgnn/models/core/gnn.py:488-540
* ``MetricalGNN``        analysisgnn/models/core/hgnn.py:323-433
* ``ResGatedGraphConv`` / ``RelEdgeConv`` / ``GATConvLayer`` / ``OnsetEmbedding``  gnn.py:212-258, 79-106, 154-209, 294-311

Convention of these layers: reduce at ``edge_index[0]`` reading ``edge_index[1]``.
A ``HeteroConv`` layer is ONE fused launch chain for all relations -- a grouped
projection, one relation-fused gather kernel on the dst-sorted CSR, one output
projection -- instead of the reference's per-relation mask / gather / 4-kernel
scatter loop through a CPU buffer (hgnn.py:480-484).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import graph, ops
from .layers import GRU, Linear


def _xavier_relu_(linear: nn.Linear):
    nn.init.xavier_uniform_(linear.weight, gain=nn.init.calculate_gain("relu"))
    if linear.bias is not None:
        nn.init.zeros_(linear.bias)


class SageConvScatter(nn.Module):
    """``z = W [x || s] + b``, ``s_i = (x_i + sum_{e: ei[0,e]=i} (Wn x + bn)[ei[1,e]]) / max(deg_i, 1)``;
    with no edges at all ``z = W [x || Wn x + bn] + b`` (gnn.py:62-76)."""

    def __init__(self, in_features, out_features, bias=True, in_edge_features=None):
        super().__init__()
        self.neigh_linear = Linear(in_features, in_features, bias=bias)
        self.linear = Linear(in_features * 2, out_features, bias=bias)
        self.in_edge_features = in_edge_features
        if in_edge_features is not None:
            self.edge_linear = Linear(in_edge_features, in_features, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        _xavier_relu_(self.linear)
        _xavier_relu_(self.neigh_linear)
        if self.in_edge_features is not None:
            _xavier_relu_(self.edge_linear)

    def forward(self, features, edge_index, edge_features=None, neigh_feats=None):
        n = features.shape[0]
        if edge_index is None:
            edge_index = torch.zeros((2, 0), dtype=torch.long, device=features.device)
        use_edge = self.in_edge_features is not None and edge_features is not None and edge_index.shape[1] > 0
        if neigh_feats is None and not use_edge:
            csr = graph.typed_csr(edge_index, None, n, 1)
            return ops.intree_sage_layer(features, self.neigh_linear.weight, self.neigh_linear.bias,
                                         self.linear.weight, self.linear.bias, csr)
        # rarely used options (edge features / separate neighbour features): per-edge messages,
        # still reduced by the segmented kernel (edge-ordered CSR, no atomics)
        h = self.neigh_linear(features if neigh_feats is None else neigh_feats)
        if edge_index.shape[1] == 0:
            return self.linear(torch.cat((features, h), dim=-1))
        msg = h.index_select(0, edge_index[1])
        if use_edge:
            msg = msg + self.edge_linear(edge_features)
        s = ops.segment_mean_self(msg, features, graph.edge_csr(edge_index[0], n))
        return self.linear(torch.cat((features, s), dim=-1))


class ResGatedGraphConv(nn.Module):
    """analysisgnn/models/core/gnn.py:212-258: ``h1 + (h1 + sum_j sigmoid(W3 x_i + W4 x_j [+ W5 e]) * W2 x_j)``
    (``h1`` enters twice in the reference, :256-257).  Projections on the tensor-core GEMM; gate, product and
    reduction in ONE kernel per direction (agnn_edge_op: the messages live in registers, no [E, F] tensor).  With
    edge features (``W5 e``, one row per edge by construction) the gate is formed per edge as in the reference."""

    def __init__(self, in_features, out_features, bias=True, in_edge_features=None):
        super().__init__()
        self.W1 = Linear(in_features, out_features, bias=bias)
        self.W2 = Linear(in_features, out_features, bias=bias)
        self.W3 = Linear(in_features, out_features, bias=bias)
        self.W4 = Linear(in_features, out_features, bias=bias)
        self.in_edge_features = in_edge_features
        if in_edge_features is not None:
            self.W5 = Linear(in_edge_features, out_features, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        for name in ("W1", "W2", "W3", "W4", "W5"):
            if hasattr(self, name):
                _xavier_relu_(getattr(self, name))

    def forward(self, features, edge_index, edge_features=None, neigh_feats=None):
        h1 = self.W1(features)
        h2 = self.W2(features if neigh_feats is None else neigh_feats)
        if edge_index is None or edge_index.shape[1] == 0:
            return h1 + h1
        if edge_features is not None and self.in_edge_features is not None:
            gate = self.W3(features).index_select(0, edge_index[0]) + self.W4(features).index_select(0, edge_index[1])
            gate = gate + self.W5(edge_features)
            msg = torch.sigmoid(gate) * h2.index_select(0, edge_index[1])
            s = ops.segment_sum_self(msg, h1, graph.edge_csr(edge_index[0], features.shape[0]))
            return h1 + s
        n = features.shape[0]
        csr = graph.typed_csr(edge_index, None, n, 1, n_cols=h2.shape[0])
        return h1 + ops.edge_gate_sum(self.W3(features), self.W4(features), h2, csr, self_add=h1)


class RelEdgeConv(nn.Module):
    """analysisgnn/models/core/gnn.py:79-106: messages ``edge_linear([h_j || |h_i - h_j|])`` (or given edge
    features), mean into ``h.clone()``, then ``linear([x || s])``.  Without given edge features the per-edge projection
    is linear in quantities that can be reduced FIRST: ``sum_j edge_linear([h_j || d_ij]) = edge_linear_w [sum_j h_j ||
    sum_j d_ij] + deg_i b`` -- two segmented reductions (the second forms ``|h_i - h_j|`` in registers, agnn_edge_op)
    and a projection over N rows instead of E (E ~ 4.6 N), no [E, 2F] tensor; changes the fp32 summation order only."""

    def __init__(self, in_node_features, out_features, bias=True, in_edge_features=None):
        super().__init__()
        self.neigh_linear = Linear(in_node_features, in_node_features, bias=bias)
        self.edge_linear = Linear(in_node_features * 2 if in_edge_features is None
                                  else in_node_features + in_edge_features, in_node_features, bias=bias)
        self.linear = Linear(in_node_features * 2, out_features, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        for lin in (self.linear, self.neigh_linear, self.edge_linear):
            _xavier_relu_(lin)

    def forward(self, features, edge_index, edge_features=None):
        h = self.neigh_linear(features)
        if edge_features is not None:
            hj = h.index_select(0, edge_index[1])
            msg = self.edge_linear(torch.cat((hj, edge_features), dim=-1))
            s = ops.segment_mean_self(msg, h, graph.edge_csr(edge_index[0], features.shape[0]))
            return self.linear(torch.cat((features, s), dim=-1))
        n = features.shape[0]
        csr = graph.typed_csr(edge_index, None, n, 1)
        s1 = ops.segment_sum(h, csr)                                   # sum_j h_j
        s2 = ops.edge_absdiff(h, h, csr)                               # sum_j |h_i - h_j|
        deg = graph.derived(edge_index, ("degree", n), lambda: torch.bincount(edge_index[0], minlength=n).to(h.dtype))
        msg = ops.linear(torch.cat((s1, s2), dim=-1), self.edge_linear.weight, None)
        if self.edge_linear.bias is not None:
            msg = msg + deg.unsqueeze(1) * self.edge_linear.bias
        s = (h + msg) / deg.clamp(min=1).unsqueeze(1)
        return self.linear(torch.cat((features, s), dim=-1))


class GATConvLayer(nn.Module):
    """analysisgnn/models/core/gnn.py:154-209.  The reference normalises the attention scores over the HEADS
    (``Softmax(dim=1)``) and then averages over the heads (:206): the weight of every edge is the constant
    ``1 / num_heads`` whatever the scores and the attention dropout are, so the layer is
    ``h_i + (1 / num_heads) sum_{e: ei[0,e]=i} h_{ei[1,e]}`` with ``h = linear(x)`` -- one projection and one segmented
    sum with the self term.  ``el`` / ``er`` / ``attnl`` / ``attnr`` (``attne`` / ``fc_fij``) are kept for the
    reference's ``state_dict`` layout; the reference's gradients for them are the rounding residue of
    ``1 - sum(softmax)`` (1e-8 of the ``linear`` gradients), here they stay ``None``."""

    def __init__(self, in_features, out_features, num_heads=3, bias=True, dropout=0.3, negative_slope=0.2,
                 in_edge_features=None):
        super().__init__()
        self.num_heads, self.in_features, self.out_features = num_heads, in_features, out_features
        self.linear = Linear(in_features, out_features, bias=bias)
        self.el = nn.Linear(in_features, in_features * num_heads, bias=bias)
        self.er = nn.Linear(in_features, in_features * num_heads, bias=bias)
        self.attnl = nn.Parameter(torch.empty(1, num_heads, in_features))
        self.attnr = nn.Parameter(torch.empty(1, num_heads, in_features))
        if in_edge_features is not None:
            self.attne = nn.Parameter(torch.empty(1, num_heads, in_features))
            self.fc_fij = nn.Linear(in_edge_features, in_features * num_heads, bias=bias)
        self.in_edge_feats = in_edge_features
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        for lin in (self.linear, self.el, self.er) + ((self.fc_fij,) if self.in_edge_feats is not None else ()):
            nn.init.xavier_normal_(lin.weight, gain=gain)
            if lin.bias is not None:
                nn.init.constant_(lin.bias, 0.0)
        nn.init.xavier_normal_(self.attnl, gain=gain)
        nn.init.xavier_normal_(self.attnr, gain=gain)
        if self.in_edge_feats is not None:
            nn.init.xavier_normal_(self.attne, gain=gain)

    def forward(self, features, edge_index, edge_features=None):
        h = self.linear(features)
        if edge_index is None or edge_index.shape[1] == 0:
            return h
        csr = graph.typed_csr(edge_index, None, features.shape[0], 1)
        return ops.segment_sum_self(h * (1.0 / self.num_heads), h, csr)


class OnsetEmbedding(nn.Module):
    """analysisgnn/models/core/gnn.py:294-311: ``W((x_i + sum_{e: ei[0,e]=i} |x_i - x_{ei[1,e]}|) / max(deg_i, 1))``;
    the self loops the reference appends add nothing to the sum and one to every divisor."""

    def __init__(self, in_feats, out_feats, bias=True, add_self_loops=True):
        super().__init__()
        self.W = Linear(in_feats, out_feats, bias=bias)
        self.add_self_loops = add_self_loops

    def forward(self, x, edge_index):
        if self.add_self_loops:
            loops = torch.arange(0, x.size(0), dtype=torch.long, device=x.device).unsqueeze(0).repeat(2, 1)
            edge_index = torch.cat([edge_index, loops], dim=1)
        csr = graph.typed_csr(edge_index, None, x.shape[0], 1)
        return self.W(ops.edge_absdiff(x, x, csr, self_add=x, mean=True))   # |x_i - x_j| formed in registers


_FOLDABLE = ("mean", "sum")


class HeteroConv(nn.Module):
    """Per-relation convolution + reduction over relations (hgnn.py:435-484).

    ``etypes`` maps relation name -> integer code found in ``edge_type``.  With
    ``module=SageConvScatter`` and ``reduction`` in {mean, sum} the whole layer is
    fused (the reduction is linear, so it is folded into the output projection).
    """

    def __init__(self, in_features, out_features, etypes, in_edge_features=None, module=SageConvScatter,
                 bias=True, reduction="mean"):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.etypes = etypes
        if reduction not in ("mean", "sum", "concat"):
            # max / min return a (values, indices) tuple in the reference and cannot run there;
            # 'lstm' (HeteroAttention) is outside the hot path
            raise NotImplementedError(f"reduction={reduction!r}")
        self.reduction = reduction
        self.conv = nn.ModuleDict({name: module(in_features, out_features, bias=bias,
                                                in_edge_features=in_edge_features) for name in etypes})
        self._lut = None
        self.reset_parameters()

    def reset_parameters(self):
        for conv in self.conv.values():
            conv.reset_parameters()

    # relation index of every edge: code -> position in ``etypes`` (-1 = not a relation of this layer)
    def _relation_ids(self, edge_type):
        codes = list(self.etypes.values())
        if codes == list(range(len(codes))):
            return edge_type
        if self._lut is None or self._lut.device != edge_type.device:
            lut = torch.full((max(codes) + 2,), -1, dtype=torch.long)
            for k, c in enumerate(codes):
                lut[c] = k
            self._lut = lut.to(edge_type.device)
        top = self._lut.numel() - 1                      # the extra last slot maps to -1
        return graph.derived(edge_type, ("lut", tuple(codes)), lambda: self._lut[
            torch.where((edge_type >= 0) & (edge_type < top), edge_type, torch.full_like(edge_type, top))])

    def _fusable(self, edge_features):
        return (self.reduction in _FOLDABLE and edge_features is None
                and all(type(c) is SageConvScatter for c in self.conv.values()))

    def forward(self, x, edge_index, edge_type, edge_features=None):
        if edge_type is None:
            raise ValueError("Edge type must be specified")
        names = list(self.etypes.keys())
        r = len(names)
        if self._fusable(edge_features):
            f = self.in_features
            convs = [self.conv[k] for k in names]
            csr = graph.typed_csr(edge_index, self._relation_ids(edge_type), x.shape[0], r)
            has_bias = convs[0].neigh_linear.bias is not None
            wn_cat = torch.cat([c.neigh_linear.weight for c in convs], dim=0)                  # [R*F, F]
            bn_cat = torch.cat([c.neigh_linear.bias for c in convs], dim=0) if has_bias else None
            scale = 1.0 / r if self.reduction == "mean" else 1.0
            w_self = torch.stack([c.linear.weight[:, :f] for c in convs], dim=0).sum(0)
            wc = torch.cat([w_self] + [c.linear.weight[:, f:] for c in convs], dim=1) * scale  # [F', (R+1)F]
            bc = torch.stack([c.linear.bias for c in convs], dim=0).sum(0) * scale if has_bias else None
            return ops.intree_sage_layer(x, wn_cat, bn_cat, wc, bc, csr)
        # generic path (other conv blocks / edge features / concat): per relation, device-resident
        rel = self._relation_ids(edge_type)
        outs = []
        for k, name in enumerate(names):
            pick = rel == k
            ef = edge_features[pick, :] if edge_features is not None else None
            outs.append(self.conv[name](x, edge_index[:, pick], ef))
        if self.reduction == "concat":
            return torch.cat(outs, dim=0)
        out = torch.stack(outs, dim=0)
        return out.mean(dim=0) if self.reduction == "mean" else out.sum(dim=0)


class MetricalConvLayer(nn.Module):
    """note -> beat/measure segmented sum, biGRU over each score's metrical sequence,
    BatchNorm (padded positions included, gnn.py:526-531), and the segmented sum back
    to the notes (gnn.py:488-540).  Both scatters run on the one CSR pair of the
    note -> metrical edges (forward graph and its transpose)."""

    def __init__(self, in_dim, out_dim, activation=None, dropout=0.2, bias=True):
        super().__init__()
        self.input_dim = in_dim
        self.output_dim = out_dim
        self.activation = nn.Identity() if activation is None else activation
        self.dropout = nn.Dropout(dropout)
        self.normalize = nn.BatchNorm1d(out_dim)
        self.neigh = Linear(in_dim, in_dim, bias=bias)
        self.conv_out = Linear(4 * in_dim, out_dim, bias=bias)
        self.seq = GRU(in_dim, in_dim, batch_first=True, bias=bias, bidirectional=True)

    def reset_parameters(self):
        self.neigh.reset_parameters()
        self.conv_out.reset_parameters()
        self.seq.reset_parameters()

    def forward(self, x_metrical, x, edge_index, lengths):
        n_m, n = x_metrical.size(0), x.size(0)
        layout = graph.sequence_layout(lengths, n_m)
        csr = graph.typed_csr(edge_index, None, n_m, 1, reduce_row=1, n_cols=n)   # rows = metrical nodes
        gathered = ops.segment_sum(self.neigh(x), csr)                            # gnn.py:510-511
        both = torch.cat((gathered, x_metrical), dim=-1)
        gath_seq, both_seq = layout.pad(gathered), layout.pad(both)
        rec = self.seq(gath_seq)[0]
        h = self.activation(self.conv_out(torch.cat((both_seq, rec), dim=-1)))
        h = self.dropout(self.normalize(h.transpose(1, 2))).transpose(1, 2)
        h = layout.unpad(h)
        out = ops.segment_sum(h, csr.t())                                         # gnn.py:539
        return out, h


class MetricalGNN(nn.Module):
    """hgnn.py:323-433.  ``jk=True`` cannot run in the reference (JumpingKnowledge is
    built with ``n_layers=hidden_features``, hgnn.py:340) and is rejected here."""

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
        self.use_knowledge = False
        self.convs = nn.ModuleList()
        self.emb_beats = Linear(input_features, hidden_features)
        self.emb_measures = Linear(input_features, hidden_features)
        self.beat_convs = nn.ModuleList()
        self.measure_convs = nn.ModuleList()
        self.project_metrical = nn.ModuleList()
        self.convs.append(HeteroConv(input_features, hidden_features, etypes=etypes,
                                     in_edge_features=in_edge_features if use_reledge else None, module=conv_block))
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
        self.project_metrical.append(Linear(h_out * 3, h_out))

    def reset_parameters(self):
        for group in (self.convs, self.beat_convs, self.measure_convs, self.project_metrical):
            for m in group:
                m.reset_parameters()
        self.emb_beats.reset_parameters()
        self.emb_measures.reset_parameters()

    # The beat and the measure layer of a metrical step are independent, and each is a chain of small launches (a GRU
    # time loop of ~120 / ~30 steps, segmented sums, BatchNorm): the measure layer runs on the side stream beside the
    # beat layer -- in the backward too, autograd replays every node on the stream of its forward.
    overlap_metrical = True

    def _metrical_step(self, k, h, h_beat, h_measure, beat_edges, measure_edges, beat_lengths, measure_lengths):
        side = ops.side_stream(h.device) if (self.overlap_metrical and h.is_cuda) else None
        if side is not None:
            main = torch.cuda.current_stream(h.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                from_measures, h_measure = self.measure_convs[k](h_measure, h, measure_edges, measure_lengths)
        from_beats, h_beat = self.beat_convs[k](h_beat, h, beat_edges, beat_lengths)
        if side is not None:
            main.wait_stream(side)
            from_measures.record_stream(main)
            h_measure.record_stream(main)
        else:
            from_measures, h_measure = self.measure_convs[k](h_measure, h, measure_edges, measure_lengths)
        h = self.project_metrical[k](torch.cat((h, from_beats, from_measures), dim=-1))
        return ops.l2norm_relu(h, relu_first=True), h_beat, h_measure

    def forward(self, x, edge_index, edge_type, beat_nodes=None, measure_nodes=None, beat_edges=None,
                measure_edges=None, rel_edge=None, beat_lengths=None, measure_lengths=None, **kwargs):
        h_beat = h_measure = None
        if self.use_metrical:                                                    # hgnn.py:405-407
            n = x.size(0)
            b_csr = graph.typed_csr(beat_edges, None, beat_nodes.size(0), 1, reduce_row=1, n_cols=n)
            m_csr = graph.typed_csr(measure_edges, None, measure_nodes.size(0), 1, reduce_row=1, n_cols=n)
            h_beat = ops.segment_sum(self.emb_beats(x), b_csr)
            h_measure = ops.segment_sum(self.emb_measures(x), m_csr)
        h = x
        for i in range(len(self.convs) - 1):
            if i != 0 and self.use_metrical:
                h, h_beat, h_measure = self._metrical_step(i - 1, h, h_beat, h_measure, beat_edges, measure_edges,
                                                           beat_lengths, measure_lengths)
            if i == 0 and self.use_reledge:
                h = self.convs[i](h, edge_index, edge_type, edge_features=rel_edge)
            else:
                h = self.convs[i](h, edge_index, edge_type)
            h = ops.l2norm_relu(h, relu_first=False)                              # normalize, then relu (hgnn.py:421-422)
            h = F.dropout(h, p=self.dropout, training=self.training)
        if self.use_metrical:
            h, h_beat, h_measure = self._metrical_step(-1, h, h_beat, h_measure, beat_edges, measure_edges,
                                                       beat_lengths, measure_lengths)
        return self.convs[-1](h, edge_index, edge_type)
