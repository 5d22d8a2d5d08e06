"""CUDA-backed PyG / graphmuse shaped encoders.

The reference builds its production encoder from the third-party graphmuse package
(``HybridGNN`` / ``HybridHGT`` / ``MetricalGNN``, analysisgnn/models/analysis.py:9,
444-473) on top of torch_geometric's ``HeteroConv`` / ``SAGEConv`` / ``HGTConv``.
Neither is part of the reference tree; the wiring below follows the reference's
in-tree statement of the same stack (analysisgnn/models/cadence.py:142-176,
229-332) and the published PyG >= 2.3 operator semantics, as fixed by this
repo's oracle (oracle/pyg.py).  Constructor / forward
arguments are those of the call site analysisgnn/models/analysis.py:444-473,
576-579; parameter names follow PyG (``convs.<i>.convs.<src>__<rel>__<dst>.lin_l``).

Convention (PyG): ``edge_index[0]`` = source j, ``edge_index[1]`` = target i.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import graph, ops
from .layers import GRU, MLP, LayerNorm, Linear

EdgeType = Tuple[str, str, str]


def rel_key(edge_type) -> str:
    return "__".join(edge_type)


class SAGEConv(nn.Module):
    """Parameter holder of one ``SAGEConv(in, out, aggr='mean', root_weight=True)``:
    ``lin_l(mean_j x_j) + lin_r(x_i)``.  The arithmetic runs in ``HeteroSAGELayer``
    for all relations at once; calling this module alone runs the same fused path
    for a single relation."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = Linear(in_channels, out_channels, bias=True)
        self.lin_r = Linear(in_channels, out_channels, bias=False)

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()

    def forward(self, x_src, x_dst, edge_index):
        same = x_src is x_dst
        types = ("n", "n") if same else ("s", "d")
        et = (types[0], "to", types[1])
        plan = LayerPlan([et], [et])
        x_dict = {"n": x_dst} if same else {"s": x_src, "d": x_dst}
        csr = graph.hetero_csr({et: edge_index}, {k: v.shape[0] for k, v in x_dict.items()})
        wcat = torch.cat((self.lin_r.weight, self.lin_l.weight), dim=1)
        xs = [x_dict[t] for t in plan.node_types]
        return ops.hetero_sage_layer(plan, csr, False, xs, [wcat, self.lin_l.bias])[0]


class LayerPlan:
    """Static wiring of one hetero layer for the edge types present in a batch."""

    def __init__(self, layer_edge_types: List[EdgeType], present: List[EdgeType], only_dst=None):
        present_set = set(present)
        self.edge_types = [et for et in layer_edge_types if et in present_set
                           and (only_dst is None or et[2] in only_dst)]
        self.incoming: Dict[str, List[EdgeType]] = {}
        self.outgoing: Dict[str, List[EdgeType]] = {}
        for et in self.edge_types:
            self.incoming.setdefault(et[2], []).append(et)
            self.outgoing.setdefault(et[0], []).append(et)
        self.dst_types = list(self.incoming.keys())
        seen = []
        for et in self.edge_types:
            for t in (et[0], et[2]):
                if t not in seen:
                    seen.append(t)
        self.node_types = seen
        for t in self.node_types:
            self.outgoing.setdefault(t, [])


class HeteroSAGELayer(nn.Module):
    """PyG ``HeteroConv({et: SAGEConv(in, out)}, aggr)`` as one gather launch + one
    GEMM per destination node type (ops.hetero_sage_layer)."""

    def __init__(self, edge_types, in_channels, out_channels, aggr="sum"):
        super().__init__()
        if aggr not in ("sum", "mean"):
            raise NotImplementedError(f"aggr={aggr!r}")
        self.edge_types = [tuple(et) for et in edge_types]
        self.aggr = aggr
        self.in_channels, self.out_channels = in_channels, out_channels
        self.convs = nn.ModuleDict({rel_key(et): SAGEConv(in_channels, out_channels) for et in self.edge_types})
        self._plans = {}

    def reset_parameters(self):
        for c in self.convs.values():
            c.reset_parameters()

    def plan_for(self, x_dict, ei_dict, only_dst=None) -> LayerPlan:
        present = tuple(et for et in self.edge_types if et in ei_dict and et[0] in x_dict and et[2] in x_dict)
        key = (present, None if only_dst is None else tuple(only_dst))
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = LayerPlan(self.edge_types, list(present), only_dst)
        return plan

    def forward(self, x_dict, ei_dict, csr: Optional[graph.HeteroCSR] = None, relu: bool = False, only_dst=None):
        """``only_dst``: compute these destination types only (the caller consumes nothing else)."""
        plan = self.plan_for(x_dict, ei_dict, only_dst)
        if csr is None:
            csr = graph.hetero_csr({et: ei_dict[et] for et in plan.edge_types},
                                   {t: x_dict[t].shape[0] for t in plan.node_types})
        params = []
        for t in plan.dst_types:
            convs = [self.convs[rel_key(et)] for et in plan.incoming[t]]
            scale = 1.0 / len(convs) if self.aggr == "mean" else 1.0
            wcat, bias = ops.sage_weights([c.lin_r.weight for c in convs], [c.lin_l.weight for c in convs],
                                          [c.lin_l.bias for c in convs], scale)
            params += [wcat, bias]
        outs = ops.hetero_sage_layer(plan, csr, relu, [x_dict[t] for t in plan.node_types], params)
        return dict(zip(plan.dst_types, outs))


def trim_inputs(layer, nodes_per_hop, edges_per_hop, x_dict, ei_dict):
    """``torch_geometric.utils.trim_to_layer`` for dict inputs, one step (as called in
    cadence.py:166-173): drop the outermost remaining hop's nodes and edges."""
    if layer <= 0 or edges_per_hop is None:
        return x_dict, ei_dict
    x_dict = {k: v[: v.size(0) - nodes_per_hop[k][-layer]] for k, v in x_dict.items()}
    ei_dict = {k: v[:, : v.size(1) - edges_per_hop[k][-layer]] for k, v in ei_dict.items()}
    return x_dict, ei_dict


def _layer_structures(conv_layers, x_dict, ei_dict, nodes_per_hop, edges_per_hop):
    """One ``HeteroCSR`` per layer (trimmed edge lists differ per layer), built in a
    single batched agnn_csr_build pass and cached per batch."""
    sizes = {t: v.shape[0] for t, v in x_dict.items()}
    if not (isinstance(ei_dict, graph.TypedEdgeDict) and edges_per_hop is None
            and all(et[0] in sizes and et[2] in sizes for et in ei_dict.keys())):
        ei_dict = {et: ei_dict[et] for et in ei_dict.keys() if et[0] in sizes and et[2] in sizes}
    ets = list(ei_dict.keys())
    if edges_per_hop is None:
        csr = graph.hetero_csr(ei_dict, sizes)
        return [csr] * len(conv_layers)
    out = []
    n_cur = dict(sizes)
    e_cur = {et: ei_dict[et].shape[1] for et in ets}
    for i in range(len(conv_layers)):
        if i > 0:
            n_cur = {t: n_cur[t] - nodes_per_hop[t][-i] for t in n_cur}
            e_cur = {et: e_cur[et] - edges_per_hop[et][-i] for et in ets}
        # slicing keeps the parent's storage, so the cache key uses (parent, length)
        out.append(graph.hetero_csr_trimmed(ei_dict, dict(e_cur), dict(n_cur)))
    return out


class HeteroSAGEStack(nn.Module):
    """cadence.py:142-176: ``trim_to_layer -> HeteroConv{SAGEConv}(aggr='sum') -> relu``, L times."""

    def __init__(self, edge_types, in_channels, hidden_channels, num_layers, aggr="sum"):
        super().__init__()
        self.convs = nn.ModuleList(
            HeteroSAGELayer(edge_types, in_channels if i == 0 else hidden_channels, hidden_channels, aggr)
            for i in range(num_layers))

    def forward(self, x_dict, ei_dict, nodes_per_hop=None, edges_per_hop=None, collect=None, final_types=None):
        """``final_types``: node types the caller reads from the result; the last layer skips the rest
        (PyG computes and discards them)."""
        structures = _layer_structures(self.convs, x_dict, ei_dict, nodes_per_hop, edges_per_hop)
        last = len(self.convs) - 1
        for i, conv in enumerate(self.convs):
            x_dict, ei_dict = trim_inputs(i, nodes_per_hop, edges_per_hop, x_dict, ei_dict)
            x_dict = conv(x_dict, ei_dict, csr=structures[i], relu=True, only_dst=final_types if i == last else None)
            if collect is not None:
                collect.append(x_dict)
        return x_dict


# ------------------------------------------------------------------------ HGT

class HGTConv(nn.Module):
    """PyG >= 2.3 ``HGTConv(in, out, metadata, heads)`` on the fused attention kernel
    (agnn_hgt_attn_fwd / _bwd): per-node-type KQV projection, per-(relation, head)
    ``k_rel`` / ``v_rel`` applied to the source type's k, v, ``p_rel`` scaling, softmax
    over ALL incoming edges of a target across relations (``joint_softmax=True``; the
    pre-2.3 per-relation softmax with ``False``), ``out_lin(gelu(.))``, gated skip."""

    def __init__(self, in_channels, out_channels, metadata, heads=1, joint_softmax=True):
        super().__init__()
        if out_channels % heads:
            raise ValueError("out_channels must be divisible by heads")
        self.node_types = list(metadata[0])
        self.edge_types = [tuple(et) for et in metadata[1]]
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.joint_softmax = joint_softmax
        d = out_channels // heads
        self.kqv_lin = nn.ModuleDict({t: Linear(in_channels, 3 * out_channels) for t in self.node_types})
        self.out_lin = nn.ModuleDict({t: Linear(out_channels, out_channels) for t in self.node_types})
        n_rel = len(self.edge_types)
        self.k_rel = nn.Parameter(torch.empty(heads * n_rel, d, d))   # index = head * n_rel + relation
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

    def _rel_ids(self, ids, device):
        """Device copy of a relation-id list (made once: a host->device copy cannot be graph-captured)."""
        key = (tuple(ids), device)
        cache = self.__dict__.setdefault("_rel_id_cache", {})
        if key not in cache:
            cache[key] = torch.tensor(ids, dtype=torch.long, device=device)
        return cache[key]

    def _plan(self, node_types, edge_types, only_dst):
        """Which relations run and which node types project what -- a function of the KEYS of the batch only."""
        present = [et for et in self.edge_types if et in edge_types and et[0] in node_types and et[2] in node_types]
        by_dst: Dict[str, List[EdgeType]] = {}
        for et in present:
            if only_dst is None or et[2] in only_dst:
                by_dst.setdefault(et[2], []).append(et)
        return present, by_dst

    def projection_weights(self, node_types, edge_types, only_dst, device):
        """One wide projection per node type: q (if the type is a target) and, for every relation leaving the type, k
        and v with the relation's per-head k_rel / v_rel folded into the weight,
            (x Wk^T + bk) blockdiag(k_rel[:, r]) = x (blockdiag^T Wk)^T + bk blockdiag,
        so all per-node-type and per-relation projections of the layer are one tensor-core GEMM per type.  The result
        depends on parameters only (~190 tiny launches per layer, autograd-tracked): ``HeteroHGTStack`` builds it for
        all layers on a side stream while the main stream is busy with the batch."""
        H, D = self.heads, self.out_channels // self.heads
        hd = H * D
        n_rel = len(self.edge_types)
        _, by_dst = self._plan(node_types, edge_types, only_dst)
        wk_all = self.k_rel.view(H, n_rel, D, D)
        wv_all = self.v_rel.view(H, n_rel, D, D)
        types, proj_w, proj_b, q_off, kv_off = [], [], [], {}, {}
        for t in node_types:
            out_rels = [et for ets in by_dst.values() for et in ets if et[0] == t]
            if t not in by_dst and not out_rels:
                continue
            w, b = self.kqv_lin[t].weight, self.kqv_lin[t].bias
            parts_w, parts_b, width = [], [], 0
            if t in by_dst:
                parts_w.append(w[hd:2 * hd])
                parts_b.append(b[hd:2 * hd])
                q_off[t], width = 0, hd
            if out_rels:
                ids = [self.edge_types.index(et) for et in out_rels]
                n_out = len(ids)
                for which, rel_w, rows in ((0, wk_all, slice(0, hd)), (1, wv_all, slice(2 * hd, 3 * hd))):
                    sel = rel_w.index_select(1, self._rel_ids(ids, device))               # [H, R_t, D, D]
                    parts_w.append(torch.einsum("hrde,hdi->rhei", sel, w[rows].view(H, D, -1)).reshape(n_out * hd, -1))
                    parts_b.append(torch.einsum("hrde,hd->rhe", sel, b[rows].view(H, D)).reshape(n_out * hd))
                    for i, et in enumerate(out_rels):
                        kv_off[(et, which)] = width + i * hd
                    width += n_out * hd
            types.append(t)
            proj_w.append(torch.cat(parts_w, dim=0))
            proj_b.append(torch.cat(parts_b, dim=0))
        return types, proj_w, proj_b, q_off, kv_off

    def forward(self, x_dict, ei_dict, csr: Optional[graph.HeteroCSR] = None, only_dst=None, weights=None):
        """``weights``: the result of ``projection_weights`` for this batch's keys, when the caller built it ahead."""
        H, D = self.heads, self.out_channels // self.heads
        hd = H * D
        present, by_dst = self._plan(list(x_dict.keys()), ei_dict.keys(), only_dst)
        if csr is None:
            csr = graph.hetero_csr({et: ei_dict[et] for et in present}, {t: v.shape[0] for t, v in x_dict.items()})
        if weights is None:
            weights = self.projection_weights(list(x_dict.keys()), ei_dict.keys(), only_dst,
                                              next(iter(x_dict.values())).device)
        types, proj_w, proj_b, q_off, kv_off = weights
        slot = {t: i for i, t in enumerate(types)}
        proj_x = [x_dict[t] for t in types]
        from .. import fused
        # the wide projections of all node types: one grouped GEMM launch (and one per direction in the backward)
        ys = fused.stage_group(proj_x, proj_w, proj_b) if proj_x else []
        scale = 1.0 / math.sqrt(D)
        targets, pscales, order = [], [], []
        for dst, ets in by_dst.items():
            rels = [(slot[et[0]], kv_off[(et, 0)], kv_off[(et, 1)], csr.fwd[et], csr.bwd[et]) for et in ets]
            ps = torch.stack([self.p_rel[rel_key(et)].reshape(H) * scale for et in ets], dim=0)
            if self.joint_softmax:
                groups = [ops.HgtGroup(rels)]
                pscales.append(ps)
            else:
                groups = [ops.HgtGroup([r]) for r in rels]
                pscales += [ps[i:i + 1] for i in range(len(rels))]
            targets.append(ops.HgtTarget(slot[dst], q_off[dst], groups))
            order.append(dst)
        aggs = ops.hgt_layer_attention(ys, targets, pscales, H, hd) if targets else ()
        out_dict = {}
        outs = fused.stage_group([F.gelu(agg) for agg in aggs], [self.out_lin[dst].weight for dst in order],
                                 [self.out_lin[dst].bias for dst in order]) if order else []
        for dst, o in zip(order, outs):
            if o.size(-1) == x_dict[dst].size(-1):
                # a * o + (1 - a) * x as one kernel (x + a (o - x)): the gate is a learned scalar per node type
                o = torch.lerp(x_dict[dst], o, self.skip[dst].sigmoid())
            out_dict[dst] = o
        return out_dict


class HeteroHGTStack(nn.Module):
    prefetch_weights = True

    def __init__(self, metadata, in_channels, hidden_channels, num_layers, heads, dropout=0.0, joint_softmax=True):
        super().__init__()
        self.dropout = dropout
        self.convs = nn.ModuleList(
            HGTConv(in_channels if i == 0 else hidden_channels, hidden_channels, metadata, heads, joint_softmax)
            for i in range(num_layers))

    def forward(self, x_dict, ei_dict, nodes_per_hop=None, edges_per_hop=None, collect=None, final_types=None):
        structures = _layer_structures(self.convs, x_dict, ei_dict, nodes_per_hop, edges_per_hop)
        last = len(self.convs) - 1
        # the composite projection weights of every layer depend on parameters only: ~570 tiny launches per step that
        # run on their own stream beside the batch's work (their backward too: autograd replays a node on its
        # forward's stream); a layer waits for its own set
        dev = next(iter(x_dict.values())).device
        ahead = [None] * len(self.convs)
        if self.prefetch_weights and dev.type == "cuda":
            main, side = torch.cuda.current_stream(dev), ops.side_stream(dev, 1)
            side.wait_stream(main)
            node_types, edge_types = list(x_dict.keys()), list(ei_dict.keys())
            with torch.cuda.stream(side):
                for i, conv in enumerate(self.convs):
                    w = conv.projection_weights(node_types, edge_types, final_types if i == last else None, dev)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    ahead[i] = (w, ev)
        for i, conv in enumerate(self.convs):
            x_dict, ei_dict = trim_inputs(i, nodes_per_hop, edges_per_hop, x_dict, ei_dict)
            weights = None
            if ahead[i] is not None and list(x_dict.keys()) == node_types and list(ei_dict.keys()) == edge_types:
                weights, ev = ahead[i]
                main.wait_event(ev)
                for t in weights[1] + weights[2]:
                    t.record_stream(main)
            out = conv(x_dict, ei_dict, csr=structures[i], only_dst=final_types if i == last else None, weights=weights)
            x_dict = {k: F.dropout(v.relu(), self.dropout, self.training) for k, v in out.items()}
            if collect is not None:
                collect.append(x_dict)
        return x_dict


# ------------------------------------------------------------ hybrid encoders

class SequenceBranch(nn.Module):
    """cadence.py:249-260, 276-285: split by graph -> pad -> 2-layer biGRU (cuDNN) ->
    LayerNorm -> MLP -> unpad.  Runs on a side stream, concurrently with the
    message-passing stack (it depends on the encoder input only)."""

    def __init__(self, in_channels, hidden_channels, dropout):
        super().__init__()
        self.rnn = GRU(input_size=in_channels, hidden_size=hidden_channels // 2, num_layers=2,
                          batch_first=True, bidirectional=True, dropout=dropout)
        self.rnn_norm = LayerNorm(hidden_channels)
        self.rnn_mlp = MLP(
            Linear(hidden_channels, hidden_channels), nn.ReLU(), LayerNorm(hidden_channels),
            nn.Dropout(dropout), Linear(hidden_channels, hidden_channels))

    def forward(self, x, batch):
        from .. import linalg
        layout = graph.batch_layout(batch)
        seq = linalg.tag_amax(layout.pad(x), linalg.known_amax(x))
        seq = self.rnn(seq)[0]
        seq = self.rnn_mlp(seq, pre_norm=self.rnn_norm)          # LayerNorm fused into the first projection's operand
        return linalg.tag_amax(layout.unpad(seq), linalg.known_amax(seq))


class JumpingKnowledge(nn.Module):
    """analysisgnn/models/core/gnn.py:345-365 (LSTM attention over layer outputs)."""

    def __init__(self, n_hidden, n_layers):
        super().__init__()
        self.lstm = nn.LSTM(n_hidden, (n_layers * n_hidden) // 2, bidirectional=True, batch_first=True)
        self.att = Linear(2 * ((n_layers * n_hidden) // 2), 1)

    def forward(self, xs):
        x = torch.stack(xs, dim=1)
        alpha, _ = self.lstm(x)
        alpha = torch.softmax(self.att(alpha).squeeze(-1), dim=-1)
        return (x * alpha.unsqueeze(-1)).sum(dim=1)


def _head(x, n):
    """``x[:n]`` without an autograd slice (zeros + copy in the backward) when it keeps every row."""
    return x if x.shape[0] == n else x[:n]


class _HybridBase(nn.Module):
    overlap_sequence_branch = True
    # The sequence branch is a chain of four ~0.8 ms recurrence kernels on a fraction of the SMs, the graph branch a
    # row of wide kernels that take every free SM.  On a high-priority stream the chain's CTAs do not queue behind
    # them.  Measured on the 100 x 500-note step (profiles/r2_u_*): HybridGNN 8.91 -> 8.73 ms; HybridHGT, whose graph
    # branch is the longer chain, 16.87 -> 17.32 ms -- hence per encoder.  AGNN_SEQ_PRIORITY=0 / 1 overrides both.
    sequence_branch_high_priority = False

    def forward(self, x_dict, edge_index_dict, batch_dict=None, batch_size=None, neighbor_mask_node=None,
                neighbor_mask_edge=None, return_edge_index=False, edge_attr_dict=None):
        x_in = x_dict["note"]
        batch_size = x_in.size(0) if batch_size is None else batch_size
        batch = batch_dict["note"][:batch_size] if batch_dict is not None else \
            torch.zeros(batch_size, dtype=torch.long, device=x_in.device)
        main = torch.cuda.current_stream(x_in.device)
        prio = os.environ.get("AGNN_SEQ_PRIORITY", "")
        prio = self.sequence_branch_high_priority if prio == "" else prio != "0"
        side = ops.side_stream(x_in.device, 0, prio) if self.overlap_sequence_branch else None
        if side is not None:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                x_seq = self.seq(_head(x_in, batch_size), batch)
        collect = [] if self.use_jk else None
        out = self.gnn(x_dict, edge_index_dict, neighbor_mask_node, neighbor_mask_edge, collect, ("note",))
        if self.use_jk:
            x_gnn = self.jk([_head(c["note"], batch_size) for c in collect])
        else:
            x_gnn = _head(out["note"], batch_size)
        if side is not None:
            main.wait_stream(side)
            x_seq.record_stream(main)
        else:
            x_seq = self.seq(_head(x_in, batch_size), batch)
        from .. import fused
        return self.cat_proj(fused.concat_cols((x_gnn, x_seq)))


class HybridGNN(_HybridBase):
    """graphmuse ``HybridGNN(metadata, input_channels, hidden_channels, num_layers, dropout,
    use_jk)`` (call site analysisgnn/models/analysis.py:454-462): HeteroSAGEStack on the
    graph + GRU branch over the target notes, joined by ``Linear(2H, H)`` (cadence.py:301-303)."""

    sequence_branch_high_priority = True

    def __init__(self, metadata, input_channels, hidden_channels, num_layers, dropout=0.5, use_jk=False):
        super().__init__()
        self.use_jk = use_jk
        self.gnn = HeteroSAGEStack(metadata[1], input_channels, hidden_channels, num_layers)
        self.seq = SequenceBranch(input_channels, hidden_channels, dropout)
        self.cat_proj = Linear(2 * hidden_channels, hidden_channels)
        if use_jk:
            self.jk = JumpingKnowledge(hidden_channels, num_layers)


class HybridHGT(_HybridBase):
    """graphmuse ``HybridHGT`` (call site analysisgnn/models/analysis.py:444-453)."""

    def __init__(self, metadata, input_channels, hidden_channels, num_layers, heads=4, dropout=0.5, use_jk=False,
                 joint_softmax=True):
        super().__init__()
        self.use_jk = use_jk
        self.gnn = HeteroHGTStack(metadata, input_channels, hidden_channels, num_layers, heads, dropout, joint_softmax)
        self.seq = SequenceBranch(input_channels, hidden_channels, dropout)
        self.cat_proj = Linear(2 * hidden_channels, hidden_channels)
        if use_jk:
            self.jk = JumpingKnowledge(hidden_channels, num_layers)


class MetricalGNN(nn.Module):
    """graphmuse-flavoured ``MetricalGNN(metadata, input_channels, hidden_channels,
    output_channels, num_layers, dropout, use_jk, fast)`` (call shape cadence.py:232-234,
    298-300; analysisgnn/models/analysis.py:463-473): HeteroSAGEStack -> note rows -> MLP.
    Returns all (trimmed) note rows; callers slice ``[:batch_size]``."""

    def __init__(self, metadata, input_channels, hidden_channels, output_channels, num_layers, dropout=0.5,
                 use_jk=False, fast=True):
        super().__init__()
        self.gnn = HeteroSAGEStack(metadata[1], input_channels, hidden_channels, num_layers)
        self.mlp = MLP(
            Linear(hidden_channels, hidden_channels), nn.ReLU(), LayerNorm(hidden_channels),
            nn.Dropout(dropout), Linear(hidden_channels, output_channels))

    def forward(self, x_dict, edge_index_dict, neighbor_mask_node=None, neighbor_mask_edge=None, **kwargs):
        out = self.gnn(x_dict, edge_index_dict, neighbor_mask_node, neighbor_mask_edge, None, ("note",))
        return self.mlp(out["note"])
