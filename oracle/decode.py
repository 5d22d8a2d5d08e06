"""CPU restatement of the reference's onset-wise logit aggregation + decode.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PINNED: compared with the reference's own
``onsetwise_logit_aggregation`` (analysisgnn/models/analysis.py:44-101) executed through
``oracle/ref_loader.load_onsetwise_decode`` and with the golden vectors that run produced
(``tests/golden/decode_*.pt``).

What the reference does, quirks included:
* the onset mean is written INTO the caller's logit tensors (``scatter_mean(..., out=v)``, :66) with the
  self term counted in the sum but not in the divisor, then softmax; the valid-label selection applies a
  second softmax (:68);
* for a single-score batch the rows of every run of equal onsets are represented by the run's first row
  (:79-82); where the arg-max of consecutive onsets changes a segment starts, and every note whose onset
  lies in segment i gets the distribution of the segment's first onset -- except in the LAST segment, which
  the loop never reaches (:96-99).
"""
from types import SimpleNamespace

import torch

RNA_KEYS = ("quality", "inversion", "degree1", "degree2")


def note_store(x, batch, onset_div, edge_index_dict):
    """Minimal stand-in for the HeteroData the reference passes as ``graph``."""
    class Graph(dict):
        pass
    g = Graph(note=SimpleNamespace(x=x, batch=batch, onset_div=onset_div))
    g.edge_index_dict = edge_index_dict
    return g


def _scatter_mean_self_(v, src_idx, dst_idx):
    """torch_scatter.scatter_mean(v[src], dst, dim=0, out=v): in place, divisor = max(#edges into the row, 1)."""
    gathered = v[src_idx]
    count = torch.zeros(v.shape[0], dtype=v.dtype)
    count.scatter_add_(0, dst_idx, torch.ones_like(dst_idx, dtype=v.dtype))
    v.scatter_add_(0, dst_idx.view(-1, 1).expand_as(gathered), gathered)
    v.div_(count.clamp_(min=1).view(-1, 1))
    return v


def onsetwise_logit_aggregation(logits_softmax_dict, graph, edge_index_dict=None, batch_size=None,
                                valid_label_mask=None, rna_keys=RNA_KEYS):
    rna_keys = list(rna_keys)
    if not (rna_keys and all(k in logits_softmax_dict for k in rna_keys)):      # :45
        return logits_softmax_dict
    note = graph["note"]
    batch_size = len(note.x) if batch_size is None else batch_size
    edge_index_dict = graph.edge_index_dict if edge_index_dict is None else edge_index_dict
    if valid_label_mask is None:
        valid_label_mask = torch.ones(batch_size, dtype=torch.bool)
    e = edge_index_dict["note", "onset", "note"]                                   # :50-55
    e = e[:, (e[0] < batch_size) & (e[1] < batch_size)]
    e = e[:, e[0] != e[1]]
    tpc = None
    if "tpc_in_label" in logits_softmax_dict:                                       # :57-59
        tpc = logits_softmax_dict["tpc_in_label"].argmax(-1).bool()
        e = e[:, tpc[e[0]] & tpc[e[1]]]
    agg = {}
    for k, v in logits_softmax_dict.items():                                        # :64-66
        if k in rna_keys:
            agg[k] = _scatter_mean_self_(v, e[0], e[1]).softmax(-1)
    agg = {k: v[valid_label_mask].softmax(-1) for k, v in agg.items()}              # :68
    logits_softmax_dict.update(agg)
    batch_id = note.batch[:batch_size][valid_label_mask]
    if torch.all(batch_id == batch_id[0]):                                          # :71
        onsets = note.onset_div[:batch_size][valid_label_mask]
        hold_between_change_points(logits_softmax_dict, onsets, tpc, rna_keys)
    return logits_softmax_dict


def hold_between_change_points(dists, onsets, tpc=None, rna_keys=RNA_KEYS):
    """The single-score stage (:72-99) on its own: ``dists[k]`` are the per-note distributions after the two
    softmaxes; rows are overwritten in place.  Only comparisons, arg-max and row copies happen here, so given the
    same ``dists`` every implementation must agree bit for bit."""
    onsets = onsets - onsets.min()
    agg = {k: dists[k] for k in rna_keys}
    if tpc is not None:
        onsets_f = onsets[tpc]
        agg = {k: v[tpc] for k, v in agg.items()}
    else:
        onsets_f = onsets
    uniq, inv = torch.unique(onsets_f, return_inverse=True)                         # :79
    heads = (inv[1:] != inv[:-1]).nonzero(as_tuple=True)[0] + 1
    heads = torch.cat([torch.zeros(1, dtype=heads.dtype), heads])
    onsetwise = {k: v[heads] for k, v in agg.items()}
    for k in rna_keys:                                                              # :85-99
        pred = onsetwise[k].argmax(-1)
        cp = (pred[1:] != pred[:-1]).nonzero(as_tuple=True)[0] + 1
        cp = torch.cat([torch.zeros(1, dtype=cp.dtype), cp])
        onset_at = uniq[cp]
        rows = onsetwise[k][cp]
        for i in range(len(cp) - 1):
            m = (onset_at[i] <= onsets) & (onsets < onset_at[i + 1])
            dists[k][m] = rows[i]
    return dists
