"""Fused projection stages: ``[LayerNorm ->] [Dropout ->] Linear [-> ReLU]`` for a GROUP of independent inputs.

Every dense block of the reference's encoder shell is a chain of such stages -- ``project_dict`` per node type,
``project_enc``, the task heads ``clf_dict`` (analysisgnn/models/analysis.py:429-443, 474-496) and the sequence MLP
(analysisgnn/models/cadence.py:252-260).  Run as separate modules a stage costs a LayerNorm kernel that writes fp32,
a dropout kernel, an amax pass and a split pass that rewrite the same matrix as the GEMM operand, the GEMM, and a
ReLU kernel -- and the same again backwards.  Here a stage is:

* forward: ``agnn_layernorm_fwd_pair`` writes the (dropped-out) normalised rows DIRECTLY as the fp16 hi / lo operand
  pair (scale from the LayerNorm bound, dropout from a counter RNG), ``agnn_gemm_grouped`` runs the projections of all
  group members (node types, task heads) in ONE launch with bias + ReLU in the epilogue, which also reports max |y| for
  whoever consumes y as an operand next;
* backward: ``agnn_grad_prepare`` (ReLU mask from the saved output, operand pair, bias gradient) with the amax its
  producer tagged, one grouped launch for all weight gradients (split-K reduced inside the launch), one for all input
  gradients, ``agnn_layernorm_bwd_dropout`` (mask recomputed, max |dx| reported).

Groups whose members are too small for the fp16 operand form (tests, tiny batches) take the same route with TF32
operand pairs and an fp32 LayerNorm output.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib, linalg, ops


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _ln_forward(x, gamma, beta, eps, p, site, want_pair: bool):
    """-> (y fp32 or None, SplitH or None, mean, rstd)."""
    rows, cols = x.shape
    dev = x.device
    mean = torch.empty(rows, dtype=torch.float32, device=dev)
    rstd = torch.empty(rows, dtype=torch.float32, device=dev)
    y = pair = None
    hi = lo = amax = None
    if want_pair:
        buf = torch.empty((2, rows, cols), dtype=torch.float16, device=dev)
        hi, lo = buf[0], buf[1]
        amax = linalg.new_amax(dev)
    else:
        y = torch.empty((rows, cols), dtype=torch.float32, device=dev)
    rng = linalg.dropout_state(dev) if p > 0 else None
    _lib.check(_lib.lib().agnn_layernorm_fwd_pair(
        x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(), y.data_ptr() if y is not None else None,
        y.stride(0) if y is not None else 0, mean.data_ptr(), rstd.data_ptr(), rows, cols, float(eps),
        hi.data_ptr() if hi is not None else None, lo.data_ptr() if lo is not None else None, cols,
        amax.data_ptr() if amax is not None else None, float(p), rng.data_ptr() if rng is not None else None,
        int(site), _stream(x)), "agnn_layernorm_fwd_pair")
    _lib.count_launches(1)
    if want_pair:
        pair = linalg.SplitH(hi, lo, amax)
    return y, pair, mean, rstd


def _ln_backward(dy, x, gamma, mean, rstd, p, site, amax_out):
    rows, cols = x.shape
    lib = _lib.lib()
    dy = dy if dy.stride(1) == 1 and dy.stride(0) % 4 == 0 and dy.data_ptr() % 16 == 0 else dy.contiguous()
    dx = torch.empty_like(x)
    blocks = lib.agnn_row_blocks(rows)
    part = torch.empty((2, blocks, cols), dtype=torch.float32, device=x.device)
    sums = torch.empty((2, cols), dtype=torch.float32, device=x.device)
    rng = linalg.dropout_state(x.device) if p > 0 else None
    _lib.check(lib.agnn_layernorm_bwd_dropout(
        dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
        dx.data_ptr(), dx.stride(0), part[0].data_ptr(), part[1].data_ptr(), sums[0].data_ptr(), sums[1].data_ptr(),
        rows, cols, float(p), rng.data_ptr() if rng is not None else None, int(site),
        amax_out.data_ptr() if amax_out is not None else None, _stream(x)), "agnn_layernorm_bwd_dropout")
    _lib.count_launches(2)
    return dx, sums[0], sums[1]


def dropout_apply(x, p, site, amax_out=None):
    """``keep ? x / (1 - p) : 0`` with the counter mask of (site, element): the same call undoes nothing -- it is its own
    backward (the mask is a function of the call site and the element index only)."""
    x = x if x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0 else x.contiguous()
    y = torch.empty((x.shape[0], x.shape[1]), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().agnn_dropout_apply(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), x.shape[0],
                                             x.shape[1], float(p), linalg.dropout_state(x.device).data_ptr(), int(site),
                                             amax_out.data_ptr() if amax_out is not None else None, _stream(x)),
               "agnn_dropout_apply")
    _lib.count_launches(1)
    return y


def _row_ok(x):
    return x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1 and x.stride(0) % 4 == 0 and \
        x.data_ptr() % 16 == 0


def _ln_shape_ok(cols):
    return cols % 4 == 0 and 4 <= cols <= 1024


class _StageGroup(torch.autograd.Function):
    """Arguments: (meta, *tensors).  ``meta``: one dict per member {ln, eps, p, site, relu, bias, amax_out}; tensors
    per member, flattened: x, weight, [bias], [gamma, beta]."""

    @staticmethod
    def forward(ctx, meta, *tensors):
        members, pos = [], 0
        for m in meta:
            x, w = tensors[pos], tensors[pos + 1]
            pos += 2
            b = g = be = None
            if m["bias"]:
                b = tensors[pos]
                pos += 1
            if m["ln"]:
                g, be = tensors[pos], tensors[pos + 1]
                pos += 2
            members.append([x if x.is_contiguous() else x.contiguous(), w, b, g, be])
        live = [i for i, mb in enumerate(members) if mb[0].shape[0] > 0]
        # operand form of the whole group: fp16 pairs if the shapes allow it and the group is big enough to care
        f16 = (linalg.parity_operands() == "f16" and bool(live)
               and max(members[i][0].shape[0] for i in live) >= linalg.F16_MIN_ROWS
               and all(_row_ok(members[i][0]) and members[i][0].shape[1] % 8 == 0 and linalg.f16_ok(members[i][1])
                       and (not meta[i]["ln"] or members[i][0].shape[1] <= 1024) for i in live))
        operands, saved, specs, shared = [None] * len(meta), [], [], {}
        for i, (m, (x, w, b, g, be)) in enumerate(zip(meta, members)):
            mean = rstd = None
            if x.shape[0] == 0:
                operands[i] = None
            elif m["ln"]:
                if not (_row_ok(x) and _ln_shape_ok(x.shape[1])):
                    raise _lib.AgnnError("fused stage: LayerNorm width must be a multiple of 4 up to 1024 (fp32)")
                y, pair, mean, rstd = _ln_forward(x, g.contiguous(), be.contiguous(), m["eps"], m["p"], m["site"], f16)
                operands[i] = pair if f16 else linalg.prepare(y)
            else:
                key = (x.data_ptr(), tuple(x.shape), x.stride(0), m["p"], m["site"])
                if key in shared:                                  # one input feeding several heads: one operand
                    operands[i] = shared[key]
                elif f16:
                    operands[i] = linalg.split_f16(x, m.get("x_amax"), dropout=(m["p"], m["site"]) if m["p"] > 0 else None)
                else:
                    y = dropout_apply(x, m["p"], m["site"]) if m["p"] > 0 else x
                    operands[i] = linalg.prepare(y)
                shared[key] = operands[i]
            saved.append((mean, rstd))
            if x.shape[0] > 0:
                specs.append(dict(a=operands[i], b=w, m=x.shape[0], n=w.shape[0], k=x.shape[1], bias=b,
                                  flags=_lib.GEMM_RELU if m["relu"] else 0, amax_out=m.get("amax_out")))
        res = iter(linalg._group(_lib.K_MAJOR, _lib.K_MAJOR, specs))
        outs = []
        for i, (x, w, b, g, be) in enumerate(members):
            outs.append(next(res) if x.shape[0] > 0 else x.new_zeros((0, w.shape[0])))
        keep = []
        for i, (m, (x, w, b, g, be)) in enumerate(zip(meta, members)):
            op = operands[i]
            first, second = linalg.pack(op) if op is not None else (None, None)
            keep += [x if m["ln"] else None, w, g, saved[i][0], saved[i][1], first, second,
                     outs[i] if m["relu"] else None]
        ctx.save_for_backward(*keep)
        ctx.meta = meta
        ctx.op_amax = [getattr(op, "amax", None) for op in operands]
        ctx.has = [(m["bias"], m["ln"]) for m in meta]
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        meta = ctx.meta
        sv = ctx.saved_tensors
        n_mem = len(meta)
        # argument positions (for needs_input_grad): 0 = meta, then the flattened tensors
        arg_pos, pos = [], 1
        for m in meta:
            d = {"x": pos, "w": pos + 1}
            pos += 2
            if m["bias"]:
                d["b"] = pos
                pos += 1
            if m["ln"]:
                d["g"], d["be"] = pos, pos + 1
                pos += 2
            arg_pos.append(d)
        grads = [None] * pos
        act, gs_list, ops_list, w_list, want_dx = [], [], [], [], []
        for i, m in enumerate(meta):
            x, w, gam, mean, rstd, first, second, out = sv[8 * i:8 * i + 8]
            g = gouts[i]
            if g is None or first is None:
                continue
            operand = linalg.unpack(first, second, ctx.op_amax[i])
            f16 = isinstance(operand, linalg.SplitH)
            n = g.shape[1]
            mult = 8 if f16 else 4
            wp = w
            if n % mult:                                           # e.g. the 185- and 50-class heads
                if m["relu"]:
                    raise _lib.AgnnError("fused stage: ReLU outputs need a width that is a multiple of 8")
                base = getattr(g, "_agnn_padded", None)            # producer wrote an aligned, zero-padded buffer
                if base is not None and base.shape[1] % mult == 0 and base.shape[1] >= n:
                    am = linalg.known_amax(g)
                    g = linalg.tag_amax(base, am)
                else:
                    g = torch.nn.functional.pad(g, (0, mult - n % mult))
                wp = torch.nn.functional.pad(w, (0, 0, 0, g.shape[1] - n))
            want_db = m["bias"] and ctx.needs_input_grad[arg_pos[i]["b"]]
            gs, db = ops.prepare_grad(g, out, want_db, f16=f16, amax=linalg.known_amax(g))
            if db is not None:
                grads[arg_pos[i]["b"]] = db[:n]
            need_dx = ctx.needs_input_grad[arg_pos[i]["x"]] or m["ln"]
            act.append((i, n, x, gam, mean, rstd))
            gs_list.append(gs)
            ops_list.append(operand)
            w_list.append(wp)
            want_dx.append(need_dx)
        if not act:
            return tuple(grads)
        # all weight gradients in one grouped launch (split-K reduced inside it), all input gradients in another
        need_w = [ctx.needs_input_grad[arg_pos[i]["w"]] for (i, *_rest) in act]
        sel = [j for j, f in enumerate(need_w) if f]
        dws = linalg.mm_tn_group([gs_list[j] for j in sel], [ops_list[j] for j in sel]) if sel else []
        for j, dw in zip(sel, dws):
            i, n = act[j][0], act[j][1]
            grads[arg_pos[i]["w"]] = dw[:n]
        sel = [j for j, f in enumerate(want_dx) if f]
        amaxes = [linalg.new_amax(gs_list[j].hi.device if hasattr(gs_list[j], "hi") else gs_list[j].device) for j in sel]
        dops = linalg.mm_group([gs_list[j] for j in sel], [w_list[j] for j in sel], amax_outs=amaxes) if sel else []
        for j, dop, am in zip(sel, dops, amaxes):
            i, n, x, gam, mean, rstd = act[j]
            m = meta[i]
            if m["ln"]:
                out_am = linalg.new_amax(dop.device)
                dx, dgam, dbe = _ln_backward(dop, x, gam.contiguous(), mean, rstd, m["p"], m["site"], out_am)
                grads[arg_pos[i]["g"]], grads[arg_pos[i]["be"]] = dgam, dbe
                if ctx.needs_input_grad[arg_pos[i]["x"]]:
                    grads[arg_pos[i]["x"]] = linalg.tag_amax(dx, out_am)
            elif m["p"] > 0:
                out_am = linalg.new_amax(dop.device)
                grads[arg_pos[i]["x"]] = linalg.tag_amax(dropout_apply(dop, m["p"], m["site"], out_am), out_am)
            else:
                grads[arg_pos[i]["x"]] = linalg.tag_amax(dop, am)
        return tuple(grads)


class _ConcatCols(torch.autograd.Function):
    """``torch.cat(parts, dim=1)`` whose backward hands the column slices of the gradient on WITH the gradient's amax
    tag (a bound of the whole matrix bounds every slice), so the projections below need no amax pass."""

    @staticmethod
    def forward(ctx, *parts):
        ctx.widths = [p.shape[1] for p in parts]
        return torch.cat(parts, dim=1)

    @staticmethod
    def backward(ctx, g):
        am = linalg.known_amax(g)
        out, c0 = [], 0
        for w in ctx.widths:
            out.append(linalg.tag_amax(g[:, c0:c0 + w], am))
            c0 += w
        return tuple(out)


def concat_cols(parts):
    return _ConcatCols.apply(*parts)


def stage_group(xs: Sequence[torch.Tensor], weights, biases, norms=None, relu=False, dropout: float = 0.0,
                training: bool = False, share_amax: bool = False) -> List[torch.Tensor]:
    """``[Linear_i(Dropout(LayerNorm_i(x_i)))]`` (LayerNorm / Dropout / ReLU optional) for independent members.

    ``norms``: per member ``None`` or ``(gamma, beta, eps)``; ``relu`` / ``dropout`` apply to every member; inputs of
    any leading shape are flattened to rows; feature counts that are not multiples of 8 (the 25 + 128 note features,
    analysis.py:574) are zero-padded.  ``share_amax``: all members report max |y| into ONE device scalar (the inputs of
    a message-passing layer share one operand scale)."""
    n = len(xs)
    norms = norms if norms is not None else [None] * n
    p = float(dropout) if training else 0.0
    if p >= 0.75:
        raise ValueError("fused dropout supports p < 0.75")
    meta, flat, leads = [], [], []
    shared = None
    for x, w, b, nm in zip(xs, weights, biases, norms):
        if not x.is_cuda:
            raise _lib.AgnnError("analysisgnn_b200 has no CPU path: tensors must live on a CUDA device")
        leads.append(x.shape[:-1])
        # (no reshape node for matrices: a view node in the autograd graph hands the gradient on as a NEW tensor
        # object, which loses the amax tag its producer attached)
        x2 = x if x.dim() == 2 else x.reshape(-1, x.shape[-1])
        am_in = linalg.known_amax(x) if nm is None else None
        k = x2.shape[1]
        if x2.dtype == torch.float32 and k % 8 and nm is None:
            pad = 8 - k % 8
            x2 = torch.nn.functional.pad(x2, (0, pad))
            w = torch.nn.functional.pad(w, (0, pad))
        if share_amax:
            shared = shared if shared is not None else linalg.new_amax(x.device)
            am = shared
        else:
            am = linalg.new_amax(x.device)
        meta.append({"ln": nm is not None, "eps": nm[2] if nm is not None else 0.0, "p": p,
                     "site": linalg.dropout_site() if p > 0 else 0, "relu": bool(relu), "bias": b is not None,
                     "amax_out": am, "x_amax": am_in})
        flat += [x2, w] + ([b] if b is not None else []) + ([nm[0], nm[1]] if nm is not None else [])
    outs = _StageGroup.apply(meta, *flat)
    res = []
    for o, m, lead, w in zip(outs, meta, leads, weights):
        linalg.tag_amax(o, m["amax_out"])
        res.append(o if len(lead) == 1 else linalg.tag_amax(o.reshape(*lead, w.shape[0]), m["amax_out"]))
    return res
