"""Thin Python wrappers over the C ABI + the autograd Functions built from them.

Tensors cross the boundary as raw device pointers (``tensor.data_ptr()``) on the
current CUDA stream; nothing here allocates inside the library or synchronises.
Dense contractions go through ``linalg`` (this repo's GEMM entry points).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import torch

from . import _lib, linalg
from .graph import CSR, HeteroCSR, TypedCSR

_DTYPES = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported feature dtype {t.dtype} (float32 and bfloat16 only)") from None


def _rows2d(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{what} must be a 2-D view with unit column stride")
    if not t.is_cuda:
        raise _lib.AgnnError(f"{what} must be a CUDA tensor: analysisgnn_b200 has no CPU path")
    return t


@dataclass
class Rel:
    """One relation of a fused gather launch (mirrors agnn_rel_t)."""
    rowptr: torch.Tensor                     # int32 [n_rows + 1]
    col: torch.Tensor                        # int32, base that rowptr values index into
    src: torch.Tensor                        # [n_src, F] view, unit column stride
    out_col: int = 0
    nbr_deg_rowptr: Optional[torch.Tensor] = None
    flags: int = 0
    n_edges: Optional[int] = None            # edges of this relation (accounting, heavy-row workspace size)
    heavy_rows: Optional[torch.Tensor] = None    # rows with >= HEAVY_ROW entries (CSR.heavy[k]) ...
    n_heavy: Optional[torch.Tensor] = None       # ... and their device-side count (CSR.n_heavy[k:k+1])


def _pack(rels: Sequence[Rel], n_feat: int, dtype) -> "ctypes.Array":
    if not 1 <= len(rels) <= _lib.MAX_REL:
        raise ValueError(f"1..{_lib.MAX_REL} relations per launch, got {len(rels)}")
    arr = (_lib.Rel * len(rels))()
    for i, r in enumerate(rels):
        src = _rows2d(r.src, "src")
        if src.dtype != dtype or src.shape[1] != n_feat:
            raise ValueError("all gathered matrices must share dtype and feature count with the output")
        arr[i].rowptr = r.rowptr.data_ptr()
        arr[i].col = r.col.data_ptr()
        arr[i].src = src.data_ptr()
        arr[i].ld_src = src.stride(0)
        arr[i].nbr_deg_rowptr = r.nbr_deg_rowptr.data_ptr() if r.nbr_deg_rowptr is not None else None
        arr[i].out_col = int(r.out_col)
        arr[i].flags = int(r.flags)
        if (len(rels) == 1 and r.flags == 0 and r.n_edges is not None and r.nbr_deg_rowptr is None
                and int(r.n_edges) < 1.5 * (int(r.rowptr.numel()) - 1)):
            arr[i].flags = _lib.REL_LOW_DEGREE           # one relation, < 1.5 entries per row: four rows per warp
        if r.heavy_rows is not None and r.n_heavy is not None:
            arr[i].heavy_rows, arr[i].n_heavy = r.heavy_rows.data_ptr(), r.n_heavy.data_ptr()
            arr[i].heavy_cap = int(r.heavy_rows.numel()) // 2        # [cap] rows + [cap] chunk prefix
    return arr


def rel_of(csr: CSR, k: int, src: torch.Tensor, **kw) -> "Rel":
    """Relation ``k`` of a CSR as a gather operand, heavy-row list included."""
    return Rel(csr.rowptr[k], csr.col, src, heavy_rows=csr.heavy[k] if csr.heavy is not None else None,
               n_heavy=csr.n_heavy[k:k + 1] if csr.n_heavy is not None else None, **kw)


class KernelTimer:
    """Optional per-launch CUDA-event timing of the aggregation kernels (bench.py's roofline):
    ``ops.timer = KernelTimer()`` records (name, algorithmic bytes, start, stop) on the launching
    stream; ``summary()`` synchronises and returns totals per kernel name."""

    def __init__(self):
        self.records = []
        self.tags = []

    def launch(self, name: str, nbytes: int, device, fn, tag=None):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.current_stream(device)
        start.record(st)
        fn()
        stop.record(st)
        self.records.append((name, nbytes, start, stop))
        self.tags.append(tag)

    def by_tag(self, name: str):
        """{tag: (launches, total ms)} of the launches recorded under ``name``."""
        torch.cuda.synchronize()
        out = {}
        for (n, _, a, b), tag in zip(self.records, self.tags):
            if n == name:
                c, ms = out.get(tag, (0, 0.0))
                out[tag] = (c + 1, ms + a.elapsed_time(b))
        return out

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, nbytes, a, b in self.records:
            d = out.setdefault(name, {"launches": 0, "bytes": 0, "ms": 0.0, "max_bytes": 0, "max_ms": 0.0})
            ms = a.elapsed_time(b)
            d["launches"] += 1
            d["bytes"] += nbytes
            d["ms"] += ms
            if nbytes > d["max_bytes"]:
                d["max_bytes"], d["max_ms"] = nbytes, ms
        return out


timer: Optional[KernelTimer] = None


def gather_bytes(rels: Sequence[Rel], n_rows: int, n_feat: int, elem: int, concat: bool, has_self: bool,
                 has_copy: bool) -> int:
    """Algorithmic bytes of one agnn_gather_reduce launch: gathered rows and their
    column ids, rows written, self / copy rows read, rowptr (and neighbour-degree reads where used)."""
    row = n_feat * elem
    total = 0
    for r in rels:
        e = int(r.n_edges)
        total += e * (row + 4) + 4 * (n_rows + 1)
        if r.nbr_deg_rowptr is not None:
            total += 8 * e
    slices = len(rels) if concat else 1
    total += n_rows * row * (slices + int(has_self) + 2 * int(has_copy))
    return total


def gather_reduce(rels: Sequence[Rel], out: torch.Tensor, n_feat: int, mean: bool, concat: bool,
                  self_add: Optional[torch.Tensor] = None, copy: Optional[torch.Tensor] = None,
                  copy_col: int = 0, out_lo: Optional[torch.Tensor] = None,
                  pair_amax: Optional[torch.Tensor] = None, amax_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """agnn_gather_reduce on ``out`` ([n_rows, >= n_feat] view).  See include/agnn.h.  With ``out_lo``
    the result is written as the TF32 hi / lo pair (``out``, ``out_lo``) that ``agnn_gemm`` consumes; with
    ``pair_amax`` too, ``out`` / ``out_lo`` are fp16 and receive the F16X3 pair (agnn_gather_reduce_f16)."""
    out = _rows2d(out, "out")
    if pair_amax is not None and (out.dtype != torch.float16 or out_lo is None or not concat):
        raise ValueError("the fp16 operand pair needs fp16 out / out_lo buffers and the concatenated layout")
    if out_lo is not None and (out_lo.shape != out.shape or out_lo.stride() != out.stride()):
        raise ValueError("out_lo must match out in shape and strides")
    n_rows = out.shape[0]
    in_dtype = torch.float32 if pair_amax is not None else out.dtype
    arr = _pack(rels, n_feat, in_dtype)
    sa = _rows2d(self_add, "self_add") if self_add is not None else None
    cp = _rows2d(copy, "copy") if copy is not None else None
    stream = torch.cuda.current_stream(out.device).cuda_stream
    ws, ws_bytes = None, 0
    if any(r.heavy_rows is not None for r in rels) and all(r.n_edges is not None for r in rels):
        # rows of >= HEAVY_ROW entries (hubs) are split across warps; a relation below that size cannot have one
        if max(int(r.n_edges) for r in rels) >= _lib.HEAVY_ROW:
            ws_bytes = _lib.lib().agnn_gather_heavy_workspace(
                sum(int(r.n_edges) for r in rels),
                sum(int(r.heavy_rows.numel()) // 2 for r in rels if r.heavy_rows is not None), n_feat)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=out.device)

    def run():
        if pair_amax is not None:
            _lib.check(_lib.lib().agnn_gather_reduce_f16(
                n_rows, n_feat, _lib.F32, _lib.SCALE_MEAN if mean else _lib.SCALE_NONE, _lib.COMBINE_CONCAT, len(rels),
                arr, sa.data_ptr() if sa is not None else None, sa.stride(0) if sa is not None else 0,
                cp.data_ptr() if cp is not None else None, cp.stride(0) if cp is not None else 0, int(copy_col),
                out.data_ptr(), out.stride(0), out_lo.data_ptr(), pair_amax.data_ptr(),
                ws.data_ptr() if ws is not None else None, ws_bytes, stream), "agnn_gather_reduce_f16")
            return
        _lib.check(_lib.lib().agnn_gather_reduce_amax(
            n_rows, n_feat, _dtype_code(out), _lib.SCALE_MEAN if mean else _lib.SCALE_NONE,
            _lib.COMBINE_CONCAT if concat else _lib.COMBINE_SUM, len(rels), arr,
            sa.data_ptr() if sa is not None else None, sa.stride(0) if sa is not None else 0,
            cp.data_ptr() if cp is not None else None, cp.stride(0) if cp is not None else 0, int(copy_col),
            out.data_ptr(), out.stride(0), out_lo.data_ptr() if out_lo is not None else None, None,
            amax_out.data_ptr() if amax_out is not None else None,
            ws.data_ptr() if ws is not None else None, ws_bytes, stream), "agnn_gather_reduce")

    if timer is not None and all(r.n_edges is not None for r in rels):
        nbytes = gather_bytes(rels, n_rows, n_feat, 4 if pair_amax is not None else out.element_size(), concat,
                              sa is not None, cp is not None)
        written = n_rows * n_feat * ((len(rels) if concat else 1) + int(cp is not None))
        if pair_amax is None and out_lo is not None:   # every written row goes out twice (hi and lo); the fp16 pair
            nbytes += written * out.element_size()     # is 2 + 2 bytes per element = one fp32 row, already counted
        timer.launch("gather_reduce", nbytes, out.device, run)
    else:
        run()
    _lib.count_launches(3 if ws is not None else 1)
    return out


def rowscale_sum(rels: Sequence[Rel], inp: torch.Tensor, out: torch.Tensor, n_feat: int,
                 base: Optional[torch.Tensor] = None) -> torch.Tensor:
    inp, out = _rows2d(inp, "in"), _rows2d(out, "out")
    arr = (_lib.Rel * len(rels))()
    for i, r in enumerate(rels):
        arr[i].rowptr = r.rowptr.data_ptr()
        arr[i].out_col = int(r.out_col)
        arr[i].flags = int(r.flags)
    b = _rows2d(base, "base") if base is not None else None
    stream = torch.cuda.current_stream(out.device).cuda_stream
    _lib.check(_lib.lib().agnn_rowscale_sum(
        out.shape[0], n_feat, _dtype_code(out), len(rels), arr, inp.data_ptr(), inp.stride(0),
        b.data_ptr() if b is not None else None, b.stride(0) if b is not None else 0,
        out.data_ptr(), out.stride(0), stream), "agnn_rowscale_sum")
    _lib.count_launches(1)
    return out


# ------------------------------------------------------------------------------
# single-relation segmented reductions with autograd (MetricalConvLayer scatters,
# onset pooling, MetricalGNN beat/measure initialisation)
# ------------------------------------------------------------------------------

class _SegmentReduce(torch.autograd.Function):
    """out[i] = s_i * (self_i + sum_{k in row i} src[col[k]])  on ``csr.fwd`` relation 0;
    backward runs the same kernel on ``csr.bwd``."""

    @staticmethod
    def forward(ctx, src, self_add, csr: TypedCSR, mean: bool):
        src = src.contiguous()
        f = src.shape[1]
        out = torch.empty((csr.n_rows, f), dtype=src.dtype, device=src.device)
        sa = self_add.contiguous() if self_add is not None else None
        gather_reduce([rel_of(csr.fwd, 0, src, n_edges=csr.n_edges)], out, f, mean=mean, concat=True,
                      self_add=sa)
        ctx.csr, ctx.mean, ctx.has_self = csr, mean, self_add is not None
        return out

    @staticmethod
    def backward(ctx, g):
        csr, f = ctx.csr, g.shape[1]
        g = g.contiguous()
        d_src = torch.empty((csr.n_cols, f), dtype=g.dtype, device=g.device)
        deg = csr.fwd.rowptr[0] if ctx.mean else None
        gather_reduce([rel_of(csr.bwd, 0, g, nbr_deg_rowptr=deg, n_edges=csr.n_edges)], d_src, f, mean=False,
                      concat=True)
        d_self = None
        if ctx.has_self:
            if ctx.mean:
                d_self = torch.empty_like(g)
                rowscale_sum([Rel(csr.fwd.rowptr[0], csr.fwd.col, g)], g, d_self, f)
            else:
                d_self = g
        return d_src, d_self, None, None


def segment_sum(src, csr: TypedCSR):
    """``scatter_add(src[e_gather], e_reduce, out=zeros)`` (gnn.py:511,539; hgnn.py:406-407)."""
    return _SegmentReduce.apply(src, None, csr, False)


def segment_sum_self(src, self_add, csr: TypedCSR):
    """``scatter(src[e_gather], e_reduce, out=self.clone(), reduce='sum')`` (gnn.py:256)."""
    return _SegmentReduce.apply(src, self_add, csr, False)


def segment_mean_self(src, self_add, csr: TypedCSR):
    """``scatter(src[e_gather], e_reduce, out=self.clone(), reduce='mean')`` (gnn.py:74; analysis.py:586)."""
    return _SegmentReduce.apply(src, self_add, csr, True)


# ------------------------------------------------------------------------------
# fused per-edge messages (agnn_edge_op): ResGatedGraphConv / RelEdgeConv / OnsetEmbedding
# ------------------------------------------------------------------------------

def _edge_op(op, csr_side: CSR, n_rows, f, row0, row1, nbr0, nbr1, self_add=None, mean=False, two_out=False):
    ok = lambda t: t is None or (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1)
    if not all(ok(t) for t in (row0, row1, nbr0, nbr1, self_add)):
        raise _lib.AgnnError("edge_op: operands must be fp32 CUDA matrices with unit column stride")
    out0 = torch.empty((n_rows, f), dtype=torch.float32, device=row0.device)
    out1 = torch.empty((n_rows, f), dtype=torch.float32, device=row0.device) if two_out else None
    ptr = lambda t: t.data_ptr() if t is not None else None
    ld = lambda t: t.stride(0) if t is not None else 0
    _lib.check(_lib.lib().agnn_edge_op(op, n_rows, f, csr_side.rowptr[0].data_ptr(), csr_side.col.data_ptr(), ptr(row0),
                                       ld(row0), ptr(row1), ld(row1), ptr(nbr0), ld(nbr0), ptr(nbr1), ld(nbr1),
                                       ptr(self_add), ld(self_add), int(mean), out0.data_ptr(), out0.stride(0),
                                       ptr(out1), ld(out1), _stream(row0)), "agnn_edge_op")
    _lib.count_launches(1)
    return (out0, out1) if two_out else out0


def _scaled_by_degree(g, csr: TypedCSR):
    out = torch.empty_like(g)
    rowscale_sum([Rel(csr.fwd.rowptr[0], csr.fwd.col, g)], g, out, g.shape[1])
    return out


class _EdgeAbsDiff(torch.autograd.Function):
    """``out_i = s_i (self_i + sum_{j in N(i)} |a_i - b_j|)`` on ``csr.fwd`` (rows = reduce side); ``s_i`` = 1 or
    ``1 / max(deg_i, 1)``.  Arguments: (a, b, self_add or None, csr, mean)."""

    @staticmethod
    def forward(ctx, a, b, self_add, csr: TypedCSR, mean: bool):
        a, b = a.contiguous(), b.contiguous()
        sa = self_add.contiguous() if self_add is not None else None
        out = _edge_op(_lib.EDGE_ABSDIFF, csr.fwd, csr.n_rows, a.shape[1], a, None, b, None, sa, mean)
        ctx.save_for_backward(a, b)
        ctx.csr, ctx.mean, ctx.has_self = csr, mean, self_add is not None
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        csr, f = ctx.csr, a.shape[1]
        g = g.contiguous()
        gs = _scaled_by_degree(g, csr) if ctx.mean else g
        da = _edge_op(_lib.EDGE_ABSDIFF_DROW, csr.fwd, csr.n_rows, f, gs, a, b, None) if ctx.needs_input_grad[0] else None
        db = _edge_op(_lib.EDGE_ABSDIFF_DNBR, csr.bwd, csr.n_cols, f, b, None, gs, a) if ctx.needs_input_grad[1] else None
        return da, db, (gs if ctx.has_self else None), None, None


class _EdgeGate(torch.autograd.Function):
    """``out_i = self_i + sum_{j in N(i)} sigmoid(a_i + b_j) * c_j`` (gnn.py:249-256).  Arguments: (a, b, c, self_add, csr)."""

    @staticmethod
    def forward(ctx, a, b, c, self_add, csr: TypedCSR):
        a, b, c = a.contiguous(), b.contiguous(), c.contiguous()
        sa = self_add.contiguous() if self_add is not None else None
        out = _edge_op(_lib.EDGE_GATE, csr.fwd, csr.n_rows, a.shape[1], a, None, b, c, sa, False)
        ctx.save_for_backward(a, b, c)
        ctx.csr, ctx.has_self = csr, self_add is not None
        return out

    @staticmethod
    def backward(ctx, g):
        a, b, c = ctx.saved_tensors
        csr, f = ctx.csr, a.shape[1]
        g = g.contiguous()
        da = _edge_op(_lib.EDGE_GATE_DROW, csr.fwd, csr.n_rows, f, g, a, b, c) if ctx.needs_input_grad[0] else None
        db = dc = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            db, dc = _edge_op(_lib.EDGE_GATE_DNBR, csr.bwd, csr.n_cols, f, b, c, g, a, two_out=True)
        return da, db, dc, (g if ctx.has_self else None), None


def edge_absdiff(a, b, csr: TypedCSR, self_add=None, mean: bool = False):
    """``s_i (self_i + sum_j |a_i - b_j|)``: no per-edge tensor on either pass."""
    return _EdgeAbsDiff.apply(a, b, self_add, csr, mean)


def edge_gate_sum(a, b, c, csr: TypedCSR, self_add=None):
    """``self_i + sum_j sigmoid(a_i + b_j) * c_j``."""
    return _EdgeGate.apply(a, b, c, self_add, csr)


# ------------------------------------------------------------------------------
# in-tree HeteroConv{SageConvScatter}: one fused layer
# ------------------------------------------------------------------------------

def _operand_buffers(rows: int, cols: int, like: torch.Tensor):
    """(hi, lo) buffers for a GEMM operand the gather kernel writes: a TF32 pair on the tcgen05 fp32 route
    (lo is None for other dtypes / widths, and hi then holds the plain values)."""
    if like.dtype == torch.float32 and cols % 4 == 0:
        buf = torch.empty((2, rows, cols), dtype=like.dtype, device=like.device)
        return buf[0], buf[1]
    return torch.empty((rows, cols), dtype=like.dtype, device=like.device), None


def _as_operand_pair(hi, lo):
    return linalg.Split(hi, lo) if lo is not None else hi


class _IntreeSageLayer(torch.autograd.Function):
    """All relations of a reference ``HeteroConv(module=SageConvScatter)`` layer
    (analysisgnn/models/core/hgnn.py:479-484 over gnn.py:62-76) in 2 GEMMs + 1 gather:

        H = x Wn_cat^T + bn_cat                              [N, R*F]
        A = [x || S_1 .. S_R],  S_r = (x + sum_j H_r[j]) / max(deg_r, 1)  (H_r itself if E_r == 0)
        Z = A Wc^T + bc                                      [N, F']

    ``Wc``/``bc`` carry the relation reduction (mean or sum) folded in by the caller.
    """

    @staticmethod
    def forward(ctx, x, wn_cat, bn_cat, wc, bc, csr: TypedCSR):
        x = x.contiguous()
        n, f = x.shape
        r = csr.n_rel
        # fp16 operand form: |S_r| <= amax(x) + amax(H), inside the 4x headroom of a scale taken from
        # max(amax(x), amax(H))
        f16 = (linalg.parity_operands() == "f16" and x.dtype == torch.float32
               and f % 8 == 0 and n > 0 and linalg.f16_ok(wn_cat) and linalg.f16_ok(wc))
        xs = linalg.split_f16(x) if f16 else x
        h = linalg.linear(xs, wn_cat, bn_cat)                                    # [N, R*F]
        rels = [rel_of(csr.fwd, k, h[:, k * f:(k + 1) * f], out_col=(k + 1) * f,
                       flags=_lib.REL_IDENTITY_IF_EMPTY, n_edges=csr.n_edges if k == 0 else 0) for k in range(r)]
        a_amax = None
        if f16:
            a_amax = linalg.amax_into(linalg.amax_into(linalg.new_amax(x.device), x), h)
            buf = torch.empty((2, n, (r + 1) * f), dtype=torch.float16, device=x.device)
            gather_reduce(rels, buf[0], f, mean=True, concat=True, self_add=x, copy=x, copy_col=0, out_lo=buf[1],
                          pair_amax=a_amax)
            a = linalg.SplitH(buf[0], buf[1], a_amax)
        else:
            a_hi, a_lo = _operand_buffers(n, (r + 1) * f, x)
            # the gather writes the GEMM operand directly as a TF32 pair: it feeds the forward and the grad-weight GEMM
            gather_reduce(rels, a_hi, f, mean=True, concat=True, self_add=x, copy=x, copy_col=0, out_lo=a_lo)
            a = _as_operand_pair(a_hi, a_lo)
        z = linalg.linear(a, wc, bc)
        ctx.save_for_backward(x, *linalg.pack(a), wn_cat, wc)
        ctx.csr = csr
        ctx.a_amax = a_amax
        ctx.has_bias = (bn_cat is not None, bc is not None)
        return z

    @staticmethod
    def backward(ctx, dz):
        x, a_first, a_second, wn_cat, wc = ctx.saved_tensors
        a = linalg.unpack(a_first, a_second, ctx.a_amax)
        f16 = isinstance(a, linalg.SplitH)
        prep = (lambda t: linalg.split_f16(t) if linalg.f16_ok(t) else linalg.prepare(t)) if f16 else linalg.prepare
        csr, (n, f), r = ctx.csr, x.shape, ctx.csr.n_rel
        dz_plain = dz.contiguous()
        dz = prep(dz_plain)
        da = linalg.mm(dz, wc)                                                   # [N, (R+1)F]
        dwc = linalg.mm_tn(dz, a) if ctx.needs_input_grad[3] else None
        dbc = colsum(dz_plain) if ctx.has_bias[1] and ctx.needs_input_grad[4] else None
        dh = torch.empty((n, r * f), dtype=x.dtype, device=x.device)
        rels_t = [rel_of(csr.bwd, k, da[:, (k + 1) * f:(k + 2) * f], out_col=k * f,
                         nbr_deg_rowptr=csr.fwd.rowptr[k], flags=_lib.REL_IDENTITY_IF_EMPTY,
                         n_edges=csr.n_edges if k == 0 else 0) for k in range(r)]
        gather_reduce(rels_t, dh, f, mean=False, concat=True)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            rels_s = [Rel(csr.fwd.rowptr[k], csr.fwd.col, da, out_col=(k + 1) * f,
                          flags=_lib.REL_IDENTITY_IF_EMPTY) for k in range(r)]
            rowscale_sum(rels_s, da, dx, f, base=da[:, :f])
            dh_op = prep(dh)
            linalg.mm(dh_op, wn_cat, out=dx, accumulate=True)
        else:
            dh_op = dh
        dwn = linalg.mm_tn(dh_op, x) if ctx.needs_input_grad[1] else None
        dbn = colsum(dh) if ctx.has_bias[0] and ctx.needs_input_grad[2] else None
        return dx, dwn, dbn, dwc, dbc, None


def intree_sage_layer(x, wn_cat, bn_cat, wc, bc, csr: TypedCSR):
    return _IntreeSageLayer.apply(x, wn_cat, bn_cat, wc, bc, csr)


# ------------------------------------------------------------------------------
# PyG-style HeteroConv{SAGEConv}: one fused layer over all node types
# ------------------------------------------------------------------------------

class _HeteroSageLayer(torch.autograd.Function):
    """PyG ``HeteroConv({et: SAGEConv}, aggr)`` (+ optional ReLU) for every destination
    type at once.  Per destination type t with incoming relations r_1..r_k:

        A_t = [x_t || mean_{r_1}(x_src) || .. || mean_{r_k}(x_src)]    (one gather launch)
        out_t = A_t [sum_r Wr_r | Wl_1 | .. | Wl_k]^T + sum_r b_r      (one GEMM)

    The GEMMs of all destination types (and, backwards, their grad-input and grad-weight products) are ONE grouped
    launch each.  The forward GEMM epilogue reports max |out| over all types into ``box["out_amax"]`` -- the next
    layer's operand scale -- and the backward gathers report max |d x|.

    argument layout: (plan, csr, relu, box, *tensors) with tensors = x per node type
    (plan.node_types order), then (Wcat_t, b_t) per destination type; ``box`` = {"x_amax": tagged scalar or None}.
    """

    @staticmethod
    def forward(ctx, plan, csr: HeteroCSR, relu: bool, box: dict, *tensors):
        nt = len(plan.node_types)
        xs = {t: tensors[i].contiguous() for i, t in enumerate(plan.node_types)}
        # F16X3 operands: every aggregated row is a copy or a mean of input rows, so the amax over the layer's inputs
        # bounds every operand the gather writes -- one scalar for all destination types
        f16 = (linalg.parity_operands() == "f16"
               and all(v.dtype == torch.float32 and v.shape[1] % 8 == 0 for v in xs.values()))
        x_amax = None
        if f16:
            x_amax = box.get("x_amax")                     # the producer's epilogue already measured it
            if x_amax is None:
                x_amax = linalg.new_amax(tensors[0].device)
                for v in xs.values():
                    linalg.stats["amax_passes"] = linalg.stats.get("amax_passes", 0) + 1
                    linalg.amax_into(x_amax, v)
        out_amax = linalg.new_amax(tensors[0].device) if f16 else None
        box["out_amax"] = out_amax
        operands, specs, saved = [], [], []
        for j, t in enumerate(plan.dst_types):
            wcat, bias = tensors[nt + 2 * j], tensors[nt + 2 * j + 1]
            x_t = xs[t]
            f = x_t.shape[1]
            rel_list = plan.incoming[t]
            rels = [rel_of(csr.fwd[et], 0, xs[et[0]], out_col=(k + 1) * f, n_edges=csr.n_edges[et])
                    for k, et in enumerate(rel_list)]
            if f16 and x_t.shape[0] > 0:
                buf = torch.empty((2, x_t.shape[0], (len(rel_list) + 1) * f), dtype=torch.float16, device=x_t.device)
                gather_reduce(rels, buf[0], f, mean=True, concat=True, copy=x_t, copy_col=0, out_lo=buf[1],
                              pair_amax=x_amax)
                a = linalg.SplitH(buf[0], buf[1], x_amax)
            else:
                a_hi, a_lo = _operand_buffers(x_t.shape[0], (len(rel_list) + 1) * f, x_t)
                gather_reduce(rels, a_hi, f, mean=True, concat=True, copy=x_t, copy_col=0, out_lo=a_lo)
                a = _as_operand_pair(a_hi, a_lo)
            operands.append(a)
            specs.append(dict(a=a, b=wcat, m=x_t.shape[0], n=wcat.shape[0], k=(len(rel_list) + 1) * f, bias=bias,
                              flags=_lib.GEMM_RELU if relu else 0,
                              amax_out=out_amax if isinstance(a, linalg.SplitH) else None))
        outs = linalg._group(_lib.K_MAJOR, _lib.K_MAJOR, specs)
        for a, sp, o in zip(operands, specs, outs):
            saved += [*linalg.pack(a), sp["b"], o]
        ctx.save_for_backward(*saved)
        ctx.x_amax = x_amax
        ctx.set_materialize_grads(False)        # an unused destination type costs nothing in backward
        ctx.plan, ctx.csr, ctx.relu = plan, csr, relu
        ctx.feat = {t: xs[t].shape[1] for t in plan.node_types}
        ctx.rows = {t: xs[t].shape[0] for t in plan.node_types}
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        plan, csr = ctx.plan, ctx.csr
        nt = len(plan.node_types)
        base = 4                                 # plan, csr, relu, box precede the tensors
        grads = [None] * (nt + 2 * len(plan.dst_types))
        act, gs_list, a_list, w_list = [], [], [], []
        for j, t in enumerate(plan.dst_types):
            a_first, a_second, wcat, o = ctx.saved_tensors[4 * j:4 * j + 4]
            g = douts[j]
            if g is None or g.shape[0] == 0:
                continue
            a = linalg.unpack(a_first, a_second, ctx.x_amax)
            g, db = prepare_grad(g, o if ctx.relu else None, ctx.needs_input_grad[base + nt + 2 * j + 1],
                                 f16=isinstance(a, linalg.SplitH))
            grads[nt + 2 * j + 1] = db
            act.append((j, t))
            gs_list.append(g)
            a_list.append(a)
            w_list.append(wcat)
        # grad-input and grad-weight products of all destination types: one grouped launch each
        das = linalg.mm_group(gs_list, w_list) if act else []
        da = {t: d for (j, t), d in zip(act, das)}
        sel = [i for i, (j, t) in enumerate(act) if ctx.needs_input_grad[base + nt + 2 * j]]
        dws = linalg.mm_tn_group([gs_list[i] for i in sel], [a_list[i] for i in sel]) if sel else []
        for i, dw in zip(sel, dws):
            grads[nt + 2 * act[i][0]] = dw
        for i, s in enumerate(plan.node_types):
            if not ctx.needs_input_grad[base + i]:
                continue
            f = ctx.feat[s]
            rels = []
            for et in plan.outgoing[s]:
                dst = et[2]
                if dst not in da:
                    continue
                k = plan.incoming[dst].index(et)
                rels.append(rel_of(csr.bwd[et], 0, da[dst][:, (k + 1) * f:(k + 2) * f],
                                   nbr_deg_rowptr=csr.fwd[et].rowptr[0], n_edges=csr.n_edges[et]))
            root = da[s][:, :f] if s in da else None
            if not rels:
                grads[i] = root.contiguous() if root is not None else None
                continue
            dx = torch.empty((ctx.rows[s], f), dtype=rels[0].src.dtype, device=rels[0].src.device)
            am = linalg.new_amax(dx.device) if dx.dtype == torch.float32 else None
            gather_reduce(rels, dx, f, mean=False, concat=False, self_add=root, amax_out=am)
            grads[i] = linalg.tag_amax(dx, am)
        return (None, None, None, None, *grads)


class _SageWeights(torch.autograd.Function):
    """The fused layer's weight for one destination type: ``[sum_r Wr_r | Wl_1 | .. | Wl_k] * scale`` and
    ``sum_r b_r * scale`` from the k relations' ``lin_r.weight``, ``lin_l.weight``, ``lin_l.bias`` (PyG SAGEConv
    parameters; HeteroConv's relation sum / mean folded in).

    Backward: the gradient of every parameter is a column block (or the root block / the bias) of the fused weight's
    gradient.  Left to autograd those are 3k strided or shared views, each of which AccumulateGrad clones with its own
    copy kernel (~100 per step at config 2).  Here ONE gather writes all 2k weight gradients as the contiguous rows of a
    ``[2k, N, F]`` buffer (and one broadcast the k bias gradients), which AccumulateGrad takes over as they are.

    Arguments: (scale, k, *lin_r weights, *lin_l weights, *lin_l biases)."""

    @staticmethod
    def forward(ctx, scale, k, *tensors):
        wr, wl, bl = tensors[:k], tensors[k:2 * k], tensors[2 * k:]
        n, f = wl[0].shape
        fused = (wl[0].is_cuda and wl[0].dtype == torch.float32 and f % 4 == 0 and k <= _lib.MAX_REL
                 and all(t.is_contiguous() and t.data_ptr() % 16 == 0 for t in tensors[:2 * k])
                 and all(t.is_contiguous() for t in bl))
        if fused:                                                  # one launch (agnn_sage_weights)
            wcat = torch.empty((n, (k + 1) * f), dtype=torch.float32, device=wl[0].device)
            bias = torch.empty(n, dtype=torch.float32, device=wl[0].device)
            _lib.check(_lib.lib().agnn_sage_weights(k, n, f, _lib.ptr_array(list(wr)), _lib.ptr_array(list(wl)),
                                                    _lib.ptr_array(list(bl)), float(scale), wcat.data_ptr(),
                                                    bias.data_ptr(), _stream(wcat)), "agnn_sage_weights")
            _lib.count_launches(1)
        else:
            if k > 1:
                root = torch.stack(wr, dim=0).sum(0)
                bias = torch.stack(bl, dim=0).sum(0)
            else:
                root, bias = wr[0], bl[0].clone()
            wcat = torch.cat([root] + list(wl), dim=1)
            if scale != 1.0:
                wcat.mul_(scale)
                bias.mul_(scale)
        ctx.scale, ctx.k = scale, k
        ctx.feat = wl[0].shape[1]
        return wcat, bias

    @staticmethod
    def backward(ctx, dwcat, dbias):
        k, scale, f = ctx.k, ctx.scale, ctx.feat
        outs = [None] * (3 * k)
        if dwcat is not None:
            n = dwcat.shape[0]
            blocks = dwcat.reshape(n, k + 1, f).permute(1, 0, 2)               # [k + 1, N, F] view
            idx = _sage_block_index(k, dwcat.device)                           # [0] * k + [1 .. k]
            g = blocks.index_select(0, idx)                                    # [2k, N, F], contiguous: one kernel
            if scale != 1.0:
                g.mul_(scale)
            for r in range(2 * k):
                outs[r] = g[r]
        if dbias is not None:
            gb = dbias.unsqueeze(0).expand(k, -1)
            gb = gb * scale if scale != 1.0 else gb.clone()      # (never in place: the incoming gradient is not ours)
            for r in range(k):
                outs[2 * k + r] = gb[r]
        return (None, None, *outs)


_sage_idx_cache = {}


def _sage_block_index(k: int, device) -> torch.Tensor:
    key = (k, str(device))
    if key not in _sage_idx_cache:
        _sage_idx_cache[key] = torch.tensor([0] * k + list(range(1, k + 1)), dtype=torch.long, device=device)
    return _sage_idx_cache[key]


def sage_weights(lin_r_weights, lin_l_weights, lin_l_biases, scale: float = 1.0):
    k = len(lin_l_weights)
    return _SageWeights.apply(float(scale), k, *lin_r_weights, *lin_l_weights, *lin_l_biases)


def hetero_sage_layer(plan, csr: HeteroCSR, relu: bool, xs: Sequence[torch.Tensor], params: Sequence[torch.Tensor]):
    """Outputs carry the amax tag of the layer (one scalar for all destination types); inputs that all carry ONE
    common tag (the previous layer's outputs, the grouped per-node-type projections) need no amax pass."""
    tags = [linalg.known_amax(x) for x in xs]
    box = {"x_amax": tags[0] if tags and all(t is not None and t is tags[0] for t in tags) else None}
    outs = _HeteroSageLayer.apply(plan, csr, relu, box, *xs, *params)
    for o in outs:
        linalg.tag_amax(o, box.get("out_amax"))
    return outs


# ------------------------------------------------------------------------------
# HGT attention (PyG HGTConv edge pipeline) -- agnn_hgt_attn_{fwd,bwd_dst,bwd_src}
# ------------------------------------------------------------------------------

def _hgt_pack(fwd_csrs: Sequence[CSR], bwd_csrs: Sequence[CSR], ks, vs, dks=None, dvs=None):
    arr = (_lib.HgtRel * len(ks))()
    for i, (cf, cb, k, v) in enumerate(zip(fwd_csrs, bwd_csrs, ks, vs)):
        k, v = _rows2d(k, "k"), _rows2d(v, "v")
        if k.stride(0) != v.stride(0):
            raise ValueError("k and v of a relation must share their row stride")
        arr[i].rowptr, arr[i].col = cf.rowptr.data_ptr(), cf.col.data_ptr()
        arr[i].t_rowptr, arr[i].t_col = cb.rowptr.data_ptr(), cb.col.data_ptr()
        arr[i].k, arr[i].v, arr[i].ld_kv = k.data_ptr(), v.data_ptr(), k.stride(0)
        arr[i].n_src = k.shape[0]
        if dks is not None:
            arr[i].dk, arr[i].dv, arr[i].ld_dkv = dks[i].data_ptr(), dvs[i].data_ptr(), dks[i].stride(0)
    return arr


class _HGTAttention(torch.autograd.Function):
    """Joint edge softmax over the given relations of one destination type.
    Arguments: (q [N_dst, H*D], pscale [R, H] fp32, heads, fwd_csrs, bwd_csrs, *ks, *vs)."""

    @staticmethod
    def forward(ctx, q, pscale, heads, fwd_csrs, bwd_csrs, *kv):
        r = len(fwd_csrs)
        ks = [t.contiguous() for t in kv[:r]]
        vs = [t.contiguous() for t in kv[r:]]
        q = _rows2d(q, "q")
        n, hd = q.shape
        d = hd // heads
        ps = pscale.detach().to(torch.float32).contiguous()
        out = torch.empty((n, hd), dtype=q.dtype, device=q.device)
        row_max = torch.empty((n, heads), dtype=torch.float32, device=q.device)
        row_den = torch.empty((n, heads), dtype=torch.float32, device=q.device)
        arr = _hgt_pack(fwd_csrs, bwd_csrs, ks, vs)
        stream = torch.cuda.current_stream(q.device).cuda_stream
        _lib.check(_lib.lib().agnn_hgt_attn_fwd(n, heads, d, _dtype_code(q), r, arr, q.data_ptr(), q.stride(0),
                                                ps.data_ptr(), out.data_ptr(), out.stride(0), row_max.data_ptr(),
                                                row_den.data_ptr(), stream), "agnn_hgt_attn_fwd")
        _lib.count_launches(1)
        ctx.save_for_backward(q, ps, out, row_max, row_den, *ks, *vs)
        ctx.heads, ctx.fwd_csrs, ctx.bwd_csrs = heads, fwd_csrs, bwd_csrs
        ctx.pscale_dtype = pscale.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        q, ps, out, row_max, row_den = ctx.saved_tensors[:5]
        r = len(ctx.fwd_csrs)
        ks, vs = ctx.saved_tensors[5:5 + r], ctx.saved_tensors[5 + r:]
        heads = ctx.heads
        n, hd = q.shape
        d = hd // heads
        dout = dout.contiguous()
        lib = _lib.lib()
        stream = torch.cuda.current_stream(q.device).cuda_stream
        dq = torch.empty((n, hd), dtype=q.dtype, device=q.device)
        delta = torch.empty((n, heads), dtype=torch.float32, device=q.device)
        blocks = lib.agnn_hgt_attn_bwd_dst_blocks(n)
        partial = torch.empty((blocks, r * heads), dtype=torch.float32, device=q.device)
        dks = [torch.empty_like(k) for k in ks]
        dvs = [torch.empty_like(v) for v in vs]
        arr = _hgt_pack(ctx.fwd_csrs, ctx.bwd_csrs, ks, vs, dks, dvs)
        if n > 0:
            _lib.check(lib.agnn_hgt_attn_bwd_dst(n, heads, d, _dtype_code(q), r, arr, q.data_ptr(), q.stride(0),
                                                 ps.data_ptr(), out.data_ptr(), out.stride(0), dout.data_ptr(),
                                                 dout.stride(0), row_max.data_ptr(), row_den.data_ptr(),
                                                 delta.data_ptr(), dq.data_ptr(), dq.stride(0), partial.data_ptr(),
                                                 stream), "agnn_hgt_attn_bwd_dst")
            dps = partial.sum(0).view(r, heads).to(ctx.pscale_dtype)
        else:
            dps = torch.zeros((r, heads), dtype=ctx.pscale_dtype, device=q.device)
        _lib.check(lib.agnn_hgt_attn_bwd_src(heads, d, _dtype_code(q), r, arr, q.data_ptr(), q.stride(0),
                                             ps.data_ptr(), dout.data_ptr(), dout.stride(0), row_max.data_ptr(),
                                             row_den.data_ptr(), delta.data_ptr(), stream), "agnn_hgt_attn_bwd_src")
        _lib.count_launches(2)
        return (dq, dps, None, None, None, *dks, *dvs)


def hgt_attention(q, ks, vs, pscale, fwd_csrs, bwd_csrs, heads: int, joint_softmax: bool = True):
    """``sum_e softmax_e(q_i . k_e * pscale) v_e`` per destination row and head.

    ``joint_softmax=True``: one softmax over the incoming edges of all relations (PyG >= 2.3);
    ``False``: one softmax per relation, results summed (pre-2.3)."""
    if joint_softmax:
        return _HGTAttention.apply(q, pscale, heads, list(fwd_csrs), list(bwd_csrs), *ks, *vs)
    out = None
    for i in range(len(ks)):
        o = _HGTAttention.apply(q, pscale[i:i + 1], heads, [fwd_csrs[i]], [bwd_csrs[i]], ks[i], vs[i])
        out = o if out is None else out + o
    return out




class HgtGroup:
    """Relations of one destination type that share a softmax: (src slot, k column, v column, fwd CSR, bwd CSR)."""

    __slots__ = ("rels",)

    def __init__(self, rels):
        self.rels = list(rels)


class HgtTarget:
    """One destination type of a layer: slot of its wide projection, column of q, softmax groups."""

    __slots__ = ("slot", "q_off", "groups")

    def __init__(self, slot, q_off, groups):
        self.slot, self.q_off, self.groups = slot, q_off, list(groups)


def _hgt_pack_wide(group: HgtGroup, ys, hd, dys=None):
    arr = (_lib.HgtRel * len(group.rels))()
    esz = ys[0].element_size()
    for i, (slot, k_off, v_off, cf, cb) in enumerate(group.rels):
        y = ys[slot]
        arr[i].rowptr, arr[i].col = cf.rowptr.data_ptr(), cf.col.data_ptr()
        arr[i].t_rowptr, arr[i].t_col = cb.rowptr.data_ptr(), cb.col.data_ptr()
        arr[i].k, arr[i].v, arr[i].ld_kv = y.data_ptr() + k_off * esz, y.data_ptr() + v_off * esz, y.stride(0)
        arr[i].n_src = y.shape[0]
        if dys is not None:
            dy = dys[slot]
            arr[i].dk, arr[i].dv, arr[i].ld_dkv = dy.data_ptr() + k_off * esz, dy.data_ptr() + v_off * esz, dy.stride(0)
    return arr


class _HGTLayerAttention(torch.autograd.Function):
    """Edge-softmax attention of every destination type of one HGT layer, reading q / k_r / v_r as column
    slices of the node types' wide projections ``ys`` and writing dq / dk_r / dv_r straight into the
    matching slices of one gradient buffer per type (no per-relation tensors on either pass).
    Arguments: (heads, hd, targets, *pscales (one [R_g, H] per group, in target/group order), *ys)."""

    @staticmethod
    def forward(ctx, heads, hd, targets, *tensors):
        n_groups = sum(len(t.groups) for t in targets)
        pscales, ys = tensors[:n_groups], tensors[n_groups:]
        for y in ys:
            if y.dim() != 2 or y.stride(1) != 1:
                raise ValueError("wide projections must be row-major 2-D tensors")
        d = hd // heads
        lib = _lib.lib()
        stream = _stream(ys[0])
        esz = ys[0].element_size()
        outs, saved, ps_saved = [], [], []
        gi = 0
        for t in targets:
            yq = ys[t.slot]
            n = yq.shape[0]
            total = None
            for g in t.groups:
                ps = pscales[gi].detach().to(torch.float32).contiguous()
                gi += 1
                out = torch.empty((n, hd), dtype=yq.dtype, device=yq.device)
                row_max = torch.empty((n, heads), dtype=torch.float32, device=yq.device)
                row_den = torch.empty((n, heads), dtype=torch.float32, device=yq.device)
                arr = _hgt_pack_wide(g, ys, hd)
                _lib.check(lib.agnn_hgt_attn_fwd(n, heads, d, _dtype_code(yq), len(g.rels), arr,
                                                 yq.data_ptr() + t.q_off * esz, yq.stride(0), ps.data_ptr(),
                                                 out.data_ptr(), out.stride(0), row_max.data_ptr(),
                                                 row_den.data_ptr(), stream), "agnn_hgt_attn_fwd")
                _lib.count_launches(1)
                saved += [out, row_max, row_den]
                ps_saved.append(ps)
                total = out if total is None else total + out
            outs.append(total)
        ctx.save_for_backward(*ys, *ps_saved, *saved)
        ctx.meta = (heads, hd, targets, len(ys), n_groups, [p.dtype for p in pscales])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        heads, hd, targets, n_y, n_groups, ps_dtypes = ctx.meta
        ys = ctx.saved_tensors[:n_y]
        pss = ctx.saved_tensors[n_y:n_y + n_groups]
        saved = ctx.saved_tensors[n_y + n_groups:]
        d = hd // heads
        lib = _lib.lib()
        stream = _stream(ys[0])
        esz = ys[0].element_size()
        covered = [0] * n_y
        for t in targets:
            covered[t.slot] += hd
            for g in t.groups:
                for slot, *_ in g.rels:
                    covered[slot] += 2 * hd
        dys = [torch.empty_like(y) if c == y.shape[1] else torch.zeros_like(y) for y, c in zip(ys, covered)]
        dpss = []
        gi = 0
        for t, dout in zip(targets, douts):
            yq, dyq = ys[t.slot], dys[t.slot]
            n = yq.shape[0]
            dout = dout.contiguous()
            q_ptr = yq.data_ptr() + t.q_off * esz
            dq_view = dyq[:, t.q_off:t.q_off + hd]
            for j, g in enumerate(t.groups):
                out, row_max, row_den = saved[3 * gi:3 * gi + 3]
                ps = pss[gi]
                r = len(g.rels)
                arr = _hgt_pack_wide(g, ys, hd, dys)
                delta = torch.empty((n, heads), dtype=torch.float32, device=yq.device)
                dq = dq_view if j == 0 else torch.empty((n, hd), dtype=yq.dtype, device=yq.device)
                if n > 0:
                    blocks = lib.agnn_hgt_attn_bwd_dst_blocks(n)
                    partial = torch.empty((blocks, r * heads), dtype=torch.float32, device=yq.device)
                    _lib.check(lib.agnn_hgt_attn_bwd_dst(n, heads, d, _dtype_code(yq), r, arr, q_ptr, yq.stride(0),
                                                         ps.data_ptr(), out.data_ptr(), out.stride(0),
                                                         dout.data_ptr(), dout.stride(0), row_max.data_ptr(),
                                                         row_den.data_ptr(), delta.data_ptr(), dq.data_ptr(),
                                                         dq.stride(0), partial.data_ptr(), stream),
                               "agnn_hgt_attn_bwd_dst")
                    dpss.append(partial.sum(0).view(r, heads).to(ps_dtypes[gi]))
                else:
                    dpss.append(torch.zeros((r, heads), dtype=ps_dtypes[gi], device=yq.device))
                _lib.check(lib.agnn_hgt_attn_bwd_src(heads, d, _dtype_code(yq), r, arr, q_ptr, yq.stride(0),
                                                     ps.data_ptr(), dout.data_ptr(), dout.stride(0),
                                                     row_max.data_ptr(), row_den.data_ptr(), delta.data_ptr(),
                                                     stream), "agnn_hgt_attn_bwd_src")
                _lib.count_launches(2)
                if j > 0:
                    dq_view += dq
                gi += 1
        return (None, None, None, *dpss, *dys)


def hgt_layer_attention(ys, targets, pscales, heads: int, hd: int):
    """See _HGTLayerAttention; returns one [N_dst, hd] tensor per target."""
    return _HGTLayerAttention.apply(heads, hd, list(targets), *pscales, *ys)


# ------------------------------------------------------------------------------
# row-wise L2 normalisation + ReLU (hgnn.py:415, 421-422) and stream helper
# ------------------------------------------------------------------------------

def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def colsum(x: torch.Tensor) -> torch.Tensor:
    """Column sums of a 2-D fp32 matrix (bias gradients): per-block partials + a fixed-order final sum.  Matrices wider
    than the row kernels' 1024 columns (HGT's folded 4864-column projection) are summed in 1024-column slices."""
    if (x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1 or x.shape[1] % 4 or x.stride(0) % 4
            or x.data_ptr() % 16 or x.shape[0] == 0):
        _lib.library_route("colsum of a matrix the row kernels do not take (width % 4, misaligned rows)")
        return x.sum(0)
    lib = _lib.lib()
    rows, cols = x.shape
    out = torch.empty(cols, dtype=torch.float32, device=x.device)
    part = torch.empty((lib.agnn_row_blocks(rows), min(cols, 1024)), dtype=torch.float32, device=x.device)
    for c0 in range(0, cols, 1024):
        w = min(1024, cols - c0)
        _lib.check(lib.agnn_colsum_partials(x.data_ptr() + 4 * c0, x.stride(0), part.data_ptr(),
                                            out.data_ptr() + 4 * c0, rows, w, _stream(x)), "agnn_colsum_partials")
        _lib.count_launches(2)
    return out


def prepare_grad(g: torch.Tensor, relu_out: Optional[torch.Tensor] = None, want_colsum: bool = False,
                 f16: bool = False, amax: Optional[torch.Tensor] = None):
    """What a projection's backward needs from its incoming gradient, in one pass (agnn_grad_prepare):
    ``g' = g * [relu_out > 0]`` as a GEMM operand and, optionally, the column sums of ``g'``.
    Returns ``(operand, colsum or None)``.  ``f16``: the operand is the fp16 pair of the F16X3 mode; its scale comes
    from ``amax`` (a bound of max |g| its producer reported), else from one extra pass over ``g``."""
    if amax is None:
        amax = linalg.known_amax(g)
    rows, cols = g.shape
    ok = (g.dtype == torch.float32 and g.is_cuda and rows > 0 and cols % 4 == 0
          and cols <= 1024 and g.stride(1) == 1 and g.stride(0) % 4 == 0 and g.data_ptr() % 16 == 0)
    if ok and relu_out is not None:
        ok = (relu_out.dtype == torch.float32 and relu_out.shape == g.shape and relu_out.stride(1) == 1
              and relu_out.stride(0) % 4 == 0 and relu_out.data_ptr() % 16 == 0)
    if not ok:
        if relu_out is not None:
            g = linalg.relu_backward(g, relu_out)
        g = g.contiguous()
        # wide gradients (e.g. HGT's folded 4864-column projection) miss the one-pass kernel but keep the operand form
        op = linalg.split_f16(g, amax) if f16 and g.is_cuda and linalg.f16_ok(g) else linalg.prepare(g)
        return op, (colsum(g) if want_colsum else None)
    lib = _lib.lib()
    f16 = f16 and cols % 8 == 0
    buf = torch.empty((2, rows, cols), dtype=torch.float16 if f16 else torch.float32, device=g.device)
    part = out = None
    if want_colsum:
        part = torch.empty((lib.agnn_row_blocks(rows), cols), dtype=torch.float32, device=g.device)
        out = torch.empty(cols, dtype=torch.float32, device=g.device)
    if f16:
        if amax is None:
            amax = linalg.amax_of(g)
        _lib.check(lib.agnn_grad_prepare_f16(g.data_ptr(), g.stride(0),
                                             relu_out.data_ptr() if relu_out is not None else None,
                                             relu_out.stride(0) if relu_out is not None else 0, amax.data_ptr(),
                                             buf[0].data_ptr(), buf[1].data_ptr(), cols,
                                             part.data_ptr() if part is not None else None,
                                             out.data_ptr() if out is not None else None, rows, cols, _stream(g)),
                   "agnn_grad_prepare_f16")
        _lib.count_launches(2 if want_colsum else 1)
        return linalg.SplitH(buf[0], buf[1], amax), out
    _lib.check(lib.agnn_grad_prepare(g.data_ptr(), g.stride(0), relu_out.data_ptr() if relu_out is not None else None,
                                     relu_out.stride(0) if relu_out is not None else 0, buf[0].data_ptr(),
                                     buf[1].data_ptr(), cols, part.data_ptr() if part is not None else None,
                                     out.data_ptr() if out is not None else None, rows, cols, _stream(g)),
               "agnn_grad_prepare")
    _lib.count_launches(2 if want_colsum else 1)
    return linalg.Split(buf[0], buf[1]), out


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False) -> torch.Tensor:
    """``nn.Linear`` arithmetic (optionally with the ReLU behind it in the GEMM epilogue) for any leading shape: one
    member of a fused projection stage (fused.stage_group), which pads feature counts that are not multiples of 8
    (e.g. the 25 + 128 note features, analysisgnn/models/analysis.py:574) to the TMA row rule."""
    from . import fused
    return fused.stage_group([x], [weight], [bias], relu=relu)[0]


class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, ignore_index, smoothing):
        rows, cols = logits.shape
        dev = logits.device
        lib = _lib.lib()
        lse = torch.empty(rows, dtype=torch.float32, device=dev)
        partials = torch.empty((lib.agnn_ce_blocks(rows), 2), dtype=torch.float32, device=dev)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        _lib.check(lib.agnn_softmax_ce_fwd(logits.data_ptr(), logits.stride(0), labels.data_ptr(), rows, cols,
                                           smoothing, ignore_index, lse.data_ptr(), partials.data_ptr(),
                                           out.data_ptr(), _stream(logits)), "agnn_softmax_ce_fwd")
        _lib.count_launches(2)
        ctx.save_for_backward(logits, labels, lse, out)
        ctx.args = (ignore_index, smoothing)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        logits, labels, lse, out = ctx.saved_tensors
        ignore_index, smoothing = ctx.args
        rows, cols = logits.shape
        g = g.to(torch.float32).contiguous()
        # rows padded to 16 bytes for the head's backward GEMMs (zeroed padding), |dx| <= |g / rows that count|
        pad = (-cols) % 8
        full = torch.empty((rows, cols + pad), dtype=torch.float32, device=logits.device)
        amax = linalg.new_amax(logits.device)
        _lib.check(_lib.lib().agnn_softmax_ce_bwd_padded(logits.data_ptr(), logits.stride(0), labels.data_ptr(),
                                                         lse.data_ptr(), rows, cols, smoothing, ignore_index,
                                                         out.data_ptr(), g.data_ptr(), full.data_ptr(), full.stride(0),
                                                         cols + pad, amax.data_ptr(), _stream(logits)),
                   "agnn_softmax_ce_bwd")
        _lib.count_launches(1)
        dx = full[:, :cols] if pad else full
        linalg.tag_amax(full, amax)
        linalg.tag_amax(dx, amax)
        if pad:
            dx._agnn_padded = full
        return dx, None, None, None


def cross_entropy(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100,
                  label_smoothing: float = 0.0) -> torch.Tensor:
    """``F.cross_entropy(logits, labels, ignore_index=..., label_smoothing=...)`` (mean over the rows that count)
    for 2-D fp32 logits (any row stride) and int64 labels."""
    if not logits.is_cuda:
        raise _lib.AgnnError("analysisgnn_b200 has no CPU path: tensors must live on a CUDA device")
    if logits.dim() != 2 or labels.dim() != 1 or labels.shape[0] != logits.shape[0]:
        raise ValueError("cross_entropy expects logits [N, C] and labels [N]")
    if logits.dtype != torch.float32:
        logits = logits.float()
    if logits.stride(1) != 1:
        logits = logits.contiguous()
    return _CrossEntropy.apply(logits, labels.to(torch.int64).contiguous(), int(ignore_index), float(label_smoothing))


class _SmallEmbedding(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, weight):
        ctx.save_for_backward(idx)
        ctx.shape = weight.shape
        return weight.index_select(0, idx.reshape(-1)).view(*idx.shape, weight.shape[1])

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        n_emb, dim = ctx.shape
        g = g.reshape(-1, dim)
        if g.stride(1) != 1:
            g = g.contiguous()
        rows = g.shape[0]
        lib = _lib.lib()
        partials = torch.empty((lib.agnn_embedding_bwd_blocks(rows), n_emb * dim), dtype=torch.float32, device=g.device)
        dw = torch.empty((n_emb, dim), dtype=torch.float32, device=g.device)
        _lib.check(lib.agnn_embedding_bwd(g.data_ptr(), g.stride(0), idx.reshape(-1).contiguous().data_ptr(), rows, dim,
                                          n_emb, partials.data_ptr(), dw.data_ptr(), _stream(g)), "agnn_embedding_bwd")
        _lib.count_launches(2)
        return None, dw


def embedding(idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """``F.embedding`` for the small fp32 tables of the encoder input (no padding_idx / max_norm); larger
    tables and other dtypes go to ATen."""
    if (not weight.is_cuda or weight.dtype != torch.float32 or weight.numel() > 3072 or idx.dtype != torch.int64
            or not weight.is_contiguous()):
        return torch.nn.functional.embedding(idx, weight)
    return _SmallEmbedding.apply(idx, weight)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        rows, cols = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().agnn_layernorm_fwd(x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(),
                                                 y.data_ptr(), y.stride(0), mean.data_ptr(), rstd.data_ptr(), rows,
                                                 cols, eps, _stream(x)), "agnn_layernorm_fwd")
        _lib.count_launches(1)
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        rows, cols = x.shape
        dy = dy.contiguous()
        lib = _lib.lib()
        dx = torch.empty_like(x)
        blocks = lib.agnn_row_blocks(rows)
        part = torch.empty((2, blocks, cols), dtype=torch.float32, device=x.device)
        sums = torch.empty((2, cols), dtype=torch.float32, device=x.device)
        _lib.check(lib.agnn_layernorm_bwd(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), gamma.data_ptr(),
                                          mean.data_ptr(), rstd.data_ptr(), dx.data_ptr(), dx.stride(0),
                                          part[0].data_ptr(), part[1].data_ptr(), sums[0].data_ptr(),
                                          sums[1].data_ptr(), rows, cols, _stream(x)), "agnn_layernorm_bwd")
        _lib.count_launches(2)
        return dx, sums[0], sums[1], None


def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """``nn.LayerNorm`` over the last dimension (fp32, width a multiple of 4 up to 1024; other cases go
    to the library kernel)."""
    if not x.is_cuda:
        raise _lib.AgnnError("analysisgnn_b200 has no CPU path: tensors must live on a CUDA device")
    cols = x.shape[-1]
    if x.dtype != torch.float32 or cols % 4 or cols > 1024 or gamma is None or beta is None or x.numel() == 0:
        if x.numel():
            _lib.library_route("layer_norm outside the row kernel's shapes (fp32, width % 4 == 0, <= 1024)")
        return torch.nn.functional.layer_norm(x, (cols,), gamma, beta, eps)
    x2 = x.reshape(-1, cols)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    return _LayerNorm.apply(x2, gamma.contiguous(), beta.contiguous(), eps).reshape(x.shape)


class _L2NormRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, relu_first):
        rows, cols = x.shape
        y = torch.empty_like(x)
        inv = torch.empty(rows, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().agnn_l2norm_relu_fwd(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0),
                                                   inv.data_ptr(), rows, cols, int(relu_first), 1e-12, _stream(x)),
                   "agnn_l2norm_relu_fwd")
        _lib.count_launches(1)
        ctx.save_for_backward(x, inv)
        ctx.relu_first = relu_first
        return y

    @staticmethod
    def backward(ctx, dy):
        x, inv = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        rows, cols = x.shape
        _lib.check(_lib.lib().agnn_l2norm_relu_bwd(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0),
                                                   inv.data_ptr(), dx.data_ptr(), dx.stride(0), rows, cols,
                                                   int(ctx.relu_first), _stream(x)), "agnn_l2norm_relu_bwd")
        _lib.count_launches(1)
        return dx, None


def l2norm_relu(h, relu_first: bool):
    """``relu_first``: ``normalize(relu(h))`` (hgnn.py:415, 431); else ``relu(normalize(h))``
    (hgnn.py:421-422).  ``normalize`` = ``h / max(||h||_2, 1e-12)`` per row."""
    if (h.is_cuda and h.dtype == torch.float32 and h.dim() == 2 and h.shape[1] % 4 == 0 and h.shape[1] <= 1024
            and h.shape[0] > 0):
        return _L2NormRelu.apply(h.contiguous(), bool(relu_first))
    if h.numel():
        _lib.library_route("l2norm_relu outside the row kernel's shapes (fp32, width % 4 == 0, <= 1024)")
    if relu_first:
        return torch.nn.functional.normalize(torch.relu(h), p=2.0, dim=-1)
    return torch.relu(torch.nn.functional.normalize(h, p=2.0, dim=-1))


_side_streams: Dict[Tuple[int, int, bool], "torch.cuda.Stream"] = {}


def side_stream(device, which: int = 0, high_priority: bool = False) -> "torch.cuda.Stream":
    """Stream ``which`` beside the caller's: 0 = the sequence branch / the measure layers, 1 = weight-only work (the
    composite projection weights of the HGT layers).  ``high_priority``: pending CTAs of this stream's kernels are
    placed before those of the default-priority streams (kept by the kernel nodes of a captured graph); for a
    branch that is a latency chain of narrow kernels beside wide ones."""
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    key = (idx, which, bool(high_priority))
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=idx, priority=-1 if high_priority else 0)
    return _side_streams[key]


class _GRULayer(torch.autograd.Function):
    """One (bi)directional GRU layer, batch first, h0 = 0.  Arguments: x [B, T, C], then per direction
    (w_ih [3H, C], w_hh [3H, H], b_ih [3H], b_hh [3H]).  Input projections and all weight / input
    gradients are tensor-core GEMMs; agnn_gru_fwd / _bwd run only the time recurrence."""

    @staticmethod
    def forward(ctx, x, opts, *params):
        n_dir = len(params) // 4
        b, t, c = x.shape
        h = params[1].shape[1]
        x2 = x.reshape(b * t, c)
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        p, site = opts.get("dropout", (0.0, 0))          # nn.GRU's inter-layer dropout, applied to this layer's input
        f16 = (linalg.parity_operands() == "f16" and x2.shape[0] >= linalg.F16_MIN_ROWS and linalg.f16_ok(x2))
        if f16:
            xs = linalg.split_f16(x2, opts.get("x_amax"), dropout=(p, site) if p > 0 else None)
        else:
            if p > 0:
                from . import fused
                x2 = fused.dropout_apply(x2, p, site)
            xs = linalg.prepare(x2)
        whh = [params[4 * d + 1].contiguous() for d in range(n_dir)]
        bhh = [params[4 * d + 3].contiguous() for d in range(n_dir)]
        # the input projections of all directions: one grouped launch
        gis = linalg.linear_group([xs] * n_dir, [params[4 * d] for d in range(n_dir)],
                                  [params[4 * d + 2] for d in range(n_dir)])
        ctx.dropout = (p, site)
        out = torch.empty((b, t, n_dir * h), dtype=x.dtype, device=x.device)
        gates = [torch.empty((b, t, 4 * h), dtype=x.dtype, device=x.device) for _ in range(n_dir)]
        _lib.check(_lib.lib().agnn_gru_fwd(b, t, h, n_dir, _lib.ptr_array(gis), _lib.ptr_array(whh),
                                           _lib.ptr_array(bhh), out.data_ptr(), _lib.ptr_array(gates), _stream(x)),
                   "agnn_gru_fwd")
        _lib.count_launches(t if _lib.lib().agnn_gru_supported(h) == _lib.GRU_STEPWISE else 1)
        ctx.save_for_backward(*linalg.pack(xs), out, *gates, *[params[4 * d] for d in range(n_dir)], *whh)
        ctx.n_dir, ctx.dims = n_dir, (b, t, c, h)
        ctx.x_amax = getattr(xs, "amax", None)
        return out

    @staticmethod
    def backward(ctx, dout):
        n_dir = ctx.n_dir
        b, t, c, h = ctx.dims
        saved = ctx.saved_tensors
        xs = linalg.unpack(saved[0], saved[1], ctx.x_amax)
        out = saved[2]
        gates = list(saved[3:3 + n_dir])
        w_ih = list(saved[3 + n_dir:3 + 2 * n_dir])
        whh = list(saved[3 + 2 * n_dir:3 + 3 * n_dir])
        dout = dout.contiguous()
        dgi = [torch.empty((b * t, 3 * h), dtype=out.dtype, device=out.device) for _ in range(n_dir)]
        dgh = [torch.empty((b * t, 3 * h), dtype=out.dtype, device=out.device) for _ in range(n_dir)]
        g_amax = [linalg.new_amax(out.device) for _ in range(n_dir)]
        if _lib.lib().agnn_gru_supported(h) == _lib.GRU_STEPWISE:
            # wide hidden sizes (MetricalConvLayer's GRU(512, 512)): one launch per time step, W_hh^T read K-major
            whh_t = [w.t().contiguous() for w in whh]
            carry = torch.empty((n_dir, b, h), dtype=out.dtype, device=out.device)
            _lib.check(_lib.lib().agnn_gru_bwd_stepwise(b, t, h, n_dir, _lib.ptr_array(whh_t), out.data_ptr(),
                                                        _lib.ptr_array(gates), dout.data_ptr(), _lib.ptr_array(dgi),
                                                        _lib.ptr_array(dgh), _lib.ptr_array(g_amax),
                                                        carry.data_ptr(), _stream(out)),
                       "agnn_gru_bwd_stepwise")
            _lib.count_launches(t)
        else:
            _lib.check(_lib.lib().agnn_gru_bwd_amax(b, t, h, n_dir, _lib.ptr_array(whh), out.data_ptr(),
                                                    _lib.ptr_array(gates), dout.data_ptr(), _lib.ptr_array(dgi),
                                                    _lib.ptr_array(dgh), _lib.ptr_array(g_amax), _stream(out)),
                       "agnn_gru_bwd")
            _lib.count_launches(1)
        for d in range(n_dir):                               # max |dgi| bounds dgh as well
            linalg.tag_amax(dgi[d], g_amax[d])
            linalg.tag_amax(dgh[d], g_amax[d])
        grads = []
        dx = None
        gi_ops, gh_ops, hp_ops = [], [], []
        out2 = out.view(b * t, n_dir * h)
        for d in range(n_dir):
            # dgi meets xs in the grad-weight GEMM: same operand form
            gi_ops.append((linalg.prepare_auto if isinstance(xs, linalg.SplitH) else linalg.prepare)(dgi[d]))
            gh_ops.append(linalg.prepare_auto(dgh[d]))               # (split_f16 finds the tags: no amax pass)
            hd2 = out2[:, d * h:(d + 1) * h]                         # this direction's states, [B T, H] strided view
            if isinstance(gh_ops[-1], linalg.SplitH) and linalg.f16_ok(hd2):
                # h_{t-1} (forward direction) / h_{t+1} (reverse): the operand pair straight from the GRU output
                hp_ops.append(linalg.split_f16(hd2, linalg.const_amax(out.device, 1.0), shift=(t, 1 if d == 0 else -1)))
                continue
            hd = out[:, :, d * h:(d + 1) * h]
            h_prev = torch.zeros((b, t, h), dtype=out.dtype, device=out.device)
            if t > 1:
                if d == 0:
                    h_prev[:, 1:] = hd[:, :-1]
                else:
                    h_prev[:, :-1] = hd[:, 1:]
            hp_ops.append(h_prev.reshape(b * t, h))
        # weight gradients of all directions (input and recurrent) in one grouped launch
        dws = linalg.mm_tn_group(gi_ops + gh_ops, [xs] * n_dir + hp_ops)
        if ctx.needs_input_grad[0]:
            dx = linalg.mm(gi_ops[0], w_ih[0])
            for d in range(1, n_dir):
                dx = linalg.mm(gi_ops[d], w_ih[d], out=dx, accumulate=True)
            p, site = ctx.dropout
            if p > 0:
                from . import fused
                dx = fused.dropout_apply(dx, p, site)
        for d in range(n_dir):
            grads += [dws[d], dws[n_dir + d], colsum(dgi[d]), colsum(dgh[d])]
        return (dx.reshape(b, t, c) if dx is not None else None, None, *grads)


def gru_layer(x: torch.Tensor, params, dropout=None) -> torch.Tensor:
    """``params``: flat list of (w_ih, w_hh, b_ih, b_hh) per direction.  ``dropout = (p, site)``: nn.GRU's
    inter-layer dropout on this layer's INPUT (counter-based mask, fused into the operand split).  The output is tagged
    with the bound |h| <= 1 (a GRU state is a convex combination of tanh values and earlier states)."""
    opts = {"x_amax": linalg.known_amax(x)}
    if dropout is not None and dropout[0] > 0:
        opts["dropout"] = dropout
    out = _GRULayer.apply(x, opts, *params)
    return linalg.tag_amax(out, linalg.const_amax(out.device, 1.0))


def gru_supported(x: torch.Tensor, hidden: int) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.shape[0] > 0 and x.shape[1] > 0
            and bool(_lib.lib().agnn_gru_supported(int(hidden))))
