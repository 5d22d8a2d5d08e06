"""Dense contractions used by the fused layers.

The per-relation / per-node-type projections are the only GEMM-shaped work on the hot path.  All of them run on
this repo's sm_100a tensor-core GEMM (csrc/gemm.cu, ``agnn_gemm``): fp32 tensors in a three-product split mode
(fp32 parity) -- 3xTF32 (``Split``), or 3 x fp16 with one power-of-two scale per tensor (``SplitH``,
``parity_operands() == "f16"``: same accuracy at twice the MMA rate) -- bf16 tensors in the bf16 mode.

There is no library GEMM and no CPU path: operands ``agnn_gemm`` cannot read as they are (rows that are not 16-byte
multiples, mixed operand forms) are first copied into padded buffers (``stats["repacked_gemms"]``; an error under
``_lib.set_strict(True)``, which bench.py sets -- a hot path must show 0), and CPU tensors raise.
"""
from __future__ import annotations

import os
from typing import Optional, Union

import torch

from . import _lib

timer = None      # set to an ops.KernelTimer by bench.py to time every agnn_gemm launch
# how many contractions had to copy their operands into padded buffers first because agnn_gemm does not take them as
# they are: misaligned rows, or a TF32 pair meeting an fp16 pair.  bench.py reports it; a hot path should show 0.
stats = {"repacked_gemms": 0}


def _need_cuda(t):
    if not t.is_cuda:
        raise _lib.AgnnError("analysisgnn_b200 has no CPU path: tensors must live on a CUDA device")


class Split:
    """An fp32 matrix as TF32-exact ``hi`` + ``lo`` parts (``agnn_split_tf32``): the operand form of the
    3xTF32 mode.  Keeping the pair lets one split feed several GEMMs (forward, grad-weight)."""
    __slots__ = ("hi", "lo")

    def __init__(self, hi, lo):
        self.hi, self.lo = hi, lo

    @property
    def shape(self):
        return self.hi.shape

    @property
    def dtype(self):
        return self.hi.dtype

    @property
    def device(self):
        return self.hi.device

    def merged(self):
        return self.hi + self.lo


class SplitH:
    """An fp32 matrix as the fp16 ``hi`` + ``lo`` pair of ``s * x`` (``agnn_split_f16``; ``s`` = the power of two
    derived from the device scalar ``amax``): the operand form of the F16X3 mode -- fp32-grade like ``Split``, at
    twice the tensor-core rate and half the bytes."""
    __slots__ = ("hi", "lo", "amax")

    def __init__(self, hi, lo, amax):
        self.hi, self.lo, self.amax = hi, lo, amax

    @property
    def shape(self):
        return self.hi.shape

    @property
    def dtype(self):
        return torch.float32

    @property
    def device(self):
        return self.hi.device

    def merged(self):
        return (self.hi.float() + self.lo.float()) / f16_scale(self.amax)


def f16_scale(amax: torch.Tensor) -> torch.Tensor:
    """The scale the kernels derive from an amax scalar (csrc/common.cuh::f16_scale_of), as a device tensor."""
    e = torch.frexp(amax)[1].to(torch.float32) - 1       # floor(log2 amax), exact (amax = m 2^e', m in [0.5, 1))
    k = (13 - e).clamp(-100, 100)
    ok = (amax > 1.1754944e-38) & torch.isfinite(amax)
    return torch.where(ok, torch.exp2(k), torch.ones_like(amax))


Operand = Union[torch.Tensor, Split, SplitH]

_PARITY_OPERANDS = os.environ.get("AGNN_PARITY_OPERANDS", "f16")


def parity_operands() -> str:
    return _PARITY_OPERANDS


def set_parity_operands(name: str) -> None:
    """Operand form of the fp32 parity mode in the fused message-passing layers: ``"tf32"`` (3xTF32) or ``"f16"``
    (3 x fp16 with per-tensor power-of-two scales; same accuracy, twice the MMA rate)."""
    global _PARITY_OPERANDS
    if name not in ("tf32", "f16"):
        raise ValueError(name)
    _PARITY_OPERANDS = name


_amax_pools = {}         # (device, stream) -> [zeroed scalars, next free]: one fill launch per 256 scalars
F16_MIN_ROWS = 16384     # rows from which the fp16 operand form pays for its amax pass


def new_amax(device) -> torch.Tensor:
    """A zeroed device scalar for ``amax_into``.  Slots come from a per-stream pool that ``begin_step`` renews, so a
    (captured) training step zeroes all of its scalars with one fill per stream."""
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    pool = _amax_pools.get(key)
    if pool is None or pool[1] >= pool[0].numel():
        pool = _amax_pools[key] = [torch.zeros(256, dtype=torch.float32, device=device), 0]
    slot = pool[0][pool[1]:pool[1] + 1]
    pool[1] += 1
    return slot


def amax_into(amax: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """``amax = max(amax, max |x|)`` on the device (agnn_amax); several tensors may share one scalar."""
    _need_cuda(x)
    if x.numel() == 0:
        return amax
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1 or x.shape[1] % 4 or x.stride(0) % 4 \
            or x.data_ptr() % 16:
        _lib.library_route("amax of a matrix agnn_amax does not take (width % 4, misaligned rows)")
        return torch.maximum(amax, x.detach().abs().max().float().reshape(1), out=amax)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(_lib.lib().agnn_amax(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), amax.data_ptr(), stream),
               "agnn_amax")
    _lib.count_launches(1)
    return amax


def f16_ok(x: torch.Tensor) -> bool:
    return (x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1 and x.shape[1] % 8 == 0
            and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0 and x.shape[0] > 0)


def split_f16(x: torch.Tensor, amax: Optional[torch.Tensor] = None) -> SplitH:
    """hi = fp16(s x), lo = fp16(s x - hi); ``amax`` (device scalar >= max |x| / 4) is measured when not given."""
    _need_cuda(x)
    if not f16_ok(x):
        raise ValueError("split_f16 needs an fp32 matrix with unit column stride, 16-byte aligned rows and a column "
                         "count that is a multiple of 8")
    if amax is None:
        amax = amax_into(new_amax(x.device), x)
    rows, cols = x.shape
    buf = torch.empty((2, rows, cols), dtype=torch.float16, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(_lib.lib().agnn_split_f16(x.data_ptr(), rows, cols, x.stride(0), amax.data_ptr(), buf[0].data_ptr(),
                                         buf[1].data_ptr(), cols, stream), "agnn_split_f16")
    _lib.count_launches(1)
    return SplitH(buf[0], buf[1], amax)


def _rows_ok(t: torch.Tensor) -> bool:
    return (t.dim() == 2 and t.stride(1) == 1 and (t.stride(0) * t.element_size()) % 16 == 0
            and t.data_ptr() % 16 == 0 and t.shape[0] > 0 and t.shape[1] > 0)


def split(x: torch.Tensor) -> Split:
    """hi = rna_tf32(x), lo = rna_tf32(x - hi) (contiguous copies)."""
    _need_cuda(x)
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1 or x.shape[1] % 4 or x.stride(0) % 4 \
            or x.data_ptr() % 16:
        raise ValueError("split needs an fp32 matrix with unit column stride and 16-byte aligned rows")
    rows, cols = x.shape
    buf = torch.empty((2, rows, cols), dtype=torch.float32, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(_lib.lib().agnn_split_tf32(x.data_ptr(), rows, cols, x.stride(0), buf[0].data_ptr(), buf[1].data_ptr(),
                                          cols, stream), "agnn_split_tf32")
    _lib.count_launches(1)
    return Split(buf[0], buf[1])


def prepare(x: torch.Tensor) -> Operand:
    """Pre-split an fp32 operand that several GEMMs will read (no-op for other dtypes / unaligned rows)."""
    if isinstance(x, torch.Tensor) and x.dtype == torch.float32 and _rows_ok(x) and x.shape[1] % 4 == 0:
        return split(x)
    return x


def prepare_auto(x: torch.Tensor) -> Operand:
    """``prepare`` in the operand form of the current parity mode: the fp16 pair for large matrices when
    ``parity_operands() == "f16"`` (every GEMM that reads it then runs in F16X3), the TF32 pair otherwise."""
    if (_PARITY_OPERANDS == "f16" and isinstance(x, torch.Tensor) and x.is_cuda and x.shape[0] >= F16_MIN_ROWS
            and f16_ok(x)):
        return split_f16(x)
    return prepare(x)


def pack(x: Operand):
    """(tensor, tensor-or-None) for ``save_for_backward``; ``unpack`` restores the operand (a ``SplitH`` also needs
    its ``amax``, which the caller keeps)."""
    return (x.hi, x.lo) if isinstance(x, (Split, SplitH)) else (x, None)


def unpack(first, second, amax=None) -> Operand:
    if second is None:
        return first
    return SplitH(first, second, amax) if first.dtype == torch.float16 else Split(first, second)


# Splits of small operands (weights) made inside a step: a weight is read by the forward GEMM and again by the
# grad-input GEMM.  Entries keep the source tensor alive (its address cannot be recycled under a live key) and are
# dropped by begin_step(), which every optimizer step / captured step function calls.
_split_cache = {}
_SPLIT_CACHE_MAX_ELEMS = 4 * 1024 * 1024


def begin_step() -> None:
    """Forget cached weight splits and the amax scalars of the last step (call once per training step, before the
    forward)."""
    _split_cache.clear()
    _amax_pools.clear()


_SPLIT_CACHE_MAX_ENTRIES = 512     # a step makes a few dozen; the cap only matters to callers that never call begin_step


def _remember(key, value):
    _split_cache[key] = value
    if len(_split_cache) > _SPLIT_CACHE_MAX_ENTRIES:
        for old in list(_split_cache)[:_SPLIT_CACHE_MAX_ENTRIES // 2]:      # oldest first (insertion order)
            del _split_cache[old]
    return value


def _cached_split(x: torch.Tensor) -> Split:
    if x.numel() > _SPLIT_CACHE_MAX_ELEMS:
        return split(x)
    key = (x.data_ptr(), tuple(x.shape), x.stride(0), x._version)
    hit = _split_cache.get(key)
    if hit is None:
        hit = _remember(key, (x, split(x)))
    return hit[1]


def _cached_split_f16(x: torch.Tensor) -> SplitH:
    if x.numel() > _SPLIT_CACHE_MAX_ELEMS:
        return split_f16(x)
    key = ("f16", x.data_ptr(), tuple(x.shape), x.stride(0), x._version)
    hit = _split_cache.get(key)
    if hit is None:
        hit = _remember(key, (x, split_f16(x)))
    return hit[1]


def _as_operand(x: Operand, f16: bool = False):
    """(hi, lo, precision, amax) for agnn_gemm, or None when the tensor cannot take the tcgen05 route.  ``f16``: the
    other operand is an fp16 pair, so a plain fp32 tensor is split the same way."""
    if isinstance(x, SplitH):
        return x.hi, x.lo, _lib.GEMM_F16X3, x.amax
    if isinstance(x, Split):
        return x.hi, x.lo, _lib.GEMM_TF32X3, None
    if x.dtype == torch.bfloat16:
        return (x, None, _lib.GEMM_BF16, None) if _rows_ok(x) else None
    if f16:
        if not f16_ok(x):
            return None
        s = _cached_split_f16(x)
        return s.hi, s.lo, _lib.GEMM_F16X3, s.amax
    if x.dtype == torch.float32 and _rows_ok(x) and x.shape[1] % 4 == 0:
        s = _cached_split(x)
        return s.hi, s.lo, _lib.GEMM_TF32X3, None
    return None


def plain(x: Operand) -> torch.Tensor:
    return x.merged() if isinstance(x, (Split, SplitH)) else x


_plain = plain


def _gemm(a: Operand, a_layout: int, b: Operand, b_layout: int, m: int, n: int, k: int, bias, flags: int,
          out: Optional[torch.Tensor], split_k: Optional[int] = None) -> Optional[torch.Tensor]:
    """agnn_gemm wrapper; returns None if the operands are not eligible as they are (``_repacked`` then copies
    them into padded buffers)."""
    f16 = isinstance(a, SplitH) or isinstance(b, SplitH)
    oa, ob = _as_operand(a, f16), _as_operand(b, f16)
    if oa is None or ob is None or oa[2] != ob[2]:
        return None
    prec = oa[2]
    dev = oa[0].device
    out_dtype = torch.bfloat16 if prec == _lib.GEMM_BF16 else torch.float32
    if out is None:
        # rows padded to 16 bytes so that the TMA-store epilogue applies (e.g. the 185-class head)
        pad = (-n) % (8 if out_dtype == torch.bfloat16 else 4)
        out = torch.empty((m, n + pad), dtype=out_dtype, device=dev)[:, :n]
    elif out.dtype != out_dtype or out.stride(1) != 1:
        return None
    if prec == _lib.GEMM_BF16:
        flags |= _lib.GEMM_OUT_BF16
    if bias is not None:
        bias = bias.float().contiguous()
    lib = _lib.lib()
    if split_k is None:
        split_k = lib.agnn_gemm_split_k(prec, m, n, k)
    ws_bytes = lib.agnn_gemm_workspace(prec, m, n, k, split_k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
    stream = torch.cuda.current_stream(dev).cuda_stream

    def run():
        if prec == _lib.GEMM_F16X3:
            _lib.check(lib.agnn_gemm_scaled(prec, a_layout, b_layout, m, n, k, oa[0].data_ptr(), oa[1].data_ptr(),
                                            oa[0].stride(0), oa[3].data_ptr(), ob[0].data_ptr(), ob[1].data_ptr(),
                                            ob[0].stride(0), ob[3].data_ptr(), out.data_ptr(), out.stride(0),
                                            bias.data_ptr() if bias is not None else None, flags, split_k,
                                            ws.data_ptr() if ws is not None else None, ws_bytes, stream),
                       "agnn_gemm_scaled")
            return
        _lib.check(lib.agnn_gemm(prec, a_layout, b_layout, m, n, k, oa[0].data_ptr(),
                                 oa[1].data_ptr() if oa[1] is not None else None, oa[0].stride(0), ob[0].data_ptr(),
                                 ob[1].data_ptr() if ob[1] is not None else None, ob[0].stride(0), out.data_ptr(),
                                 out.stride(0), bias.data_ptr() if bias is not None else None, flags, split_k,
                                 ws.data_ptr() if ws is not None else None, ws_bytes, stream), "agnn_gemm")

    if timer is not None:                       # bench.py: per-launch CUDA events, algorithmic flops = 2 M N K
        timer.launch("gemm", 2 * m * n * k, dev, run, tag=(a_layout, b_layout, m, n, k, split_k))
    else:
        run()
    _lib.count_launches(2 if split_k > 1 else 1)
    return out


def _pad2(t: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    """Contiguous zero-padded copy [rows, cols] of a 2-D tensor (elementwise copy kernels, not a GEMM)."""
    if t.shape[0] == rows and t.shape[1] == cols and t.is_contiguous() and t.data_ptr() % 16 == 0:
        return t
    out = torch.zeros((rows, cols), dtype=t.dtype, device=t.device)
    out[:t.shape[0], :t.shape[1]] = t
    return out


def _repacked(a: Operand, a_layout: int, b: Operand, b_layout: int, m: int, n: int, k: int, bias, flags: int,
              out: Optional[torch.Tensor]) -> torch.Tensor:
    """The same contraction for operands agnn_gemm cannot read in place: plain copies padded to 16-byte rows (zeros
    do not change the result), the product on the tensor core as usual, the result copied / added into ``out``."""
    from . import _lib as lib_mod
    lib_mod.library_route("agnn_gemm operand repack (misaligned rows or mixed operand forms)", counter=stats,
                          key="repacked_gemms")
    ap, bp = plain(a), plain(b)
    if ap.dtype != bp.dtype:
        ap, bp = ap.float(), bp.float()
    mult = 8 if ap.dtype == torch.bfloat16 else 4
    up = lambda v: (v + mult - 1) // mult * mult
    mp, np_, kp = up(m), up(n), up(k)
    ap = _pad2(ap, *((kp, mp) if a_layout == _lib.MN_MAJOR else (m, kp)))
    bp = _pad2(bp, *((kp, np_) if b_layout == _lib.MN_MAJOR else (n, kp)))
    m2 = mp if a_layout == _lib.MN_MAJOR else m
    n2 = np_ if b_layout == _lib.MN_MAJOR else n
    if bias is not None and n2 != n:
        bias = torch.nn.functional.pad(bias.float(), (0, n2 - n))
    y = _gemm(ap, a_layout, bp, b_layout, m2, n2, kp, bias, flags & ~_lib.GEMM_ACCUMULATE, None)
    if y is None:
        raise _lib.AgnnError("agnn_gemm refused padded operands")
    y = y[:m, :n]
    if out is None:
        return y
    if flags & _lib.GEMM_ACCUMULATE:
        return out.add_(y.to(out.dtype))
    return out.copy_(y)


def linear(x: Operand, weight: Operand, bias=None, relu: bool = False):
    """``x @ weight.T + bias`` (optionally ReLU'd), ``weight`` [out, in]."""
    _need_cuda(x if isinstance(x, torch.Tensor) else x.hi)
    m, k = x.shape
    n = weight.shape[0]
    flags = _lib.GEMM_RELU if relu else 0
    y = _gemm(x, _lib.K_MAJOR, weight, _lib.K_MAJOR, m, n, k, bias, flags, None)
    return y if y is not None else _repacked(x, _lib.K_MAJOR, weight, _lib.K_MAJOR, m, n, k, bias, flags, None)


def mm(a: Operand, b: Operand, out=None, accumulate: bool = False):
    """``a @ b`` with ``b`` stored [K, N]; with ``out`` and ``accumulate`` adds into ``out`` in place."""
    _need_cuda(a if isinstance(a, torch.Tensor) else a.hi)
    m, k = a.shape
    n = b.shape[1]
    flags = _lib.GEMM_ACCUMULATE if accumulate else 0
    y = _gemm(a, _lib.K_MAJOR, b, _lib.MN_MAJOR, m, n, k, None, flags, out)
    return y if y is not None else _repacked(a, _lib.K_MAJOR, b, _lib.MN_MAJOR, m, n, k, None, flags, out)


def mm_tn(a: Operand, b: Operand):
    """``a.T @ b`` (weight gradients): ``a`` [R, M], ``b`` [R, N], reduction over the R rows (split-K)."""
    _need_cuda(a if isinstance(a, torch.Tensor) else a.hi)
    r, m = a.shape
    n = b.shape[1]
    y = _gemm(a, _lib.MN_MAJOR, b, _lib.MN_MAJOR, m, n, r, None, 0, None)
    return y if y is not None else _repacked(a, _lib.MN_MAJOR, b, _lib.MN_MAJOR, m, n, r, None, 0, None)


def relu_backward(grad, out):
    """Gradient through ``out = relu(.)`` given the saved output."""
    return torch.where(out > 0, grad, torch.zeros((), dtype=grad.dtype, device=grad.device))
