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


def tag_amax(t: torch.Tensor, amax: Optional[torch.Tensor]) -> torch.Tensor:
    """Remember that ``amax`` (device scalar) bounds ``max |t|`` -- written by the kernel that produced ``t`` (GEMM /
    gather / LayerNorm-backward epilogues) -- so that a consumer needing the fp16 operand scale of ``t`` does not pass
    over it again.  The tag dies with any in-place modification of ``t`` (version check in ``known_amax``)."""
    if amax is not None:
        t._agnn_amax = (amax, t._version)
    return t


def known_amax(t) -> Optional[torch.Tensor]:
    rec = getattr(t, "_agnn_amax", None) if isinstance(t, torch.Tensor) else None
    if rec is not None and rec[1] == t._version:
        return rec[0]
    return None


trace_amax = None        # a dict when bench.py wants the amax passes attributed
_const_scalars = {}


def const_amax(device, value: float) -> torch.Tensor:
    """A device scalar holding a known bound (e.g. 1.0 for GRU states, |h| <= 1)."""
    device = torch.device(device)
    key = (torch.cuda.current_device() if device.index is None else device.index, float(value))
    t = _const_scalars.get(key)
    if t is None:
        t = _const_scalars[key] = torch.full((1,), float(value), dtype=torch.float32, device=device)
    return t


def amax_of(t: torch.Tensor) -> torch.Tensor:
    """The tagged amax of ``t`` if its producer left one, else one agnn_amax pass."""
    am = known_amax(t)
    if am is None:
        stats["amax_passes"] = stats.get("amax_passes", 0) + 1
        if trace_amax is not None:                 # bench.py: which tensors still need a pass, by shape and caller
            import sys
            f = sys._getframe(1)
            site = f"{f.f_code.co_name}<{f.f_back.f_code.co_name}" if f.f_back else f.f_code.co_name
            key = f"{tuple(t.shape)} {site}"
            trace_amax[key] = trace_amax.get(key, 0) + 1
        am = amax_into(new_amax(t.device), t)
    return am


# ---- counter-based dropout (csrc/common.cuh): device state {seed, step} per device, call-site ids per step
_dropout_states = {}
_dropout_calls = [0]


def dropout_state(device) -> torch.Tensor:
    device = torch.device(device)
    idx = torch.cuda.current_device() if device.index is None else device.index
    st = _dropout_states.get(idx)
    if st is None:
        seed = torch.initial_seed() & 0x7FFFFFFFFFFFFFFF
        st = _dropout_states[idx] = torch.tensor([seed, 0], dtype=torch.int64, device=torch.device("cuda", idx))
    return st


def dropout_site() -> int:
    """A fresh call-site id for a fused dropout (ids restart with every ``begin_step``, so a captured step and its
    replays use the same ids while the device-side step counter changes the masks)."""
    _dropout_calls[0] += 1
    return _dropout_calls[0]


def f16_ok(x: torch.Tensor) -> bool:
    return (x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1 and x.shape[1] % 8 == 0
            and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0 and x.shape[0] > 0)


def split_f16(x: torch.Tensor, amax: Optional[torch.Tensor] = None, dropout=None, shift=None) -> SplitH:
    """hi = fp16(s x), lo = fp16(s x - hi); ``amax`` (device scalar >= max |x| / 4) is the producer's tag or measured
    when not given.  ``dropout = (p, site)``: the pair of the dropped-out matrix (counter-based mask, 1 / (1 - p)
    scaling; p < 0.75 keeps the result inside the scale's 4x headroom).  ``shift = (seq_len, s)``: the pair of the matrix
    whose row r is row r - s of the same length-``seq_len`` sequence, zeros elsewhere (agnn_split_f16_shifted)."""
    _need_cuda(x)
    if not f16_ok(x):
        raise ValueError("split_f16 needs an fp32 matrix with unit column stride, 16-byte aligned rows and a column "
                         "count that is a multiple of 8")
    if amax is None:
        amax = amax_of(x)
    rows, cols = x.shape
    buf = torch.empty((2, rows, cols), dtype=torch.float16, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    p, site = dropout if dropout is not None else (0.0, 0)
    if p >= 0.75:
        raise ValueError("fused dropout supports p < 0.75")
    seq_len, sh = shift if shift is not None else (1, 0)
    _lib.check(_lib.lib().agnn_split_f16_shifted(x.data_ptr(), rows, cols, x.stride(0), amax.data_ptr(),
                                                 buf[0].data_ptr(), buf[1].data_ptr(), cols, float(p),
                                                 dropout_state(x.device).data_ptr() if p > 0 else None, int(site),
                                                 int(seq_len), int(sh), stream), "agnn_split_f16")
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
    _dropout_calls[0] = 0
    for st in _dropout_states.values():        # new masks for the new step (inside a captured step: on every replay)
        with torch.cuda.device(st.device):
            _lib.check(_lib.lib().agnn_dropout_advance(st.data_ptr(), torch.cuda.current_stream(st.device).cuda_stream),
                       "agnn_dropout_advance")
        _lib.count_launches(1)


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


def presplit_f16(tensors) -> None:
    """The fp16 operand pairs of several small fp32 matrices (the weights of a grouped launch) in ONE launch
    (agnn_split_f16_multi), left in the per-step split cache where ``_as_operand`` finds them."""
    todo, seen = [], set()
    for x in tensors:
        if not (isinstance(x, torch.Tensor) and x.is_cuda and f16_ok(x) and x.numel() <= _SPLIT_CACHE_MAX_ELEMS):
            continue
        key = ("f16", x.data_ptr(), tuple(x.shape), x.stride(0), x._version)
        if key in _split_cache or key in seen:
            continue
        seen.add(key)
        todo.append((key, x))
    if not todo:
        return
    lib = _lib.lib()
    for lo in range(0, len(todo), _lib.SPLIT_MULTI_MAX):
        chunk = todo[lo:lo + _lib.SPLIT_MULTI_MAX]
        dev = chunk[0][1].device
        arr = (_lib.SplitItem * len(chunk))()
        made = []
        for q, (key, x) in zip(arr, chunk):
            buf = torch.empty((2, x.shape[0], x.shape[1]), dtype=torch.float16, device=dev)
            am = new_amax(dev)
            q.x, q.rows, q.cols, q.ld_x = x.data_ptr(), x.shape[0], x.shape[1], x.stride(0)
            q.hi, q.lo, q.ld_out, q.amax = buf[0].data_ptr(), buf[1].data_ptr(), x.shape[1], am.data_ptr()
            made.append((key, x, SplitH(buf[0], buf[1], am)))
        _lib.check(lib.agnn_split_f16_multi(len(chunk), arr, torch.cuda.current_stream(dev).cuda_stream),
                   "agnn_split_f16_multi")
        _lib.count_launches(1)
        for key, x, sp in made:
            _remember(key, (x, sp))


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


SPLITK_IN_KERNEL = os.environ.get("AGNN_SPLITK", "kernel") == "tickets"
# the large single products on CTA pairs (agnn_gemm_pair, csrc/gemm2.cu)
GEMM_PAIR = os.environ.get("AGNN_GEMM_PAIR", "0") not in ("", "0")
GEMM_PAIR_MIN_FLOPS = 2 * 16384 * 256 * 256
_ticket_pools = {}       # (device, stream) -> zeroed int32 counters for the in-kernel split-K reduction
_TICKETS = 16384


def _tickets(dev) -> torch.Tensor:
    """Per-stream ticket counters (agnn_gemm_grouped): zero before every launch and zero again after it, so launches
    that are ordered on one stream share one array."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    t = _ticket_pools.get(key)
    if t is None:
        t = _ticket_pools[key] = torch.zeros(_TICKETS, dtype=torch.int32, device=dev)
    return t


class _Problem:
    """One contraction of a (possibly grouped) launch, operands already in the form agnn_gemm reads."""
    __slots__ = ("oa", "ob", "m", "n", "k", "bias", "flags", "out", "split_k", "ws", "ws_bytes", "amax_out")


def _resolve(a: Operand, b: Operand, m: int, n: int, k: int, bias, flags: int, out, split_k, amax_out):
    """-> (precision, _Problem) or None if the operands cannot be read as they are."""
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
    p = _Problem()
    p.oa, p.ob, p.m, p.n, p.k, p.flags, p.out, p.amax_out = oa, ob, m, n, k, flags, out, amax_out
    p.bias = bias.float().contiguous() if bias is not None else None
    p.split_k, p.ws, p.ws_bytes = split_k, None, 0      # chosen per launch group (_launch)
    return prec, p


def _launch(prec: int, a_layout: int, b_layout: int, problems) -> None:
    """One agnn_gemm_grouped launch per AGNN_GEMM_MAX_GROUP problems of one precision / layout combination."""
    lib = _lib.lib()
    ptr = lambda t: t.data_ptr() if t is not None else None
    if GEMM_PAIR and a_layout == _lib.K_MAJOR:
        rest = []
        for p in problems:
            if (p.m > 0 and p.n > 0 and p.out.dtype == torch.float32 and (p.out.stride(0) * 4) % 16 == 0
                    and p.out.data_ptr() % 16 == 0 and 2 * p.m * p.n * p.k >= GEMM_PAIR_MIN_FLOPS
                    and lib.agnn_gemm_pair_supported(prec, a_layout, p.m, p.n, p.k, p.flags)):
                dev = p.oa[0].device
                stream = torch.cuda.current_stream(dev).cuda_stream

                def run(p=p):
                    _lib.check(lib.agnn_gemm_pair(b_layout, p.m, p.n, p.k, ptr(p.oa[0]), ptr(p.oa[1]), p.oa[0].stride(0),
                                                  ptr(p.oa[3]), ptr(p.ob[0]), ptr(p.ob[1]), p.ob[0].stride(0),
                                                  ptr(p.ob[3]), p.out.data_ptr(), p.out.stride(0), ptr(p.bias), p.flags,
                                                  ptr(p.amax_out), stream), "agnn_gemm_pair")

                if timer is not None:
                    timer.launch("gemm", 2 * p.m * p.n * p.k, dev, run, tag=(a_layout, b_layout, p.m, p.n, p.k, 1, -2))
                else:
                    run()
                _lib.count_launches(1)
                stats["gemm_launches"] = stats.get("gemm_launches", 0) + 1
                stats["gemm_problems"] = stats.get("gemm_problems", 0) + 1
                stats["gemm_pair_launches"] = stats.get("gemm_pair_launches", 0) + 1
            else:
                rest.append(p)
        problems = rest
    for lo in range(0, len(problems), _lib.GEMM_MAX_GROUP):
        chunk = [p for p in problems[lo:lo + _lib.GEMM_MAX_GROUP] if p.m > 0 and p.n > 0]
        if not chunk:
            continue
        dev = chunk[0].oa[0].device
        arr = (_lib.GemmProblem * len(chunk))()
        need = 0
        import ctypes as C
        n = len(chunk)
        ms, ns, ks = ((C.c_int64 * n)(*[getattr(p, f) for p in chunk]) for f in ("m", "n", "k"))
        splits = (C.c_int32 * n)()
        _lib.check(lib.agnn_gemm_group_split_k(prec, n, ms, ns, ks, splits), "agnn_gemm_group_split_k")
        for p, sk in zip(chunk, splits):
            if p.split_k is None:
                p.split_k = int(sk)
            p.ws_bytes = lib.agnn_gemm_workspace(prec, p.m, p.n, p.k, p.split_k)
            p.ws = torch.empty(p.ws_bytes, dtype=torch.uint8, device=dev) if p.ws_bytes else None
        for q, p in zip(arr, chunk):
            q.M, q.N, q.K = p.m, p.n, p.k
            q.a_hi, q.a_lo, q.lda, q.amax_a = ptr(p.oa[0]), ptr(p.oa[1]), p.oa[0].stride(0), ptr(p.oa[3])
            q.b_hi, q.b_lo, q.ldb, q.amax_b = ptr(p.ob[0]), ptr(p.ob[1]), p.ob[0].stride(0), ptr(p.ob[3])
            q.c, q.ldc, q.bias, q.flags, q.split_k = p.out.data_ptr(), p.out.stride(0), ptr(p.bias), p.flags, p.split_k
            q.workspace, q.workspace_bytes, q.amax_out = ptr(p.ws), p.ws_bytes, ptr(p.amax_out)
            need += lib.agnn_gemm_tickets(p.m, p.n, p.split_k)
        # split-K partials: added in split order either by ONE grouped reduce kernel behind the launch (default) or, with
        # ticket counters, inside the launch by the CTA that stores a tile's last partial (AGNN_SPLITK=tickets)
        tickets = _tickets(dev) if (SPLITK_IN_KERNEL and 0 < need <= _TICKETS) else None
        stream = torch.cuda.current_stream(dev).cuda_stream

        def run():
            _lib.check(lib.agnn_gemm_grouped(prec, a_layout, b_layout, len(chunk), arr, ptr(tickets),
                                             tickets.numel() if tickets is not None else 0, stream), "agnn_gemm_grouped")

        if timer is not None:                   # bench.py: per-launch CUDA events, algorithmic flops = 2 M N K
            flops = sum(2 * p.m * p.n * p.k for p in chunk)
            big = max(chunk, key=lambda p: p.m * p.n * p.k)
            timer.launch("gemm", flops, dev, run, tag=(a_layout, b_layout, big.m, big.n, big.k, big.split_k, len(chunk)))
        else:
            run()
        _lib.count_launches(1 if tickets is not None or need == 0 else 2)
        stats["gemm_launches"] = stats.get("gemm_launches", 0) + 1
        stats["gemm_problems"] = stats.get("gemm_problems", 0) + len(chunk)


def _gemm(a: Operand, a_layout: int, b: Operand, b_layout: int, m: int, n: int, k: int, bias, flags: int,
          out: Optional[torch.Tensor], split_k: Optional[int] = None, amax_out=None) -> Optional[torch.Tensor]:
    """One contraction on agnn_gemm_grouped; returns None if the operands are not eligible as they are
    (``_repacked`` then copies them into padded buffers)."""
    r = _resolve(a, b, m, n, k, bias, flags, out, split_k, amax_out)
    if r is None:
        return None
    _launch(r[0], a_layout, b_layout, [r[1]])
    return r[1].out


def _group(a_layout: int, b_layout: int, specs):
    """``specs``: list of dicts(a, b, m, n, k, bias, flags, out, amax_out).  Independent contractions of one layout
    combination in as few launches as their precisions allow (one per precision, AGNN_GEMM_MAX_GROUP problems each);
    problems whose operands need a repack run alone.  Returns the outputs in order."""
    outs = [None] * len(specs)
    by_prec = {}
    # plain fp32 operands that will meet an fp16 pair (the weights): all their pairs in one launch
    presplit_f16([sp["b"] if isinstance(sp["a"], SplitH) else sp["a"] for sp in specs
                  if isinstance(sp["a"], SplitH) != isinstance(sp["b"], SplitH) and sp["m"] > 0 and sp["n"] > 0])
    for i, sp in enumerate(specs):
        if sp["m"] == 0 or sp["n"] == 0 or sp["k"] == 0:       # an empty member (a node type without nodes in this batch)
            if sp.get("out") is not None:
                outs[i] = sp["out"] if sp.get("flags", 0) & _lib.GEMM_ACCUMULATE else sp["out"].zero_()
            else:
                a0 = sp["a"].hi if isinstance(sp["a"], (Split, SplitH)) else sp["a"]
                dt = torch.bfloat16 if a0.dtype == torch.bfloat16 else torch.float32
                outs[i] = torch.zeros((sp["m"], sp["n"]), dtype=dt, device=a0.device)
            continue
        r = _resolve(sp["a"], sp["b"], sp["m"], sp["n"], sp["k"], sp.get("bias"), sp.get("flags", 0), sp.get("out"),
                     None, sp.get("amax_out"))
        if r is None:
            outs[i] = _repacked(sp["a"], a_layout, sp["b"], b_layout, sp["m"], sp["n"], sp["k"], sp.get("bias"),
                                sp.get("flags", 0), sp.get("out"))
            if sp.get("amax_out") is not None:
                amax_into(sp["amax_out"], outs[i])
            continue
        by_prec.setdefault(r[0], []).append((i, r[1]))
    for prec, items in by_prec.items():
        _launch(prec, a_layout, b_layout, [p for _, p in items])
        for i, p in items:
            outs[i] = p.out
    return outs


def linear_group(xs, weights, biases=None, relu: bool = False, amax_outs=None, outs=None):
    """``[x_i @ w_i.T + b_i]`` for independent projections (per node type, per task head, per GRU direction) in one
    grouped launch."""
    biases = biases if biases is not None else [None] * len(xs)
    amax_outs = amax_outs if amax_outs is not None else [None] * len(xs)
    outs = outs if outs is not None else [None] * len(xs)
    flags = _lib.GEMM_RELU if relu else 0
    for x in xs:
        _need_cuda(x if isinstance(x, torch.Tensor) else x.hi)
    return _group(_lib.K_MAJOR, _lib.K_MAJOR,
                  [dict(a=x, b=w, m=x.shape[0], n=w.shape[0], k=x.shape[1], bias=bb, flags=flags, out=o, amax_out=am)
                   for x, w, bb, am, o in zip(xs, weights, biases, amax_outs, outs)])


def mm_group(as_, bs, outs=None, accumulate: bool = False, amax_outs=None):
    """``[a_i @ b_i]`` with ``b_i`` stored [K, N] (grad-input products of independent projections)."""
    outs = outs if outs is not None else [None] * len(as_)
    amax_outs = amax_outs if amax_outs is not None else [None] * len(as_)
    flags = _lib.GEMM_ACCUMULATE if accumulate else 0
    return _group(_lib.K_MAJOR, _lib.MN_MAJOR,
                  [dict(a=a, b=b, m=a.shape[0], n=b.shape[1], k=a.shape[1], flags=flags, out=o, amax_out=am)
                   for a, b, o, am in zip(as_, bs, outs, amax_outs)])


def mm_tn_group(as_, bs):
    """``[a_i.T @ b_i]`` (grad-weight products of independent projections; split-K inside the launch)."""
    return _group(_lib.MN_MAJOR, _lib.MN_MAJOR,
                  [dict(a=a, b=b, m=a.shape[1], n=b.shape[1], k=a.shape[0]) for a, b in zip(as_, bs)])


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


def linear(x: Operand, weight: Operand, bias=None, relu: bool = False, amax_out=None, out=None):
    """``x @ weight.T + bias`` (optionally ReLU'd), ``weight`` [out, in].  ``amax_out``: device scalar that receives
    ``max(amax_out, max |y|)`` from the GEMM epilogue; ``out``: write into this view (e.g. a column slice)."""
    return linear_group([x], [weight], [bias], relu, [amax_out], [out])[0]


def mm(a: Operand, b: Operand, out=None, accumulate: bool = False, amax_out=None):
    """``a @ b`` with ``b`` stored [K, N]; with ``out`` and ``accumulate`` adds into ``out`` in place."""
    _need_cuda(a if isinstance(a, torch.Tensor) else a.hi)
    return mm_group([a], [b], [out], accumulate, [amax_out])[0]


def mm_tn(a: Operand, b: Operand):
    """``a.T @ b`` (weight gradients): ``a`` [R, M], ``b`` [R, N], reduction over the R rows (split-K)."""
    _need_cuda(a if isinstance(a, torch.Tensor) else a.hi)
    return mm_tn_group([a], [b])[0]


def relu_backward(grad, out):
    """Gradient through ``out = relu(.)`` given the saved output."""
    return torch.where(out > 0, grad, torch.zeros((), dtype=grad.dtype, device=grad.device))
