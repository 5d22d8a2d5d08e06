"""Data-parallel training step around the message-passing modules.

The reference trains through Lightning: DDP's bucketed gradient allreduce when
several devices are listed (analysisgnn/train/train_analysisgnn.py:138-145,
246-248), ``gradient_clip_val=1.0`` (:254) and ``torch.optim.AdamW``
(analysisgnn/models/analysis.py:1380).  Here the same arithmetic is:

* one flat fp32 gradient arena (``p.grad`` are views into it), zeroed with one memset;
  parameters that get no gradient in a step (task-dependent heads, the reason for the
  reference's ``find_unused_parameters``) simply contribute zeros;
* one in-place ``all_reduce(sum)`` over the arena (NCCL over NVLink / NVSwitch);
* agnn_sumsq_partials + agnn_adamw_clip_step: 1/world averaging, global-norm clipping
  and AdamW fused in two launches (include/agnn.h).

One process per GPU; batches are sharded across ranks by subgraph (no edge crosses
subgraphs, so message passing needs no exchange).
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import _lib


def shard_indices(n_items: int, rank: int, world: int):
    """Subgraphs ``{g : g mod world == rank}`` of a global batch."""
    return list(range(rank, n_items, world))


class GradArena:
    """Flat fp32 buffers for gradients and AdamW moments + the device chunk table
    that maps arena ranges back to the modules' own parameter tensors."""

    def __init__(self, params: Iterable[torch.nn.Parameter], chunk_elems: int):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                raise ValueError("GradArena needs contiguous fp32 parameters on one device")
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4           # 16-byte aligned slices
        self.numel = off
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.grad)
        self.exp_avg_sq = torch.zeros_like(self.grad)
        for p, o in zip(self.params, self.offsets):
            p.grad = self.grad[o:o + p.numel()].view_as(p)
        chunks = []
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            for lo in range(0, n, chunk_elems):
                chunks.append((p.data_ptr(), lo, o + lo, min(chunk_elems, n - lo),
                               int((p.data_ptr() + 4 * lo) % 16 == 0)))
        table = (_lib.ParamChunk * len(chunks))()
        for i, c in enumerate(chunks):
            table[i].param, table[i].param_off, table[i].arena_off, table[i].count, table[i].param_aligned = c
        raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8)
        self.n_chunks = len(chunks)
        self.table = raw.to(dev) if dev.type == "cuda" else raw
        self._ptrs = [p.data_ptr() for p in self.params]

    def zero_(self):
        self.grad.zero_()

    def views(self):
        return [self.grad[o:o + p.numel()].view_as(p) for p, o in zip(self.params, self.offsets)]

    def detach_views(self):
        """``p.grad = None``: autograd then hands over each gradient tensor as produced (no in-place add per
        parameter); ``collect`` brings them into the arena."""
        for p in self.params:
            p.grad = None

    def collect(self):
        """Copy the parameters' gradient tensors into the arena with one multi-tensor copy, zero the slices of
        parameters that received none (task-dependent heads), and point every ``p.grad`` at its slice again."""
        views = self.views()
        src, dst, unused = [], [], []
        for p, v in zip(self.params, views):
            g = p.grad
            if g is None:
                unused.append(v)
            elif g.data_ptr() != v.data_ptr():
                src.append(g if g.dtype == v.dtype else g.to(v.dtype))
                dst.append(v)
        if unused:
            torch._foreach_zero_(unused)                  # one multi-tensor launch, not one fill per parameter
        if dst:
            torch._foreach_copy_(dst, src)
        for p, v in zip(self.params, views):
            p.grad = v

    def check_views(self):
        """Autograd accumulates in place into an existing ``.grad``; anything that replaced
        it (``zero_grad(set_to_none=True)``) would silently detach the arena."""
        for p, o, ptr in zip(self.params, self.offsets, self._ptrs):
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                raise RuntimeError("a parameter's .grad no longer aliases the gradient arena "
                                   "(use trainer.zero_grad(), not zero_grad(set_to_none=True))")
            if p.data_ptr() != ptr:
                raise RuntimeError("a parameter tensor was reallocated after the arena was built")


class DataParallelTrainer:
    """``zero_grad() -> loss.backward() -> step()``; ``step`` = allreduce + clip + AdamW."""

    def __init__(self, model: torch.nn.Module, lr=5e-3, weight_decay=5e-3, betas=(0.9, 0.999), eps=1e-8,
                 max_norm: float = 1.0, process_group=None, world_size: Optional[int] = None,
                 collect_grads: bool = False):
        import torch.distributed as dist
        self.model = model
        self.lr, self.weight_decay, self.betas, self.eps, self.max_norm = lr, weight_decay, betas, eps, max_norm
        self.group = process_group
        if world_size is None:
            world_size = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.world_size = world_size
        self.step_count = 0
        # collect_grads: ``zero_grad`` detaches the views and ``collect`` (called by ``allreduce`` / ``step``, or
        # explicitly inside a captured step) copies the finished gradients in -- ~180 in-place adds per step less
        self.collect_grads = collect_grads
        self._collected = True
        first = next(p for p in model.parameters() if p.requires_grad)
        if first.is_cuda:
            lib = _lib.lib()
            self.arena = GradArena(model.parameters(), lib.agnn_optim_chunk_elems())
            self.n_partials = lib.agnn_sumsq_blocks(self.arena.numel)
            self.partials = torch.empty(self.n_partials, dtype=torch.float32, device=first.device)
        else:   # host-side logic only (gloo tests): arena + allreduce, no optimizer kernels
            self.arena = GradArena(model.parameters(), 4096)
            self.n_partials, self.partials = 0, None
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=first.device)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=first.device)   # device-side step number
        # the learning rate lives on the device: the optimizer launch reads it, so the host's scheduler
        # (``set_lr``; the reference's warm-up + cosine schedule, analysis.py:1380-1400) keeps working when the launch
        # is replayed from a CUDA graph
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=first.device)
        if self.world_size > 1:
            self.broadcast_state()

    def set_lr(self, lr: float) -> None:
        """Write a new learning rate (host scheduler) into the device scalar the optimizer kernel reads."""
        self.lr = float(lr)
        self.lr_dev.fill_(self.lr)

    def broadcast_state(self, src: int = 0) -> None:
        """What DDP does when it wraps a module (train_analysisgnn.py:138-145): every rank starts from rank ``src``'s
        parameters AND buffers (BatchNorm running statistics of MetricalConvLayer, gnn.py:498-531).  Buffers are
        rank-local afterwards, as under Lightning's DDP without SyncBatchNorm (``broadcast_buffers`` re-syncs them
        from rank 0 before every forward there; call this again to do the same)."""
        import torch.distributed as dist
        with torch.no_grad():
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t.data, src=src, group=self.group)

    def zero_grad(self):
        if self.collect_grads:
            self.arena.detach_views()
            self._collected = False
        else:
            self.arena.zero_()

    def collect(self):
        if self.collect_grads and not self._collected:
            self.arena.collect()
            self._collected = True

    def allreduce(self):
        self.collect()
        if self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(self.arena.grad, op=dist.ReduceOp.SUM, group=self.group)

    def step(self):
        a = self.arena
        if not a.grad.is_cuda:
            raise _lib.AgnnError("DataParallelTrainer.step needs CUDA parameters: there is no CPU optimizer path")
        self.collect()
        a.check_views()
        self.allreduce()
        self.step_count += 1                    # host mirror; the kernels use the device counter
        lib = _lib.lib()
        stream = torch.cuda.current_stream(a.grad.device).cuda_stream
        _lib.check(lib.agnn_sumsq_partials(a.grad.data_ptr(), a.numel, self.partials.data_ptr(),
                                           self.step_dev.data_ptr(), stream), "agnn_sumsq_partials")
        _lib.check(lib.agnn_adamw_clip_step(
            a.table.data_ptr(), a.n_chunks, a.grad.data_ptr(), a.exp_avg.data_ptr(), a.exp_avg_sq.data_ptr(),
            self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, 0, self.step_dev.data_ptr(),
            self.lr_dev.data_ptr(), 1.0 / self.world_size, self.max_norm if self.max_norm else 0.0, self.partials.data_ptr(),
            self.n_partials, self.grad_norm.data_ptr(), stream), "agnn_adamw_clip_step")
        _lib.count_launches(2)
        from . import linalg
        linalg.begin_step()                     # the weights changed: cached TF32 splits are stale


class GraphedStep:
    """The device work of a training step (CSR build, forward, backward; optionally the optimizer)
    captured once in a CUDA graph and replayed: the step is ~2 500 short launches, and launching them
    one by one from Python costs more than the GPU needs to run them.  Everything libagnn enqueues is
    capturable (it never allocates or synchronises; the optimizer's step number lives on the device).
    With several ranks keep ``trainer.step()`` (the NCCL allreduce) outside the captured function.

    ``fn(inputs) -> loss`` must read its data from the tensors in ``inputs`` (static buffers: write the
    next batch into them with ``copy_`` -- shapes are fixed by the capture, so batches are padded to
    the captured capacity with relation id -1 edges and ``ignore_index`` labels) and must not read
    device data on the host.  Host-side layouts that need such a read (sequence lengths) are taken
    from the warm-up steps, which run eagerly before the capture.

    Per-batch device structures are part of the captured work: before every warm-up call and before the capture this
    class itself drops the CSR cache (``graph.clear_cache``) and the per-step weight splits / amax scalars
    (``linalg.begin_step``), so the ``agnn_csr_build`` launches and the scalar resets are IN the graph and every replay
    rebuilds them from whatever the static inputs hold then -- ``fn`` does not have to remember to.  If the optimizer
    launch is captured too, the learning rate it uses is the trainer's device scalar (``DataParallelTrainer.set_lr``),
    not a constant baked into the graph."""

    def __init__(self, fn, inputs, warmup: int = 3):
        from . import graph as _graph, linalg as _linalg

        def step(x):
            _graph.clear_cache()
            _linalg.begin_step()
            return fn(x)

        self.fn, self.inputs = step, inputs
        dev = torch.cuda.current_device()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step(inputs)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        before = _lib.launches()
        self.graph = torch.cuda.CUDAGraph()
        _graph.freeze_host_layouts(True)
        try:
            with torch.cuda.graph(self.graph):
                self.loss = step(inputs)
        finally:
            _graph.freeze_host_layouts(False)
        self.launches_per_replay = _lib.launches() - before

    def __call__(self):
        self.graph.replay()
        _lib.count_launches(self.launches_per_replay)
        return self.loss
