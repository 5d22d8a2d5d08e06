#!/usr/bin/env python
"""Benchmark of the message-passing hot path (BASELINE.json metric: score-graph
nodes/s fwd+bwd, HybridGNN 3 layers / hidden 256).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype fp32|bf16]

Workload = BASELINE.json configs[1]: HybridGNN 3L/256 with beat + measure nodes and
the three multi-task heads (cadence / localkey / romanNumeral) on a batch of 100
synthetic 500-note score subgraphs per GPU; one step = CSR build + forward +
backward + gradient allreduce (N > 1) + clip + AdamW, train mode (dropout 0.3).

Prints ONE JSON line (rank 0).  `value` times the step with inputs resident in HBM;
`e2e` times the same step from pinned host buffers (host->device copies of the batch
and a device->host read of the loss inside the timed region; the copy of step i+1's
batch runs on a copy stream while step i computes, into a second staging set).  `roofline` is the
aggregation kernel (agnn_gather_reduce) measured with CUDA events inside the timed
region; `cpu_baseline` is the CPU oracle on a bounded sample of the same workload.
`--impl reference` times the reference arm: this repo's CPU restatement of the
reference's encoder (oracle/pyg.py; the third-party graphmuse / PyG stack is not
installable here) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TASKS = {"cadence": 4, "localkey": 50, "romanNumeral": 185}   # analysisgnn/train/train_analysisgnn.py:22-45
CFG = dict(graphs=100, notes=500, voices=4, in_features=25, hidden=256, out=128, layers=3, dropout=0.3,
           lr=5e-3, weight_decay=5e-3, max_norm=1.0)
METRIC = "score-graph nodes/sec fwd+bwd (HybridGNN 3L/256)"
N_HOST_BATCHES = 4        # distinct host batches the end-to-end arm rotates through (a different one every step)


def peaks():
    """(HBM GB/s, bf16 TFLOP/s burst, source)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def traffic_of(kernel):
    """{"traffic": DRAM bytes per launch or None, "traffic_source": ...} from profiles/traffic.json."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as fh:
            t = json.load(fh).get(kernel)
    except (OSError, ValueError):
        t = None
    if not t:
        return {"traffic": None, "traffic_source": "no ncu --set full capture of this kernel committed for this round"}
    return {"traffic": t["dram_bytes"], "traffic_source": t["source"],
            "traffic_algorithmic_bytes_same_launch": t.get("algorithmic_bytes")}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_batch(seed, graphs):
    from analysisgnn_b200 import synth
    return synth.hetero_batch(graphs, CFG["notes"], seed, voices=CFG["voices"], in_features=CFG["in_features"],
                              task_dict=TASKS)


def batch_tensors(b):
    """Flat name -> tensor map of everything a step consumes (what crosses host->device)."""
    t = {"pitch_spelling": b["pitch_spelling"], "key_signature": b["key_signature"]}
    for k, v in b["x_dict"].items():
        t[f"x.{k}"] = v
    for k, v in b["edge_index_dict"].items():
        t["ei." + "__".join(k)] = v
    for k, v in b["batch_dict"].items():
        t[f"batch.{k}"] = v
    for k, v in b["labels"].items():
        t[f"label.{k}"] = v
    return t


def unflatten(t, b):
    return dict(pitch_spelling=t["pitch_spelling"], key_signature=t["key_signature"],
                x_dict={k: t[f"x.{k}"] for k in b["x_dict"]},
                edge_index_dict={k: t["ei." + "__".join(k)] for k in b["edge_index_dict"]},
                batch_dict={k: t[f"batch.{k}"] for k in b["batch_dict"]},
                labels={k: t[f"label.{k}"] for k in b["labels"]}, batch_size=b["batch_size"])


# --------------------------------------------------------------------------- CPU arm

def cpu_step_fn(graphs, seed=0):
    """The reference arm / cpu_baseline: oracle restatement on the host cores, same step."""
    import torch
    from oracle import pyg as opyg
    b = make_batch(seed, graphs)
    torch.manual_seed(0)
    model = opyg.AnalysisEncoderShell(b["metadata"], CFG["in_features"], CFG["hidden"], CFG["out"], TASKS,
                                      CFG["layers"], dropout=CFG["dropout"])
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=CFG["lr"], weight_decay=CFG["weight_decay"])

    def step():
        opt.zero_grad(set_to_none=True)
        logits = model(b["pitch_spelling"], b["key_signature"], b["x_dict"], b["edge_index_dict"], b["batch_dict"],
                       b["batch_size"], None, None)
        loss = opyg.multitask_ce(logits, b["labels"])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), CFG["max_norm"])
        opt.step()
        return float(loss.detach())

    return step, b["batch_size"]


def cpu_arm(steps, warmup, budget_s, max_graphs):
    """Times the CPU oracle on a bounded sample: the number of subgraphs per step is chosen
    so that warmup + steps fit in ``budget_s`` seconds."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    probe_graphs = min(4, max_graphs)
    step, nodes = cpu_step_fn(probe_graphs)
    step()
    t0 = time.perf_counter()
    step()
    per_graph = (time.perf_counter() - t0) / probe_graphs
    graphs = int(max(1, min(max_graphs, budget_s / max(per_graph * (steps + warmup), 1e-9))))
    if graphs != probe_graphs:
        step, nodes = cpu_step_fn(graphs)
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": nodes * steps / total, "unit": "nodes/s", "cores": cores, "kind": "port",
            "sample": f"{graphs} of {CFG['graphs']} subgraphs x {CFG['notes']} notes per step, {steps} timed steps "
                      f"(+{warmup} warm-up), oracle/pyg.py AnalysisEncoderShell fwd+bwd+clip+AdamW, fp32, "
                      f"torch {cores} threads",
            "ms_per_step": 1e3 * total / steps, "nodes_per_step": nodes}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    r = cpu_arm(steps, warmup, budget_s=150.0, max_graphs=CFG["graphs"])
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "nodes/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, cpu=True),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, cpu=False):
    return {"workload": "BASELINE configs[1]: HybridGNN 3L/256 + beat/measure nodes + cadence/localkey/romanNumeral "
                        "heads, 100 synthetic 500-note subgraphs per GPU (4 voices, 25 note features), "
                        "step = CSR build + fwd + bwd + clip(1.0) + AdamW, train mode dropout 0.3",
            "subgraphs_per_gpu": CFG["graphs"], "notes_per_subgraph": CFG["notes"], "hidden": CFG["hidden"],
            "layers": CFG["layers"], "parallelism": f"dp{args.gpus}",
            "parity_gemm_operands": "n/a (CPU)" if cpu else getattr(args, "operands", "tf32"),
            "degree_bound_hint": None if cpu else CFG["notes"],
            "l2": "n/a (CPU)" if cpu else "L2 flushed between timed steps (256 MiB write); per-step activations "
                                           "(~2 GB) also exceed the 126 MB L2"}


# --------------------------------------------------------------------------- GPU arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    from analysisgnn_b200 import _lib, graph, ops
    from analysisgnn_b200 import nn as ann
    from analysisgnn_b200.train import DataParallelTrainer, GraphedStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference "
                         "for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if args.dtype != "fp32":
        raise SystemExit("bench.py: the training step runs in fp32 (the reference's precision); the bf16 mode is "
                         "covered by the parity tests only")
    dtype = torch.float32

    # N_HOST_BATCHES different synthetic batches (different scores, features, labels, edge counts), padded to common
    # shapes so that one captured graph serves them all: edge lists with -1 entries (dropped by agnn_csr_build), beat /
    # measure node arrays with isolated zero-feature nodes.  The end-to-end arm copies a DIFFERENT one every step.
    raw = [make_batch(seed=1000 + 17 * i + rank, graphs=CFG["graphs"]) for i in range(N_HOST_BATCHES)]
    b = raw[0]
    flats = [batch_tensors(x) for x in raw]
    caps = {k: max(f[k].shape[-1] if k.startswith("ei.") else f[k].shape[0] for f in flats) for k in flats[0]}

    def padded(f):
        out = {}
        for k, v in f.items():
            if k.startswith("ei."):
                out[k] = torch.nn.functional.pad(v, (0, caps[k] - v.shape[1]), value=-1)
            elif k.startswith("batch.") and v.shape[0] < caps[k]:
                out[k] = torch.cat((v, v[-1:].expand(caps[k] - v.shape[0])))
            elif v.shape[0] < caps[k]:
                out[k] = torch.cat((v, v.new_zeros((caps[k] - v.shape[0],) + tuple(v.shape[1:]))))
            else:
                out[k] = v
        return out

    hosts = [{k: v.contiguous().pin_memory() for k, v in padded(f).items()} for f in flats]
    host = hosts[0]
    resident = {k: v.to(dev) for k, v in host.items()}
    staging = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    n_nodes = b["batch_size"]
    n_edges = sum(v.shape[1] for k, v in b["edge_index_dict"].items())

    torch.manual_seed(0)
    model = ann.AnalysisEncoder(b["metadata"], CFG["in_features"], CFG["hidden"], CFG["out"], TASKS, CFG["layers"],
                                dropout=CFG["dropout"]).to(dev)
    model.train()
    trainer = DataParallelTrainer(model, lr=CFG["lr"], weight_decay=CFG["weight_decay"], max_norm=CFG["max_norm"],
                                  world_size=world, collect_grads=True)

    from analysisgnn_b200 import linalg as _lin
    _lin.set_parity_operands(args.operands)
    # the measured path must stay on the hand-written kernels: an operand repack or a library kernel raises
    _lib.set_strict(True)
    # every batch of this bench is a disjoint union of CFG["notes"]-note subgraphs: no CSR row (in- / out-neighbours, the
    # notes of a beat or measure, the nodes of a pooled graph) is longer than that, which is below the hub-row threshold
    # -> the CSR build skips the hub-row lists and the aggregations their two hub-row launches (a hint: see graph.py)
    graph.set_degree_bound(CFG["notes"])

    def fwd_bwd(tensors):
        _lin.begin_step()                          # weight splits are per step (a captured step re-splits on replay)
        graph.clear_cache()                        # a new batch every step: the CSR build is part of the step
        d = unflatten(tensors, b)
        trainer.zero_grad()
        logits = model(d["pitch_spelling"], d["key_signature"], d["x_dict"], d["edge_index_dict"],
                       d["batch_dict"], d["batch_size"], None, None)
        loss = ann.multitask_ce(logits, d["labels"])
        loss.backward()
        trainer.collect()                          # gradients -> flat arena (inside the captured region)
        return loss

    def step(tensors):
        loss = fwd_bwd(tensors)
        trainer.step()                             # NCCL allreduce (N > 1) + fused clip + AdamW
        return loss

    not_pinned = [k for h in hosts for k, v in h.items() if v.numel() and not v.is_pinned()]
    if not_pinned or not loss_host.is_pinned():
        raise SystemExit(f"bench.py: host buffers are not pinned: {not_pinned}")

    # End to end: the batch comes from pinned host memory every step.  Two device staging sets: while step i computes
    # from set i % 2, the copy stream brings step i+1's batch into the other set (what a prefetching loader does).
    # Every timed step's host->device copy is issued and completes inside the timed region (the first one exposed,
    # the last step prefetches nothing), and every step reads its loss back into pinned host memory.
    staging2 = [staging, {k: torch.empty_like(v, device=dev) for k, v in host.items()}]
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event(), torch.cuda.Event()]      # set s holds its batch
    consumed = [torch.cuda.Event(), torch.cuda.Event()]    # the step reading set s has finished

    class E2E:
        slot = 0
        remaining = 0          # steps left in the current run (0 = unknown: always prefetch)
        pending = [False, False]
        used = [False, False]
        copies = 0

        @classmethod
        def reset(cls, steps):
            torch.cuda.synchronize()
            cls.slot, cls.remaining, cls.pending, cls.copies = 0, steps, [False, False], 0

        @classmethod
        def fetch(cls, s):
            main = torch.cuda.current_stream(dev)
            if cls.used[s]:
                copy_stream.wait_event(consumed[s])
            else:
                copy_stream.wait_stream(main)
            with torch.cuda.stream(copy_stream):
                for k, v in hosts[cls.copies % N_HOST_BATCHES].items():     # a different batch every step
                    if v.numel():
                        staging2[s][k].copy_(v, non_blocking=True)
                copied[s].record(copy_stream)
            cls.pending[s] = True
            cls.copies += 1

        @classmethod
        def run(cls, compute):
            s = cls.slot
            main = torch.cuda.current_stream(dev)
            if not cls.pending[s]:
                cls.fetch(s)
            main.wait_event(copied[s])
            cls.pending[s] = False
            if cls.remaining != 1:
                cls.fetch(1 - s)                   # overlaps this step's compute
            compute(s)
            consumed[s].record(main)
            cls.used[s] = True
            trainer.step()
            cls.slot = 1 - s
            if cls.remaining > 0:
                cls.remaining -= 1

    def e2e_compute(s):
        loss = fwd_bwd(staging2[s])
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        return loss

    def e2e_step():
        E2E.run(e2e_compute)

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        t0 = time.perf_counter()
        enqueue = 0.0
        for a, z in ev:
            flush.fill_(1.0)                       # L2 flush, outside the event pair
            a.record()
            t1 = time.perf_counter()
            fn()
            enqueue += time.perf_counter() - t1
            z.record()
        barrier()
        wall = time.perf_counter() - t0
        timed.enqueue_ms = 1e3 * enqueue / steps   # host time spent launching one step
        ms = sum(a.elapsed_time(z) for a, z in ev)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(resident)
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()

    use_graph = not args.no_graph
    if use_graph:
        # the step is launch-bound from Python (host enqueue ~ step time in eager mode): capture the
        # CSR build + forward + backward once; the allreduce and the two optimizer launches stay eager
        g_resident = GraphedStep(fwd_bwd, resident, warmup=1)
        g_e2e = [GraphedStep(e2e_compute, 0, warmup=1), GraphedStep(e2e_compute, 1, warmup=1)]

        def run_resident():
            g_resident()
            trainer.step()

        def run_e2e():
            E2E.run(lambda s: g_e2e[s]())
    else:
        run_resident, run_e2e = (lambda: step(resident)), e2e_step
    for _ in range(warm):
        run_resident()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launches()
    torch.cuda.profiler.start()                    # ncu --profile-from-start off captures the timed region only
    ms, wall = timed(run_resident, args.steps)
    torch.cuda.profiler.stop()
    launches = _lib.launches() - launches0
    enqueue_ms = timed.enqueue_ms
    E2E.reset(0)
    for _ in range(2):
        run_e2e()
    E2E.reset(args.steps)
    e2e_ms, _ = timed(run_e2e, args.steps)
    assert E2E.copies == args.steps, (E2E.copies, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    final_loss = float(loss_host.item())
    # per-kernel timing of the aggregation kernels: CUDA events cannot be read inside a graph, so the same
    # K steps run once more eagerly with an event pair around every agnn_gather_reduce launch
    eager_ms, _ = timed(lambda: step(resident), args.steps)
    eager_enqueue_ms = timed.enqueue_ms
    # ... with the sequence branch in line (not on its side stream), so that each launch is timed alone
    from analysisgnn_b200.nn import hetero as _hetero
    _hetero._HybridBase.overlap_sequence_branch = False
    from analysisgnn_b200 import linalg as _linalg
    ops.timer = _linalg.timer = ops.KernelTimer()
    _linalg.trace_amax = {}
    stats0 = dict(_linalg.stats)
    serial_ms, _ = timed(lambda: step(resident), args.steps)
    amax_sites = {k: v / args.steps for k, v in sorted(_linalg.trace_amax.items(), key=lambda kv: -kv[1])}
    _linalg.trace_amax = None
    gemm_stats = {k: (_linalg.stats.get(k, 0) - stats0.get(k, 0)) / args.steps
                  for k in ("gemm_launches", "gemm_problems", "amax_passes", "repacked_gemms")}
    ktimes = ops.timer.summary()
    if os.environ.get("AGNN_DUMP_GEMM") and rank == 0:      # per-shape GEMM time of one step, to stderr
        lay = {0: "K", 1: "MN"}
        rows = sorted(ops.timer.by_tag("gemm").items(), key=lambda kv: -kv[1][1])
        for (la, lb, m_, n_, k_, sk, grp_), (cnt_, ms_) in rows:     # largest member of every (grouped) launch
            print(f"gemm A:{lay.get(la, la)} B:{lay.get(lb, lb)} M={m_:6d} N={n_:5d} K={k_:6d} split={sk:2d} "
                  f"group={grp_:2d}  x{cnt_ / args.steps:5.1f}/step  {ms_ / args.steps * 1e3:8.1f} us/step",
                  file=sys.stderr)
    ops.timer = _linalg.timer = None
    _hetero._HybridBase.overlap_sequence_branch = True

    if rank == 0:
        hbm_peak, bf16_peak, peak_src = peaks()
        zero = {"launches": 0, "bytes": 0, "ms": 0.0, "max_bytes": 0, "max_ms": 0.0}
        g = ktimes.get("gather_reduce", zero)
        mm = ktimes.get("gemm", zero)
        achieved = g["bytes"] / (g["ms"] * 1e-3) / 1e9 if g["ms"] > 0 else 0.0
        big = g["max_bytes"] / (g["max_ms"] * 1e-3) / 1e9 if g["max_ms"] > 0 else 0.0
        # TF32 runs at half the bf16 tensor rate: the measured bf16 cuBLAS burst / 2 is the denominator
        tf32_peak = bf16_peak / 2.0
        f16_ops = args.operands == "f16"
        # f16 operand mode: the big layer GEMMs run on the f16 MMA (the bf16 burst rate), the small ones stay 3xTF32;
        # the denominator is the faster pipe's rate
        mma_peak = bf16_peak if f16_ops else tf32_peak
        mm_tflops = mm["bytes"] / (mm["ms"] * 1e-3) / 1e12 if mm["ms"] > 0 else 0.0
        mm_big = mm["max_bytes"] / (mm["max_ms"] * 1e-3) / 1e12 if mm["max_ms"] > 0 else 0.0
        timed_in = ("eager, single-stream re-run of the same K steps (CUDA events around each launch; events cannot "
                    "be read inside the replayed graph)")
        roofline_gemm = {
            "bound": "tensor", "kernel": "agnn gemm_kernel (tcgen05 " + ("3xF16 message-passing layers + 3xTF32 rest" if f16_ops
                                                                          else "3xTF32") + ", all launches of the step)",
            "achieved": mm_tflops, "peak": mma_peak, "unit": "TFLOP/s", "frac": mm_tflops / mma_peak,
            # what the tensor pipe actually executes in the fp32-parity mode: 3 TF32 MMAs per algorithmic product
            "mma_achieved": mm_tflops * (3 if dtype == torch.float32 else 1),
            "mma_frac": mm_tflops * (3 if dtype == torch.float32 else 1) / (mma_peak if dtype == torch.float32
                                                                              else bf16_peak),
            # dram__bytes_read.sum + dram__bytes_write.sum of the largest launch from the ncu --set full capture of THIS
            # round's kernel (profiles/traffic.json, written by tools/ncu_traffic.py from the committed raw page);
            # null when no capture of the current kernel is committed
            **traffic_of("gemm_f16" if f16_ops else "gemm_tf32"),
            "peak_source": peak_src + (": bf16 burst (f16 MMA rate)" if f16_ops else ": bf16 burst / 2 (TF32 rate)"),
            "launches": mm["launches"],
            "algorithmic_flops_per_step": mm["bytes"] / max(args.steps, 1),
            "kernel_ms_per_step": mm["ms"] / max(args.steps, 1),
            "share_of_step": mm["ms"] / serial_ms if serial_ms else None, "timed_in": timed_in,
            "note": "algorithmic flops = 2*M*N*K per GEMM; the fp32-parity modes spend 3 MMAs per product, so "
                    "frac <= 0.33 by construction (tensor-pipe active cycles are in profiles/)",
            "largest_launch": {"flops": mm["max_bytes"], "ms": mm["max_ms"], "achieved": mm_big,
                               "frac": mm_big / mma_peak}}
        roofline_gather = {
            "bound": "hbm", "kernel": "agnn gather_reduce_kernel (all launches of the step)",
            "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            **traffic_of("gather_f16" if f16_ops else "gather_tf32"),
            "peak_source": peak_src, "launches": g["launches"],
            "algorithmic_bytes_per_step": g["bytes"] / max(args.steps, 1),
            "kernel_ms_per_step": g["ms"] / max(args.steps, 1),
            "share_of_step": g["ms"] / serial_ms if serial_ms else None, "timed_in": timed_in,
            "largest_launch": {"bytes": g["max_bytes"], "ms": g["max_ms"], "achieved": big, "frac": big / hbm_peak}}
        dominant_is_gemm = mm["ms"] >= g["ms"]
        cpu = cpu_arm(steps=2, warmup=1, budget_s=20.0, max_graphs=20) if not args.skip_cpu else \
            {"value": None, "unit": "nodes/s", "cores": os.cpu_count(), "kind": "port", "sample": "skipped (--skip-cpu)"}
    # ---- the other BASELINE configs, as sub-objects of the same line (bench_configs.py); collective ones on all ranks
    extras = {}
    main_routes = dict(_lib.library_routes)        # of the main arm alone (strict mode: must be empty)
    if not args.no_extras:
        import bench_configs as bc
        _lib.set_strict(False)                     # the sub-configs report their library routes instead of refusing them
        if use_graph:
            del g_resident, g_e2e
        torch.cuda.empty_cache()
        ctx = bc.Ctx(dev, world, rank, peaks()[0])
        extras["config4_dp"] = bc.guarded(bc.config4, ctx)
        if world > 1:
            extras["config1_strong"] = bc.guarded(bc.config1_strong, ctx, CFG, TASKS)
        elif rank == 0:
            extras["config1_strong"] = {"scaling": "strong", "n_gpus": 1, "ms_per_step": ms / args.steps,
                                        "nodes_per_s": n_nodes * args.steps / (ms * 1e-3),
                                        "workload": "= the main line at one GPU"}
            extras["e2e_loader"] = bc.guarded(bc.loader_e2e, ctx, CFG, TASKS)
            extras["config3_hgt"] = bc.guarded(bc.config3, ctx, CFG, TASKS)
            extras["config5_inference"] = bc.guarded(bc.config5, ctx)
            extras["library_baseline"] = bc.guarded(bc.library_baseline, ctx, CFG, TASKS)
    if rank == 0:
        value = world * n_nodes * args.steps / (ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": "nodes/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if dtype == torch.float32 else "bf16",
            "data": "synthetic", "config": workload_config(args),
            "edges_per_s": world * n_edges * CFG["layers"] * args.steps / (ms * 1e-3),
            "e2e": {"value": world * n_nodes * args.steps / (e2e_ms * 1e-3), "unit": "nodes/s",
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps,
                    "h2d": "pinned host -> one of two device staging sets on a copy stream, overlapped with the "
                           "previous step's compute; K copies for K timed steps, the first one exposed; the copies "
                           f"rotate through {N_HOST_BATCHES} DIFFERENT synthetic batches (padded to common shapes), so "
                           "every timed step trains on other data than the step before"},
            "gpu_launches": launches, "library_routes": main_routes,
            "per_step": {**gemm_stats, "amax_passes_by_shape_and_site": amax_sites},
            "host_enqueue_ms_per_step": enqueue_ms,
            "execution": ("CSR build + fwd + bwd replayed as one CUDA graph per step, then allreduce + fused "
                          "clip/AdamW launched eagerly" if use_graph else "eager launches"),
            "eager": {"ms_per_step": eager_ms / args.steps, "host_enqueue_ms_per_step": eager_enqueue_ms},
            # the dominant kernel of the step (largest share) and the aggregation kernel next to it
            "roofline": roofline_gemm if dominant_is_gemm else roofline_gather,
            "roofline_aggregation": roofline_gather, "roofline_gemm": roofline_gemm,
            "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "clocks": clocks, "wall_s": wall, "loss": final_loss,
        }
        line.update(extras)
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--operands", default=os.environ.get("AGNN_PARITY_OPERANDS", "f16"), choices=["tf32", "f16"],
                    help="operand form of the fp32 parity GEMMs in the message-passing layers: 3xTF32 or 3 x fp16 with "
                         "per-tensor power-of-two scales (same accuracy, twice the MMA rate)")
    ap.add_argument("--skip-cpu", action="store_true", help="leave out the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-extras", action="store_true",
                    help="main workload only: leave out the config 3 / 4 / 5, strong-scaling and library-baseline "
                         "sub-objects (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
