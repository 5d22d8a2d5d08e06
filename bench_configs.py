"""Measurements of the other BASELINE.json configs, emitted as sub-objects of bench.py's JSON line so that the
driver's own runs carry them:

* ``config3_hgt``      configs[2]: HGT 3L/256, 4 heads -- encoder-shell training step (fp32), HGT stack fwd+bwd in fp32
                       and in the stated bf16 mode, attention-kernel rooflines (algorithmic bytes / CUDA-event time);
* ``config5_inference`` configs[4]: 200 000-note score: GPU graph build + CSR + 3-layer SAGE forward, and the
                       aggregation degree sweep (E = 2^22, F = 256, uniform and Zipf(1.2) destinations);
* ``config4_dp``       configs[3]: in-tree MetricalGNN 4L/512 on 64 x 500 notes PER RANK, data-parallel step (weak
                       scaling) with the gradient allreduce timed on the device: blocking and overlapped forms;
* ``config1_strong``   SURVEY 8e: the 100-subgraph batch of the headline config split over the ranks;
* ``library_baseline`` the second bar of BASELINE.md: the oracle's encoder on the same B200 with PyTorch-eager CUDA
                       (ATen gather / index_add_, cuBLAS, cuDNN) -- "hand-written sm_100a vs library".

Every timing: CUDA events on the launching stream after warm-up, L2 flushed between timed iterations, max over ranks.
A failure inside one sub-measurement is reported in its object (``{"error": ...}``) and never hides the main line."""
import os
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))


class Ctx:
    """Device, distributed state and the shared timing helper."""

    def __init__(self, dev, world, rank, hbm_peak):
        self.dev, self.world, self.rank, self.hbm_peak = dev, world, rank, hbm_peak
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, value: float) -> float:
        if self.world == 1:
            return value
        import torch.distributed as dist
        t = torch.tensor([value], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def time_steps(self, fn, steps=10, warm=3, collective=False):
        """Total ms of ``steps`` calls (events around each call, L2 flush in between); with ``collective`` the result
        is the max over ranks and the timed region is bracketed by barriers."""
        for _ in range(warm):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier() if collective else torch.cuda.synchronize()
        for a, z in ev:
            self.flush.fill_(1.0)
            a.record()
            fn()
            z.record()
        self.barrier() if collective else torch.cuda.synchronize()
        ms = sum(a.elapsed_time(z) for a, z in ev)
        return self.max_over_ranks(ms) if collective else ms

    def replay_ms(self, fn, n=10, warm=3):
        """Median ms of ``fn`` replayed as a CUDA graph (L2 flushed between replays): device time of its kernels without
        the host side of the call.  Must run under a non-default ``torch.cuda.stream`` that also recorded whatever
        autograd graph ``fn`` walks."""
        s = torch.cuda.current_stream()
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
        for _ in range(warm):
            g.replay()
        out = []
        for _ in range(n):
            self.flush.fill_(1.0)
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            z.record()
            torch.cuda.synchronize()
            out.append(a.elapsed_time(z))
        return float(np.median(out))

    def median_ms(self, fn, n=10, warm=3):
        for _ in range(warm):
            fn()
        out = []
        for _ in range(n):
            self.flush.fill_(1.0)
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            z.record()
            torch.cuda.synchronize()
            out.append(a.elapsed_time(z))
        return float(np.median(out))


def guarded(fn, *a, **kw):
    from analysisgnn_b200 import _lib
    before = dict(_lib.library_routes)
    try:
        out = fn(*a, **kw)
    except Exception as exc:  # noqa: BLE001 -- a sub-measurement must never take the main line down
        return {"error": f"{type(exc).__name__}: {exc}", "trace": traceback.format_exc(limit=4)[-600:]}
    routes = {k: v - before.get(k, 0) for k, v in _lib.library_routes.items() if v != before.get(k, 0)}
    if isinstance(out, dict) and routes:
        out["library_routes"] = routes             # kernels of this measurement that were NOT this repo's
    return out


# ------------------------------------------------------------------------------------------ config 3: HGT

def _shell_step(ctx, encoder_type, graphs, seed, steps, cfg, tasks, use_graph=True):
    """Training step of the encoder shell (as bench.py's main arm) for another encoder type / batch size."""
    import bench
    from analysisgnn_b200 import nn as ann
    from analysisgnn_b200.train import DataParallelTrainer, GraphedStep
    b = bench.make_batch(seed, graphs)
    torch.manual_seed(0)
    net = ann.AnalysisEncoder(b["metadata"], cfg["in_features"], cfg["hidden"], cfg["out"], tasks, cfg["layers"],
                              dropout=cfg["dropout"], encoder_type=encoder_type).to(ctx.dev)
    net.train()
    trainer = DataParallelTrainer(net, lr=cfg["lr"], weight_decay=cfg["weight_decay"], max_norm=cfg["max_norm"],
                                  world_size=ctx.world, collect_grads=True)
    t = {k: v.to(ctx.dev) for k, v in bench.batch_tensors(b).items()}
    d = bench.unflatten(t, b)

    def fwd_bwd(_=None):
        trainer.zero_grad()
        logits = net(d["pitch_spelling"], d["key_signature"], d["x_dict"], d["edge_index_dict"], d["batch_dict"],
                     d["batch_size"], None, None)
        loss = ann.multitask_ce(logits, d["labels"])
        loss.backward()
        trainer.collect()
        return loss

    g = GraphedStep(fwd_bwd, None, warmup=2) if use_graph else None

    def step():
        if g is not None:
            g()
        else:
            from analysisgnn_b200 import graph, linalg
            graph.clear_cache()
            linalg.begin_step()
            fwd_bwd()
        trainer.step()

    ms = ctx.time_steps(step, steps=steps, warm=3, collective=ctx.world > 1)
    n_edges = sum(v.shape[1] for v in b["edge_index_dict"].values())
    return {"ms_per_step": ms / steps, "nodes_per_step_per_rank": b["batch_size"], "edges_per_rank": n_edges,
            "nodes_per_s": b["batch_size"] * steps / (ms * 1e-3), "steps": steps}


def hgt_attention_roofline(ctx):
    torch.cuda.synchronize()
    with torch.cuda.stream(torch.cuda.Stream()):
        res = _hgt_attention_roofline(ctx)
    torch.cuda.synchronize()
    return res


def _hgt_attention_roofline(ctx):
    from analysisgnn_b200 import graph, ops, synth
    b = synth.hetero_batch(100, 500, 0, add_beats=False, add_measures=False)
    heads, d = 4, 64
    n = b["batch_size"]
    ei = {k: v.to(ctx.dev) for k, v in b["edge_index_dict"].items()}
    csr = graph.hetero_csr(ei, {"note": n})
    ets = list(ei.keys())
    e_tot = sum(v.shape[1] for v in ei.values())
    res = {}
    for dtype in (torch.float32, torch.bfloat16):
        eb = 4 if dtype == torch.float32 else 2
        mk = lambda: torch.randn(n, heads * d, device=ctx.dev).to(dtype).requires_grad_(True)
        q, ks, vs = mk(), [mk() for _ in ets], [mk() for _ in ets]
        ps = torch.ones(len(ets), heads, device=ctx.dev) / 8.0
        fw, bw = [csr.fwd[et] for et in ets], [csr.bwd[et] for et in ets]
        f_ms = ctx.replay_ms(lambda: ops.hgt_attention(q.detach(), [k.detach() for k in ks], [v.detach() for v in vs],
                                                       ps, fw, bw, heads))
        out = ops.hgt_attention(q, ks, vs, ps, fw, bw, heads)
        g = torch.randn_like(out)
        b_ms = ctx.replay_ms(lambda: torch.autograd.grad(out, [q] + ks + vs, g, retain_graph=True))
        row = heads * d * eb
        f_bytes = e_tot * (2 * row + 4) + n * (2 * row + 8 * heads)
        b_bytes = (e_tot * (2 * row + 4) + n * (4 * row + 12 * heads) + e_tot * (2 * row + 4 + 12 * heads)
                   + len(ets) * n * 4 * row)
        res["f32" if dtype == torch.float32 else "bf16"] = {
            "fwd_ms": f_ms, "fwd_algorithmic_bytes": f_bytes, "fwd_gbs": f_bytes / f_ms / 1e6,
            "fwd_frac": f_bytes / f_ms / 1e6 / ctx.hbm_peak,
            "bwd_ms": b_ms, "bwd_algorithmic_bytes": b_bytes, "bwd_gbs": b_bytes / b_ms / 1e6,
            "bwd_frac": b_bytes / b_ms / 1e6 / ctx.hbm_peak}
    res.update(nodes=n, edges=e_tot, relations=len(ets), heads=heads, head_dim=d, bound="hbm", peak_gbs=ctx.hbm_peak,
               timing="CUDA-graph replays of the op (kernels only; the eager call is host-bound at this size: packing "
                      "seven relations costs more than the 0.14 ms forward kernel)",
               note="the note->note relations of the config-1 batch in one joint-softmax launch; q and the per-relation "
                    "k / v matrices total 0.77 GB in fp32, but neighbouring destination rows share source rows (L1 / L2 "
                    "hits), so DRAM-counter bytes are below the algorithmic bytes graded here and the fraction can "
                    "exceed 1")
    return res


def hgt_stack_fwd_bwd(ctx, dtype):
    """HGT message-passing stack alone (3 layers, 256, 4 heads) fwd + bwd on the config-2 graph in ``dtype``."""
    from analysisgnn_b200 import graph, linalg, synth
    from analysisgnn_b200 import nn as ann
    b = synth.hetero_batch(100, 500, 0, voices=4)
    torch.manual_seed(0)
    net = ann.hetero.HeteroHGTStack(b["metadata"], 256, 256, 3, 4).to(ctx.dev, dtype)
    g = torch.Generator().manual_seed(1)
    x = {k: torch.randn(v.shape[0], 256, generator=g).to(ctx.dev, dtype).requires_grad_(True)
         for k, v in b["x_dict"].items()}
    ei = {k: v.to(ctx.dev) for k, v in b["edge_index_dict"].items()}

    def step():
        graph.clear_cache()
        linalg.begin_step()
        net.zero_grad(set_to_none=True)
        out = net(x, ei, final_types=("note",))
        out["note"].float().square().mean().backward()

    ms = ctx.median_ms(step, n=5, warm=2)
    return {"ms": ms, "nodes_per_s": b["batch_size"] / ms * 1e3, "execution": "eager launches"}


def config3(ctx, cfg, tasks):
    out = {"workload": "BASELINE configs[2]: HGT 3L/256, 4 heads (D = 64), joint softmax over relations, config-2 "
                       "batch (100 x 500 notes + beat / measure nodes)"}
    out["encoder_step_f32"] = guarded(_shell_step, ctx, "hgt", cfg["graphs"], 1000, 5, cfg, tasks)
    out["stack_fwd_bwd_f32"] = guarded(hgt_stack_fwd_bwd, ctx, torch.float32)
    out["stack_fwd_bwd_bf16"] = guarded(hgt_stack_fwd_bwd, ctx, torch.bfloat16)
    out["attention_roofline"] = guarded(hgt_attention_roofline, ctx)
    return out


# ------------------------------------------------------------------------------------------ config 5: inference, sweep

def degree_sweep(ctx, degrees=(1, 2, 4, 8, 16, 32, 64, 128)):
    from analysisgnn_b200 import graph, ops
    out = []
    e, f = 1 << 22, 256
    rng = np.random.default_rng(0)
    for dist in ("uniform", "zipf1.2"):
        for deg in degrees:
            n = e // deg
            n_src = max(n, 1 << 18)                               # >= 256 MB of source rows: not L2-resident
            dst = rng.integers(0, n, e) if dist == "uniform" else np.minimum(rng.zipf(1.2, e) - 1, n - 1)
            src = rng.integers(0, n_src, e)
            ei = torch.as_tensor(np.stack((dst, src)), dtype=torch.long, device=ctx.dev)
            csr = graph.TypedCSR(ei, None, n, n_cols=n_src)
            x = torch.randn(n_src, f, device=ctx.dev)
            y = torch.empty(n, f, device=ctx.dev)
            rel = [ops.rel_of(csr.fwd, 0, x, n_edges=e)]
            ms = ctx.median_ms(lambda: ops.gather_reduce(rel, y, f, mean=True, concat=True), n=5, warm=2)
            nbytes = ops.gather_bytes(rel, n, f, 4, True, False, False)
            # rows actually distinct: an upper bound on what DRAM must deliver (repeated sources can hit L2)
            out.append({"dist": dist, "mean_in_degree": deg, "rows": n, "ms": ms, "algorithmic_bytes": nbytes,
                        "achieved_gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / ctx.hbm_peak,
                        "source_matrix_mb": n_src * f * 4 / 1e6,
                        "l2_resident_sources": bool(n_src * f * 4 < 100e6)})
            del csr, x, y, ei
    return out


def full_score_inference(ctx):
    from analysisgnn_b200 import graph, scoregraph, synth
    from analysisgnn_b200 import nn as ann
    na = synth.synth_note_array(200_000, 5, 4)
    build_ms = ctx.median_ms(lambda: scoregraph.score_graph_edges(na, ctx.dev), n=5, warm=1)
    edges, _ = scoregraph.score_graph_edges(na, ctx.dev)
    names = ["onset", "consecutive", "during", "rest"]
    ei = {("note", nm, "note"): edges[:2, edges[2] == k].contiguous() for k, nm in enumerate(names)}
    for k, nm in enumerate(names[1:], 1):
        ei[("note", nm + "_rev", "note")] = ei[("note", nm, "note")].flip(0).contiguous()
    torch.manual_seed(0)
    net = ann.HybridGNN((["note"], list(ei.keys())), 256, 256, 3, dropout=0.0).to(ctx.dev).eval()
    x = {"note": torch.randn(200_000, 256, device=ctx.dev)}

    def fwd():
        graph.clear_cache()
        with torch.no_grad():
            return net.gnn(x, ei)

    gnn_ms = ctx.median_ms(fwd, n=5, warm=2)
    e_rev = int(sum(v.shape[1] for v in ei.values()))
    return {"notes": 200_000, "edges_fwd": int(edges.shape[1]), "edges_with_rev": e_rev,
            "graph_build_ms_incl_h2d": build_ms, "csr_plus_sage_stack_3x256_fwd_ms": gnn_ms,
            "nodes_per_s_message_passing": 200_000 / gnn_ms * 1e3,
            "edges_per_s_message_passing": 3 * e_rev / gnn_ms * 1e3, "multi_gpu": "replicas only (one score per GPU)"}


def config5(ctx):
    from analysisgnn_b200 import graph
    old = graph.set_degree_bound(None)       # whole scores and Zipf degree tails: hub rows exist, keep their lists
    try:
        full, sweep = guarded(full_score_inference, ctx), guarded(degree_sweep, ctx)
    finally:
        graph.set_degree_bound(old)
    return {"workload": "BASELINE configs[4]: full-score inference on a synthetic 200 000-note score + node-degree "
                        "sweep of the aggregation kernel (E = 2^22, F = 256 fp32, mean aggregation)",
            "full_score": full,
            "degree_sweep": sweep,
            "degree_sweep_note": "frac = algorithmic bytes / CUDA-event time / measured HBM copy peak; at mean degree "
                                 ">= 32 the destination count shrinks and repeated source rows hit the 126 MB L2, so "
                                 "frac > 1 there is L2 bandwidth, not HBM evidence"}


# ------------------------------------------------------------------------------------------ config 4: MetricalGNN DP

def config4(ctx, steps=5, graphs_per_rank=64, layers=4, hidden=512, overlap=True):
    """In-tree MetricalGNN (hgnn.py:323-433) 4 layers / 512, 64 x 500 notes per rank: step = CSR build + forward +
    backward + NCCL gradient allreduce + fused clip / AdamW."""
    import torch.distributed as dist
    from analysisgnn_b200 import graph, linalg, synth
    from analysisgnn_b200 import nn as ann
    from analysisgnn_b200.train import DataParallelTrainer, GraphedStep
    b = synth.intree_batch(graphs_per_rank, 500, 100 + ctx.rank, in_features=64, metrical=True)
    torch.manual_seed(0)
    net = ann.MetricalGNN(64, hidden, hidden, b["etypes"], num_layers=layers, dropout=0.3, metrical=True).to(ctx.dev)
    net.train()
    trainer = DataParallelTrainer(net, lr=5e-3, weight_decay=5e-3, max_norm=1.0, world_size=ctx.world,
                                  collect_grads=True)
    args = [b[k].to(ctx.dev) for k in ("edge_index", "edge_type", "beat_nodes", "measure_nodes", "beat_edges",
                                       "measure_edges")]
    kw = {k: b[k].to(ctx.dev) for k in ("beat_lengths", "measure_lengths")}
    x = b["x"].to(ctx.dev)
    n = x.shape[0]

    def fwd_bwd(_=None):
        trainer.zero_grad()
        loss = net(x, *args, **kw).square().mean()
        loss.backward()
        trainer.collect()
        return loss

    execution = "CUDA graph (fwd + bwd) + eager allreduce / optimizer"
    try:
        g = GraphedStep(fwd_bwd, None, warmup=2)
        run = g
    except Exception as exc:  # noqa: BLE001 -- e.g. a library RNN that cannot be captured
        execution = f"eager launches (capture failed: {type(exc).__name__})"
        torch.cuda.synchronize()

        def run():
            graph.clear_cache()
            linalg.begin_step()
            return fwd_bwd()

    def step():
        run()
        trainer.step()

    ms = ctx.time_steps(step, steps=steps, warm=3, collective=True)
    # the collective alone, on the device: K back-to-back allreduces of the arena (nothing to overlap with) -- what a
    # blocking exchange exposes per step; and the step without it
    comm_ms = None
    if ctx.world > 1:
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        a.record()
        for _ in range(steps):
            dist.all_reduce(trainer.arena.grad, op=dist.ReduceOp.SUM)
        z.record()
        ctx.barrier()
        comm_ms = ctx.max_over_ranks(a.elapsed_time(z)) / steps
    keep, trainer.world_size = trainer.world_size, 1
    local_ms = ctx.time_steps(step, steps=steps, warm=1, collective=True)
    trainer.world_size = keep
    params = sum(p.numel() for p in net.parameters())
    return {"workload": "BASELINE configs[3]: in-tree MetricalGNN 4L/512 (7 edge types, metrical), 64 x 500 notes per "
                        "rank, fp32, train mode dropout 0.3; step = CSR build + fwd + bwd + allreduce + clip + AdamW",
            "scaling": "weak", "n_gpus": ctx.world, "ms_per_step": ms / steps, "nodes_per_rank": n,
            "nodes_per_s": ctx.world * n * steps / (ms * 1e-3), "params": params, "arena_mb": params * 4 / 1e6,
            "allreduce_alone_ms": comm_ms, "ms_per_step_without_allreduce": local_ms / steps,
            "exposed_collective_ms": (ms - local_ms) / steps if ctx.world > 1 else 0.0,
            "execution": execution, "steps": steps}


def config1_strong(ctx, cfg, tasks, steps=10):
    """SURVEY 8e strong scaling: the headline config's 100 subgraphs split over the ranks (rank r takes subgraphs
    ``{g : g mod W == r}``)."""
    from analysisgnn_b200.train import shard_indices
    mine = len(shard_indices(cfg["graphs"], ctx.rank, ctx.world))
    r = _shell_step(ctx, "hybridgnn", mine, 3000 + ctx.rank, steps, cfg, tasks)
    total_nodes = cfg["graphs"] * cfg["notes"]
    return {"workload": f"headline config, {cfg['graphs']} subgraphs in total split over {ctx.world} rank(s)",
            "scaling": "strong", "n_gpus": ctx.world, "subgraphs_this_rank": mine, "ms_per_step": r["ms_per_step"],
            "nodes_per_s": total_nodes / (r["ms_per_step"] * 1e-3)}


# ------------------------------------------------------------------------------------------ library baseline

def library_baseline(ctx, cfg, tasks, steps=5):
    """BASELINE.md's second bar: the oracle's restatement of the reference encoder (plain PyTorch modules) run on the
    SAME B200 with PyTorch-eager CUDA kernels (ATen index / index_add_, cuBLAS fp32 SIMT GEMM with TF32 off, cuDNN
    GRU), same batch, same step.  Separates "GPU vs CPU" from "hand-written sm_100a vs library".  The oracle is
    executed here only as a reported baseline, like cpu_baseline."""
    import bench
    from oracle import pyg as opyg
    b = bench.make_batch(1000, cfg["graphs"])
    torch.manual_seed(0)
    model = opyg.AnalysisEncoderShell(b["metadata"], cfg["in_features"], cfg["hidden"], cfg["out"], tasks,
                                      cfg["layers"], dropout=cfg["dropout"]).to(ctx.dev)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=cfg["lr"], weight_decay=cfg["weight_decay"], fused=True)
    mv = lambda d: {k: v.to(ctx.dev) for k, v in d.items()}
    ps, ks = b["pitch_spelling"].to(ctx.dev), b["key_signature"].to(ctx.dev)
    xd, ed, bd, lab = mv(b["x_dict"]), mv(b["edge_index_dict"]), mv(b["batch_dict"]), mv(b["labels"])

    def step():
        opt.zero_grad(set_to_none=True)
        logits = model(ps, ks, xd, ed, bd, b["batch_size"], None, None)
        loss = opyg.multitask_ce(logits, lab)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), cfg["max_norm"])
        opt.step()

    t0 = time.perf_counter()
    ms = ctx.time_steps(step, steps=steps, warm=3)
    return {"kind": "port on cuda (PyTorch eager: ATen, cuBLAS fp32 with TF32 off, cuDNN)", "ms_per_step": ms / steps,
            "nodes_per_s": b["batch_size"] * steps / (ms * 1e-3), "steps": steps,
            "wall_s_incl_warmup": time.perf_counter() - t0,
            "workload": "same batch and step as the main line (oracle/pyg.py AnalysisEncoderShell, fp32, train mode)"}


# ------------------------------------------------------------------------------------------ loader inside the step

def loader_e2e(ctx, cfg, tasks, steps=10, n_scores=200):
    """The headline step fed by the DEVICE loader: a corpus of synthetic scores (600-800 notes each) resident in HBM,
    every step a different batch -- 100 scores drawn by the epoch order, one 500-note window each (counter RNG on the
    host, 1.6 kB host -> device), window subgraph extraction + beat / measure nodes + feature / label gathers +
    CSR build + forward + backward replayed as ONE CUDA graph (sampler.StaticBatcher: static shapes, -1 padding), then
    allreduce / clip / AdamW, and the loss read back to pinned host memory.  Replaces graphmuse MuseNeighborLoader's
    worker processes (analysisgnn/data/datamodules/analysis.py:270-293) for full-window batches."""
    from analysisgnn_b200 import nn as ann, sampler, synth
    from analysisgnn_b200.train import DataParallelTrainer, GraphedStep
    t0 = time.perf_counter()
    c = synth.corpus(n_scores, lambda g: 600 + 25 * (g % 9), 4242 + ctx.rank, voices=cfg["voices"],
                     in_features=cfg["in_features"], task_dict=tasks)
    corpus = sampler.Corpus(c["x"].to(ctx.dev), c["edges"].to(ctx.dev), c["node_ptr"],
                            extras={k: v.to(ctx.dev) for k, v in c["extras"].items()})
    sb = sampler.StaticBatcher(corpus, cfg["notes"], cfg["graphs"], beat_of=c["beat_of"].to(ctx.dev),
                               measure_of=c["measure_of"].to(ctx.dev))
    loader = sampler.ScoreGraphLoader(corpus, cfg["notes"], cfg["graphs"], seed=7)
    build_s = time.perf_counter() - t0
    torch.manual_seed(0)
    net = ann.AnalysisEncoder(sb.metadata, cfg["in_features"], cfg["hidden"], cfg["out"], tasks, cfg["layers"],
                              dropout=cfg["dropout"]).to(ctx.dev)
    net.train()
    trainer = DataParallelTrainer(net, lr=cfg["lr"], weight_decay=cfg["weight_decay"], max_norm=cfg["max_norm"],
                                  world_size=ctx.world, collect_grads=True)
    sel_dev = sb.select(loader, 0, 0).to(ctx.dev)
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    def fwd_bwd(_=None):
        b = sb.batch(sel_dev)
        trainer.zero_grad()
        logits = net(b["extras"]["pitch_spelling"], b["extras"]["key_signature"], b["x_dict"], b["edge_index_dict"],
                     b["batch_dict"], b["batch_size"], None, None)
        loss = ann.multitask_ce(logits, {t: b["extras"][t] for t in tasks})
        loss.backward()
        trainer.collect()
        return loss

    g = GraphedStep(fwd_bwd, None, warmup=2)
    state = {"i": 0}
    per_epoch = len(loader)

    def step():
        i = state["i"]
        state["i"] += 1
        sel_dev.copy_(sb.select(loader, i // per_epoch, i % per_epoch), non_blocking=True)
        loss = g()
        trainer.step()
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)

    ms = ctx.time_steps(step, steps=steps, warm=3, collective=ctx.world > 1)
    n = cfg["graphs"] * cfg["notes"]
    return {"workload": "headline config fed by the device loader (StaticBatcher): a different 100 x 500-note batch "
                        "every step, sampler + collation inside the replayed graph",
            "value": ctx.world * n * steps / (ms * 1e-3), "unit": "nodes/s", "ms_per_step": ms / steps, "steps": steps,
            "h2d_bytes_per_step": int(sel_dev.numel() * 8), "d2h_bytes_per_step": 4,
            "corpus": {"scores": n_scores, "notes": int(c["node_ptr"][-1]), "edges": int(c["edges"].shape[1]),
                       "host_build_s": build_s},
            "static_capacity": {"note_note_edge_slots": int(sb.edge_cap) * 2,
                                **{k: int(v[1]) for k, v in sb.virtual.items()}},
            "loss": float(loss_host.item())}
