#!/usr/bin/env python
"""Secondary measurements for the other BASELINE.json configs (not the driver's bench contract; see bench.py).

    python bench_extra.py [--out profiles/extra.json]

* degree sweep (config 5): agnn_gather_reduce on E = 2^22 edges, F = 256 fp32, mean in-degree 1..128,
  uniform and Zipf(1.2) destination distributions, random sources over a matrix larger than L2 --
  achieved algorithmic GB/s vs the measured HBM copy peak;
* HGT attention kernels (config 3): forward / bwd_dst / bwd_src on the config-1 graph, fp32 and bf16;
* full-score inference (config 5): 200 000-note score: GPU graph build -> CSR -> HybridGNN forward;
* in-tree MetricalGNN 4L/512, 64 x 500 notes (config 4): forward + backward step time;
* onset-wise logit aggregation + decode of a 200 000-note score.
Every timing: CUDA events on the launching stream, 3 warm-up + 10 timed runs, L2 flushed between runs.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from analysisgnn_b200 import _lib, graph, ops, scoregraph, synth  # noqa: E402
from analysisgnn_b200 import nn as ann  # noqa: E402

DEV = torch.device("cuda:0")


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.isfile(p) else 6650.0


_flush = None


def timeit(fn, n=10, warm=3):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=DEV)
    for _ in range(warm):
        fn()
    ms = []
    for _ in range(n):
        _flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def degree_sweep():
    out, pk = [], peak()
    e, f = 1 << 22, 256
    rng = np.random.default_rng(0)
    for dist in ("uniform", "zipf1.2"):
        for deg in (1, 2, 4, 8, 16, 32, 64, 128):
            n = e // deg
            n_src = max(n, 1 << 18)                                   # >= 256 MB of source rows: not L2-resident
            if dist == "uniform":
                dst = rng.integers(0, n, e)
            else:
                dst = np.minimum(rng.zipf(1.2, e) - 1, n - 1)
            src = rng.integers(0, n_src, e)
            ei = torch.as_tensor(np.stack((dst, src)), dtype=torch.long, device=DEV)
            csr = graph.TypedCSR(ei, None, n, n_cols=n_src)
            x = torch.randn(n_src, f, device=DEV)
            y = torch.empty(n, f, device=DEV)
            rel = [ops.rel_of(csr.fwd, 0, x, n_edges=e)]
            ms = timeit(lambda: ops.gather_reduce(rel, y, f, mean=True, concat=True))
            nbytes = ops.gather_bytes(rel, n, f, 4, True, False, False)
            out.append({"dist": dist, "mean_in_degree": deg, "rows": n, "ms": ms, "algorithmic_bytes": nbytes,
                        "achieved_gbs": nbytes / ms / 1e6, "frac_of_measured_peak": nbytes / ms / 1e6 / pk})
            del csr, x, y, ei
    return out


def hgt_kernels():
    b = synth.hetero_batch(100, 500, 0, add_beats=False, add_measures=False)
    res, pk = [], peak()
    heads, d = 4, 64
    n = b["batch_size"]
    ei = {k: v.to(DEV) for k, v in b["edge_index_dict"].items()}
    csr = graph.hetero_csr(ei, {"note": n})
    ets = list(ei.keys())
    e_tot = sum(v.shape[1] for v in ei.values())
    for dtype in (torch.float32, torch.bfloat16):
        eb = 4 if dtype == torch.float32 else 2
        q = torch.randn(n, heads * d, device=DEV).to(dtype).requires_grad_(True)
        ks = [torch.randn(n, heads * d, device=DEV).to(dtype).requires_grad_(True) for _ in ets]
        vs = [torch.randn(n, heads * d, device=DEV).to(dtype).requires_grad_(True) for _ in ets]
        ps = torch.ones(len(ets), heads, device=DEV) / 8.0
        fw = [csr.fwd[et] for et in ets]
        bw = [csr.bwd[et] for et in ets]
        f_ms = timeit(lambda: ops.hgt_attention(q.detach(), [k.detach() for k in ks], [v.detach() for v in vs], ps, fw, bw, heads))
        out = ops.hgt_attention(q, ks, vs, ps, fw, bw, heads)
        g = torch.randn_like(out)
        b_ms = timeit(lambda: torch.autograd.grad(out, [q] + ks + vs, g, retain_graph=True))
        row = heads * d * eb
        f_bytes = e_tot * (2 * row + 4) + n * (2 * row + 8 * heads)
        b_bytes = e_tot * (2 * row + 4) + n * (4 * row + 12 * heads) + e_tot * (2 * row + 4 + 12 * heads) + \
            len(ets) * n * 4 * row
        res.append({"dtype": str(dtype).split(".")[1], "nodes": n, "edges": e_tot, "relations": len(ets),
                    "fwd_ms": f_ms, "fwd_gbs": f_bytes / f_ms / 1e6, "fwd_frac": f_bytes / f_ms / 1e6 / pk,
                    "bwd_ms": b_ms, "bwd_gbs": b_bytes / b_ms / 1e6, "bwd_frac": b_bytes / b_ms / 1e6 / pk})
    return res


def full_score_inference():
    na = synth.synth_note_array(200_000, 5, 4)
    build_ms = timeit(lambda: scoregraph.score_graph_edges(na, DEV), n=5, warm=1)
    edges, _ = scoregraph.score_graph_edges(na, DEV)
    names = ["onset", "consecutive", "during", "rest"]
    ei = {("note", nm, "note"): edges[:2, edges[2] == k].contiguous() for k, nm in enumerate(names)}
    for k, nm in enumerate(names[1:], 1):
        ei[("note", nm + "_rev", "note")] = ei[("note", nm, "note")].flip(0).contiguous()
    torch.manual_seed(0)
    meta = (["note"], list(ei.keys()))
    net = ann.HybridGNN(meta, 256, 256, 3, dropout=0.0).to(DEV).eval()
    x = {"note": torch.randn(200_000, 256, device=DEV)}
    batch = {"note": torch.zeros(200_000, dtype=torch.long, device=DEV)}

    def fwd():
        graph.clear_cache()
        with torch.no_grad():
            return net.gnn(x, ei)

    gnn_ms = timeit(fwd, n=5, warm=2)
    return {"notes": 200_000, "edges_fwd": int(edges.shape[1]), "edges_with_rev": int(sum(v.shape[1] for v in ei.values())),
            "graph_build_ms_incl_h2d": build_ms, "csr_plus_sage_stack_3x256_fwd_ms": gnn_ms,
            "nodes_per_s_message_passing": 200_000 / gnn_ms * 1e3}


def metrical_gnn_step():
    b = synth.intree_batch(64, 500, 0, in_features=64, metrical=True)
    torch.manual_seed(0)
    net = ann.MetricalGNN(64, 512, 512, b["etypes"], num_layers=4, dropout=0.3, metrical=True).to(DEV)
    net.train()
    args = [b[k].to(DEV) for k in ("edge_index", "edge_type", "beat_nodes", "measure_nodes", "beat_edges", "measure_edges")]
    kw = {k: b[k].to(DEV) for k in ("beat_lengths", "measure_lengths")}
    x = b["x"].to(DEV)

    def step():
        graph.clear_cache()
        net.zero_grad(set_to_none=True)
        net(x, *args, **kw).square().mean().backward()

    ms = timeit(step, n=5, warm=2)
    n = x.shape[0]
    return {"nodes": n, "edges": int(b["edge_index"].shape[1]), "params": sum(p.numel() for p in net.parameters()),
            "fwd_bwd_ms": ms, "nodes_per_s": n / ms * 1e3}


def hgt_encoder_step(encoder_type="hgt"):
    """BASELINE.json configs[3]: the config[1] batch through the HGT encoder (3 layers, 256, 4 heads)."""
    import bench
    b = bench.make_batch(0, bench.CFG["graphs"])
    torch.manual_seed(0)
    net = ann.AnalysisEncoder(b["metadata"], bench.CFG["in_features"], bench.CFG["hidden"], bench.CFG["out"],
                              bench.TASKS, bench.CFG["layers"], dropout=bench.CFG["dropout"],
                              encoder_type=encoder_type).to(DEV)
    net.train()
    t = {k: v.to(DEV) for k, v in bench.batch_tensors(b).items()}
    d = bench.unflatten(t, b)

    from analysisgnn_b200 import linalg
    from analysisgnn_b200.train import DataParallelTrainer, GraphedStep
    trainer = DataParallelTrainer(net, lr=bench.CFG["lr"], weight_decay=bench.CFG["weight_decay"],
                                  max_norm=bench.CFG["max_norm"], world_size=1, collect_grads=True)

    def fwd_bwd(_=None):
        linalg.begin_step()
        graph.clear_cache()
        trainer.zero_grad()
        logits = net(d["pitch_spelling"], d["key_signature"], d["x_dict"], d["edge_index_dict"], d["batch_dict"],
                     d["batch_size"], None, None)
        loss = ann.multitask_ce(logits, d["labels"])
        loss.backward()
        trainer.collect()
        return loss

    def step():
        fwd_bwd()
        trainer.step()

    ms = timeit(step, n=5, warm=3)
    if os.environ.get("AGNN_PROFILE"):
        from torch.profiler import profile, ProfilerActivity
        shapes = bool(os.environ.get("AGNN_PROFILE_SHAPES"))
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=shapes) as prof:
            step()
            torch.cuda.synchronize()
        print(prof.key_averages(group_by_input_shape=shapes).table(sort_by="cuda_time_total", row_limit=160,
                                                                   max_name_column_width=60))
    graphed = GraphedStep(fwd_bwd, None)

    def graphed_step():
        graphed()
        trainer.step()

    ms_graph = timeit(graphed_step, n=10, warm=3)
    if os.environ.get("AGNN_TRACE"):                      # kernel timeline of two replayed steps (chrome trace)
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            graphed_step()
            graphed_step()
            torch.cuda.synchronize()
        prof.export_chrome_trace(os.environ["AGNN_TRACE"])
    n = b["batch_size"]
    return {"nodes": n, "edges": sum(v.shape[1] for v in b["edge_index_dict"].values()), "step_ms_eager": ms,
            "step_ms_graph": ms_graph, "nodes_per_s": n / ms_graph * 1e3}


def decode_full_score(n_notes=200_000, cpu_notes=20_000, cpu_port=False):
    """Onset-wise logit aggregation + decode (analysisgnn/models/analysis.py:44-101) of one 200 k-note score: the call
    as ``predict`` makes it (device tensors in, device tensors out; its own host synchronisations included), timed
    with CUDA events.  ``cpu_port`` (--with-cpu-port) adds the cpu_baseline leg: the CPU restatement (oracle/decode.py,
    pinned to the reference's function; test infrastructure, used here only as the reported baseline) on a bounded
    sample -- its loop over change points is O(segments x notes), so the sample is smaller and the per-note figure
    FAVOURS the CPU."""
    import time
    from types import SimpleNamespace
    from analysisgnn_b200 import decode

    def note_store(x, batch, onset_div, edge_index_dict):
        class Graph(dict):
            pass
        g = Graph(note=SimpleNamespace(x=x, batch=batch, onset_div=onset_div))
        g.edge_index_dict = edge_index_dict
        return g

    rna_keys = decode.RNA_KEYS

    def prep(case, dev):
        mv = lambda t: None if t is None else t.to(dev)
        g = note_store(mv(case["x"]), mv(case["batch"]), mv(case["onset_div"]),
                       {k: mv(v) for k, v in case["edge_index_dict"].items()})
        return {k: v.to(dev) for k, v in case["logits"].items()}, g

    case = synth.decode_case(n_notes, 21, smooth=12)
    logits0, g = prep(case, DEV)

    def run():
        graph.clear_cache()
        decode.onsetwise_logit_aggregation({k: v.clone() for k, v in logits0.items()}, g, batch_size=n_notes)

    ms = timeit(run, n=10, warm=3)
    cpu = None
    if cpu_port:
        from oracle import decode as odecode
        small = synth.decode_case(cpu_notes, 21, smooth=12)
        lc, gc = prep(small, "cpu")
        torch.set_num_threads(os.cpu_count())
        t0 = time.perf_counter()
        odecode.onsetwise_logit_aggregation(lc, gc, batch_size=cpu_notes)
        cpu_s = time.perf_counter() - t0
        cpu = {"notes": cpu_notes, "seconds": cpu_s, "notes_per_s": cpu_notes / cpu_s, "cores": os.cpu_count(),
               "kind": "port"}
    cols = sum(synth.DECODE_TASKS[k] for k in rna_keys)
    e_kept = int(case["edge_index_dict"][("note", "onset", "note")].shape[1])
    # algorithmic bytes of one call: the packed logits read (edges + self) and written by the mean, read and written
    # by the two softmaxes, read by the arg-max of one row per onset, rows copied by the assignment (read + write)
    alg = 4 * cols * (e_kept + 2 * n_notes) + 2 * 4 * cols * n_notes + 2 * 4 * cols * n_notes
    return {"notes": n_notes, "onset_edges": e_kept, "task_columns": cols, "gpu_ms": ms,
            "gpu_notes_per_s": n_notes / ms * 1e3, "algorithmic_gb": alg / 1e9,
            "achieved_gbs_whole_call": alg / ms / 1e6,
            "cpu_port": cpu}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only-decode", action="store_true")
    ap.add_argument("--with-cpu-port", action="store_true",
                    help="decode: also time the CPU restatement (oracle/decode.py) on a bounded sample")
    ap.add_argument("--only-encoders", action="store_true",
                    help="in-tree MetricalGNN 4L/512 and the HGT encoder step (AGNN_PARITY_OPERANDS=tf32|f16 selects "
                         "the operand form)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "extra.json"))
    ap.add_argument("--only-sweep", action="store_true")
    ap.add_argument("--only-hgt", action="store_true")
    ap.add_argument("--only-hybrid", action="store_true", help="config[1] step through the same harness (profiling)")
    args = ap.parse_args()
    if args.only_encoders:
        from analysisgnn_b200 import linalg
        r = {"parity_operands": linalg.parity_operands(), "metrical_gnn_4L512": metrical_gnn_step(),
             "hgt_encoder": hgt_encoder_step()}
        r["library_routes"] = dict(_lib.library_routes)
        print("encoders", r)
        with open(args.out, "w") as fh:
            json.dump(r, fh, indent=1)
        return
    if args.only_decode:
        r = decode_full_score(cpu_port=args.with_cpu_port)
        print("decode", r)
        with open(args.out, "w") as fh:
            json.dump({"decode_full_score": r}, fh, indent=1)
        return
    if args.only_hybrid:
        print("hybridgnn-encoder", hgt_encoder_step("hybridgnn"))
        return
    if args.only_hgt:
        print("hgt-encoder", hgt_encoder_step())
        return
    res = {"hbm_peak_gbs": peak(), "degree_sweep": degree_sweep()}
    if args.only_sweep:
        res.update(hgt_attention=[], full_score_inference={}, metrical_gnn_4L512={})
    else:
        res.update(hgt_attention=hgt_kernels(), full_score_inference=full_score_inference(),
                   metrical_gnn_4L512=metrical_gnn_step(), hgt_encoder=hgt_encoder_step(),
                   decode_full_score=decode_full_score(cpu_port=args.with_cpu_port))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)
    for r in res["degree_sweep"]:
        print(f"sweep {r['dist']:8s} deg {r['mean_in_degree']:4d}: {r['ms']:.3f} ms {r['achieved_gbs']:.0f} GB/s "
              f"({r['frac_of_measured_peak']:.2f})")
    for r in res["hgt_attention"]:
        print("hgt", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()})
    print("full-score", res["full_score_inference"])
    print("metricalgnn", res["metrical_gnn_4L512"])
    print("hgt-encoder", res.get("hgt_encoder"))
    print("decode", res.get("decode_full_score"))


if __name__ == "__main__":
    main()
