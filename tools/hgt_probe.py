#!/usr/bin/env python
"""The HGT attention kernels of config 3 alone (fp32 and bf16), timed as CUDA-graph replays so that the host side of
``ops.hgt_attention`` (packing seven relations, allocating outputs) is not in the number: forward, and the backward pair
(bwd_dst + bwd_src), L2 flushed between replays."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from analysisgnn_b200 import graph, ops, synth  # noqa: E402


def replay_ms(fn, flush, n=10, warm=3):
    """Runs on the current (side) stream: the autograd graph being replayed was recorded there too."""
    s = torch.cuda.current_stream()
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        fn()
    for _ in range(warm):
        g.replay()
    ms = []
    for _ in range(n):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def main():
    dev = torch.device("cuda:0")
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(pk))["hbm_gbs"]) if os.path.isfile(pk) else 6650.0
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    b = synth.hetero_batch(100, 500, 0, add_beats=False, add_measures=False)
    heads, d = 4, 64
    n = b["batch_size"]
    ei = {k: v.to(dev) for k, v in b["edge_index_dict"].items()}
    csr = graph.hetero_csr(ei, {"note": n})
    ets = list(ei.keys())
    e_tot = sum(v.shape[1] for v in ei.values())
    res = {"nodes": n, "edges": e_tot, "relations": len(ets), "peak_gbs": peak}
    for dtype in (torch.float32, torch.bfloat16):
        eb = 4 if dtype == torch.float32 else 2
        mk = lambda: torch.randn(n, heads * d, device=dev).to(dtype).requires_grad_(True)
        q, ks, vs = mk(), [mk() for _ in ets], [mk() for _ in ets]
        ps = torch.ones(len(ets), heads, device=dev) / 8.0
        fw, bw = [csr.fwd[et] for et in ets], [csr.bwd[et] for et in ets]
        f_ms = replay_ms(lambda: ops.hgt_attention(q.detach(), [k.detach() for k in ks], [v.detach() for v in vs],
                                                   ps, fw, bw, heads), flush)
        out = ops.hgt_attention(q, ks, vs, ps, fw, bw, heads)
        g = torch.randn_like(out)
        b_ms = replay_ms(lambda: torch.autograd.grad(out, [q] + ks + vs, g, retain_graph=True), flush)
        row = heads * d * eb
        f_bytes = e_tot * (2 * row + 4) + n * (2 * row + 8 * heads)
        b_bytes = (e_tot * (2 * row + 4) + n * (4 * row + 12 * heads) + e_tot * (2 * row + 4 + 12 * heads)
                   + len(ets) * n * 4 * row)
        res["f32" if dtype == torch.float32 else "bf16"] = {
            "fwd_ms": f_ms, "fwd_frac": f_bytes / f_ms / 1e6 / peak, "bwd_ms": b_ms, "bwd_frac": b_bytes / b_ms / 1e6 / peak}
    print(json.dumps(res))


if __name__ == "__main__":
    with torch.cuda.stream(torch.cuda.Stream()):
        main()
