#!/usr/bin/env python
"""A/B of the heavy-row constants (AGNN_HEAVY_ROW / AGNN_HEAVY_CHUNK are build-time): run under
AGNN_LIB_PATH=<variant .so>.  Prints the aggregation degree sweep (bench_extra.degree_sweep), a skewed multi-relation
COMBINE_SUM gather (the backward of hub source nodes) and the CSR build time of a config-1-sized typed graph."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench_extra  # noqa: E402
from analysisgnn_b200 import _lib, graph, ops  # noqa: E402

DEV = torch.device("cuda:0")


def skewed_sum(n_rel=6, n=60_000, e=400_000, f=256):
    rng = np.random.default_rng(1)
    rels, nbytes = [], 0
    for _ in range(n_rel):
        dst = np.minimum(rng.zipf(1.6, e) - 1, n - 1)
        src = rng.integers(0, n, e)
        ei = torch.as_tensor(np.stack((dst, src)), dtype=torch.long, device=DEV)
        csr = graph.TypedCSR(ei, None, n, n_cols=n)
        rels.append(ops.rel_of(csr.fwd, 0, torch.randn(n, f, device=DEV), n_edges=e))
    base = torch.randn(n, f, device=DEV)
    out = torch.empty(n, f, device=DEV)
    ms = bench_extra.timeit(lambda: ops.gather_reduce(rels, out, f, mean=True, concat=False, self_add=base))
    nbytes = ops.gather_bytes(rels, n, f, 4, False, True, False)
    return {"ms": ms, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / bench_extra.peak()}


def csr_build_ms(n=50_000, n_rel=12, e=2_600_000):
    rng = np.random.default_rng(2)
    ei = torch.as_tensor(rng.integers(0, n, (2, e)), dtype=torch.long, device=DEV)
    et = torch.as_tensor(rng.integers(0, n_rel, e), dtype=torch.long, device=DEV)

    def run():
        graph.clear_cache()
        graph.TypedCSR(ei, et, n, n_rel=n_rel).fwd
    return bench_extra.timeit(run)


if __name__ == "__main__":
    res = {"heavy": _lib.heavy_params(), "lib": _lib.LIB_PATH, "skewed_sum": skewed_sum()}
    try:
        res["csr_build_ms"] = csr_build_ms()
    except Exception as ex:  # noqa: BLE001
        res["csr_build_ms"] = repr(ex)
    res["sweep"] = [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()
                     if k in ("dist", "mean_in_degree", "ms", "frac")} for r in bench_extra.degree_sweep()]
    print(json.dumps(res))
