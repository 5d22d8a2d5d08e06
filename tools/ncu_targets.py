#!/usr/bin/env python
"""The two launches whose DRAM traffic bench.py's roofline objects quote, alone, for one `ncu --set full` capture:

    ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|gather_reduce_kernel' -c 8 \
        -o gpurun_out/r2_targets python tools/ncu_targets.py

(1) the largest GEMM of the step: 50 000 x 256 x 2560, fp16 hi / lo operand pairs (fp32 parity mode), as the fused
message-passing layer launches it; (2) the largest aggregation launch: one HeteroSAGELayer forward on the config-2
batch (9 relations into the 50 000 note rows, output written as the fp16 operand pair)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from analysisgnn_b200 import _lib, graph, linalg, synth  # noqa: E402
from analysisgnn_b200 import nn as ann  # noqa: E402

DEV = "cuda:0"


def main():
    torch.manual_seed(0)
    m, n, k = 50000, 256, 2560
    x = linalg.split_f16(torch.randn(m, k, device=DEV))
    w = linalg.split_f16(torch.randn(n, k, device=DEV) * 0.1)
    for _ in range(2):
        linalg.linear(x, w, torch.zeros(n, device=DEV), relu=True)
    b = synth.hetero_batch(100, 500, 1000, voices=4)
    layer = ann.HeteroSAGELayer(b["metadata"][1], 256, 256).to(DEV)
    g = torch.Generator().manual_seed(1)
    xd = {t: torch.randn(v.shape[0], 256, generator=g).to(DEV) for t, v in b["x_dict"].items()}
    ei = {et: v.to(DEV) for et, v in b["edge_index_dict"].items()}
    with torch.no_grad():
        for _ in range(2):
            graph.clear_cache()
            layer(xd, ei, relu=True)
    torch.cuda.synchronize()
    print("algorithmic bytes of the gemm launch:", 2 * 2 * (m * k + n * k) + 4 * m * n)


if __name__ == "__main__":
    main()
