#!/usr/bin/env python
"""Standalone timing of agnn_gemm on the shapes of the training step (used with ncu for kernel work).

    python tools/gemm_probe.py [--reps 20] [--one NAME] [--operands f16|tf32]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from analysisgnn_b200 import linalg  # noqa: E402

SHAPES = [  # name, kind, M, N, K
    ("fwd_256", "linear", 50000, 256, 256),
    ("fwd_2560", "linear", 50000, 256, 2560),
    ("fwd_wide", "linear", 50000, 4864, 256),
    ("dgrad_256", "mm", 50000, 256, 256),
    ("dgrad_2560", "mm", 50000, 2560, 256),
    ("wgrad_256", "mm_tn", 256, 256, 50000),
    ("wgrad_2560", "mm_tn", 256, 2560, 50000),
]
# exactly r persistent rounds of 148 tiles (N = 256 -> 2 column tiles): per-round time and fixed cost
ROUNDS = [(f"rounds_{r}_k{k}", "linear", 9472 * r, 256, k) for k in (256, 2560) for r in (1, 2, 3, 6)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--one", default=None)
    ap.add_argument("--rounds", action="store_true")
    ap.add_argument("--operands", default="f16", choices=["f16", "tf32"],
                    help="operand form of the parity mode: fp16 pairs (AGNN_GEMM_F16X3) or TF32 pairs (AGNN_GEMM_TF32X3)")
    args = ap.parse_args()
    split = linalg.split_f16 if args.operands == "f16" else linalg.split
    dev = torch.device("cuda:0")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    torch.manual_seed(0)
    for name, kind, m, n, k in (ROUNDS if args.rounds else SHAPES):
        if args.one and name != args.one:
            continue
        if kind == "linear":
            a = split(torch.randn(m, k, device=dev))
            b = split(torch.randn(n, k, device=dev))
            fn = lambda: linalg.linear(a, b, None)
        elif kind == "mm":
            a = split(torch.randn(m, k, device=dev))
            b = split(torch.randn(k, n, device=dev))
            fn = lambda: linalg.mm(a, b)
        else:
            a = split(torch.randn(k, m, device=dev))
            b = split(torch.randn(k, n, device=dev))
            fn = lambda: linalg.mm_tn(a, b)
        for _ in range(3):
            fn()
        times = {}
        for mode in ("warm", "flushed"):
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
            torch.cuda.synchronize()
            for s, e in ev:
                if mode == "flushed":
                    flush.fill_(1.0)
                s.record()
                fn()
                e.record()
            torch.cuda.synchronize()
            ts = sorted(s.elapsed_time(e) for s, e in ev)
            times[mode] = ts[len(ts) // 2] * 1e3
        fl = 2.0 * m * n * k
        print(f"{name:11s} M={m:6d} N={n:5d} K={k:6d}  warm {times['warm']:8.1f} us ({fl / times['warm'] / 1e6:6.1f} TF/s)"
              f"  flushed {times['flushed']:8.1f} us ({fl / times['flushed'] / 1e6:6.1f} TF/s)", flush=True)


if __name__ == "__main__":
    main()
