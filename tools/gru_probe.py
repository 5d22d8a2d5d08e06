#!/usr/bin/env python
"""Standalone forward + backward of the bidirectional GRU layer of the sequence branch (100 sequences x 500
steps, hidden 128 per direction) -- used with ncu for work on agnn_gru_fwd / _bwd.

    python tools/gru_probe.py [--reps 5]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from analysisgnn_b200.nn.layers import GRU  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=100)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--inputs", type=int, default=256)
    ap.add_argument("--library", action="store_true", help="time torch.nn.GRU (cuDNN, TF32 off) on the same shapes too")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    x = torch.randn(args.batch, args.steps, args.inputs, device=dev, requires_grad=True)

    def time_layer(rnn, what):
        for _ in range(2):
            rnn(x)[0].square().mean().backward()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        for _ in range(args.reps):
            ev[0].record()
            y = rnn(x)[0]
            ev[1].record()
            y.square().mean().backward()
            ev[2].record()
            torch.cuda.synchronize()
            tf += ev[0].elapsed_time(ev[1])
            tb += ev[1].elapsed_time(ev[2])
        print(f"GRU layer B={args.batch} T={args.steps} H={args.hidden}x2: forward {tf / args.reps:.3f} ms, "
              f"backward {tb / args.reps:.3f} ms ({what})", flush=True)

    rnn = GRU(args.inputs, args.hidden, num_layers=1, batch_first=True, bidirectional=True).to(dev)
    time_layer(rnn, "projections on agnn_gemm + recurrence kernels; eager launches")
    if args.library:
        lib = torch.nn.GRU(args.inputs, args.hidden, num_layers=1, batch_first=True, bidirectional=True).to(dev)
        time_layer(lib, "torch.nn.GRU: cuDNN, TF32 off")


if __name__ == "__main__":
    main()
