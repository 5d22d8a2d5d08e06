#!/usr/bin/env python
"""Run ONE of bench_configs.py's sub-objects on cuda:0 without the main bench line (development runs).

    python tools/config_probe.py config4 [--steps 5]
    python tools/config_probe.py config3 | config5 | library | loader
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bench_configs as bc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["config3", "config4", "config5", "library", "loader"])
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    from analysisgnn_b200 import _lib
    ctx = bc.Ctx(dev, 1, 0, bench.peaks()[0])
    _lib.reset_library_routes() if hasattr(_lib, "reset_library_routes") else None
    if args.which == "config4":
        out = bc.config4(ctx, steps=args.steps)
    elif args.which == "config3":
        out = bc.config3(ctx, bench.CFG, bench.TASKS)
    elif args.which == "config5":
        out = bc.config5(ctx)
    elif args.which == "library":
        out = bc.library_baseline(ctx, bench.CFG, bench.TASKS, steps=args.steps)
    else:
        out = bc.loader_e2e(ctx, bench.CFG, bench.TASKS, steps=args.steps)
    out["library_routes"] = dict(getattr(_lib, "library_routes", {}))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
