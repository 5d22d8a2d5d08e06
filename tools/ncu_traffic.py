#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --page raw --csv` export of tools/ncu_targets.py: DRAM bytes (read + write) of
the largest gemm_kernel and gather_reduce_kernel launches, with the tensor-pipe / DRAM utilisation next to them.

    ncu -i gpurun_out/r2_targets.ncu-rep --page raw --csv > profiles/r2_targets_raw.csv
    python tools/ncu_traffic.py profiles/r2_targets_raw.csv"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(v):
    try:
        return float(str(v).replace(",", ""))
    except ValueError:
        return None


def main(path):
    with open(path) as fh:
        rows = list(csv.reader(fh))
    head = rows[0]
    units = rows[1]
    col = {name: i for i, name in enumerate(head)}

    def val(r, name):
        i = col.get(name)
        if i is None:
            return None
        v, u = num(r[i]), units[i]
        if v is None:
            return None
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "us": 1e-6, "msecond": 1e-3,
                 "ms": 1e-3, "nsecond": 1e-9, "ns": 1e-9, "second": 1.0, "s": 1.0}.get(u, 1.0)
        return v * scale

    best = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        key = "gemm" if "gemm_kernel" in name else "gather" if "gather_reduce_kernel" in name else None
        if key is None:
            continue
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        dur = val(r, "gpu__time_duration.sum")
        if rd is None or wr is None:
            continue
        if key not in best or rd + wr > best[key]["dram_bytes"]:
            best[key] = {"dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "duration_s": dur,
                         "kernel": name[:80],
                         "sm__pipe_tensor_cycles_active_pct": val(
                             r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                         "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
    out = {}
    src = os.path.relpath(path, ROOT)
    if "gemm" in best:
        m, n, k = 50000, 256, 2560
        out["gemm_f16"] = dict(best["gemm"], source=f"{src} (tools/ncu_targets.py: 50000 x 256 x 2560, fp16 pairs)",
                               algorithmic_bytes=2 * 2 * (m * k + n * k) + 4 * m * n)
    if "gather" in best:
        out["gather_f16"] = dict(best["gather"], source=f"{src} (tools/ncu_targets.py: note rows of the config-2 batch, "
                                                         "9 relations, fp16 pair output)",
                                 algorithmic_bytes=1108569428,
                                 note="the 51 MB source matrix and the column ids stay in the 126 MB L2: DRAM sees the "
                                      "operand pair being written and little else")
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
