"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/*_launches.csv):
time per kernel family and the share of this repo's kernels.  Usage: python profiles/summarize.py FILE [N]"""
import collections
import csv
import re
import sys


def main(path, top=25):
    with open(path) as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}[row["Metric Unit"]]
        name = row["Kernel Name"]
        m = re.search(r"(agnn::\S*?::)?(\w+)(<|\()", name)
        key = ("agnn::" if "agnn" in name else "") + (m.group(2) if m else name[:60])
        tot[key][0] += 1
        tot[key][1] += v
    total = sum(v[1] for v in tot.values())
    ours = sum(v[1] for k, v in tot.items() if k.startswith("agnn::"))
    print(f"total {total / 1e3:.2f} ms over {sum(v[0] for v in tot.values())} launches; "
          f"agnn kernels {ours / 1e3:.2f} ms ({100 * ours / total:.1f}%)")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t:10.1f} us {100 * t / total:5.1f}%  n={n:5d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
