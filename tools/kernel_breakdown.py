"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name, share of the total."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=25):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"mb::\(anonymous namespace\)::", "", name)
        rows.append((name, v * scale))
    tot = sum(v for _, v in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, v in rows:
        agg[n][0] += 1
        agg[n][1] += v
    print(f"{len(rows)} launches, {tot / 1e3:.3f} ms total (serialised, cold-cache: compare SHARES)")
    print(f"{'kernel':70s} {'launches':>8s} {'ms':>9s} {'share':>7s} {'us/launch':>10s}")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{n[:70]:70s} {c:8d} {v / 1e3:9.3f} {100 * v / tot:6.1f}% {v / c:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
