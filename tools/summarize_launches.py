"""Per-kernel / per-grid summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [title] > profiles/<name>_summary.txt"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"ganffn::", "", name)
    return re.sub(r"\(.*$", "", name)[:90]


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        grid = r.get("Grid Size", "")
        rows.append((short(r["Kernel Name"]), grid, us))
    total = sum(r[2] for r in rows)
    print(f"# {title}")
    print(f"# {len(rows)} launches, {total / 1e3:.2f} ms summed (cold-cache, serialised by ncu)")
    by = defaultdict(lambda: [0.0, 0])
    for k, g, us in rows:
        by[k][0] += us
        by[k][1] += 1
    print("\n## by kernel")
    for k, (us, n) in sorted(by.items(), key=lambda kv: -kv[1][0]):
        print(f"{us:9.0f} us {100 * us / total:5.1f}%  n={n:5d}  avg={us / n:7.1f} us  {k}")
    byg = defaultdict(lambda: [0.0, 0])
    for k, g, us in rows:
        byg[(k, g)][0] += us
        byg[(k, g)][1] += 1
    print("\n## by kernel and grid (top 50)")
    for (k, g), (us, n) in sorted(byg.items(), key=lambda kv: -kv[1][0])[:50]:
        print(f"{us:9.0f} us {100 * us / total:5.1f}%  n={n:5d}  avg={us / n:7.1f} us  {k} grid={g}")


if __name__ == "__main__":
    main()
