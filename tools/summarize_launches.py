"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`).

usage: python tools/summarize_launches.py X.csv ["header line"] > X_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((r["Kernel Name"], v * scale))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for name, ms in rows:
        short = re.sub(r"\(.*$", "", name)
        short = short.replace("oi::<", "").replace("oi::(anonymous namespace)::", "")
        tot[short] += ms
        cnt[short] += 1
    total = sum(tot.values())
    if len(sys.argv) > 2:
        print("# " + sys.argv[2])
    print(f"total {total:.1f} ms over {len(rows)} launches")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"{tot[k]:9.2f} ms {100 * tot[k] / total:5.1f}% n={cnt[k]:5d} avg={tot[k] / cnt[k]:8.4f} ms  {k}")


if __name__ == "__main__":
    main()
