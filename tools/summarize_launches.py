"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name: total us, count, share.

usage: python tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/rN_launch_list_summary.txt
Cold-cache, serialised timings: compare SHARES with the live numbers, not absolutes."""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"((?:sis3d::|at::native::|at::)?[\w:]+(?:<[^()]*?>)?)\(", name)
    return (m.group(1) if m else name)[:84]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
            rows.append((short(r["Kernel Name"]), us, r["Grid Size"]))
    rows = rows[skip:]
    tot = sum(u for _, u, _ in rows)
    agg = defaultdict(lambda: [0.0, 0])
    for n, u, _ in rows:
        agg[n][0] += u
        agg[n][1] += 1
    print(f"# {len(rows)} launches (first {skip} skipped), total {tot:.1f} us")
    print(f"{'us_total':>10} {'count':>6} {'us_avg':>8} {'share':>7}  kernel")
    for n, (u, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{u:10.1f} {c:6d} {u / c:8.2f} {100 * u / tot:6.1f}%  {n}")


if __name__ == "__main__":
    main()
