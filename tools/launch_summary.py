"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the per-step sequence."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for x in csv.DictReader(lines):
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        name = re.sub(r"\(.*", "", x["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
        rows.append((name, x["Grid Size"], x["Block Size"], v))
    return rows


def main():
    rows = load(sys.argv[1])
    marker = sys.argv[2] if len(sys.argv) > 2 else "row_valid_kernel16"
    idx = [i for i, r in enumerate(rows) if marker in r[0]]
    steps = len(idx)
    if len(idx) >= 2:
        rows_step = rows[idx[-2]:idx[-1]]
    else:
        rows_step = rows
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, grid, blk, v in rows_step:
        key = name[:60]
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    print("one step (between the last two '%s' launches): %d launches, %.1f us of kernel time" % (marker, len(rows_step), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-62s n=%4d total=%9.1f us avg=%8.1f us %5.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    if "--seq" in sys.argv:
        for name, grid, blk, v in rows_step:
            print("%-52s grid=%-18s blk=%-12s %8.1f us" % (name[:52], grid, blk, v))


if __name__ == "__main__":
    main()
