"""Turn an `ncu --csv --metrics a,b,c` log (one row per kernel and metric) into one line per launch:
    python tools/ncu_metric_table.py <ncu.csv> [kernel-substring]
Columns: id, kernel, grid, then every metric found in the log (durations in us, byte counts in MB)."""
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
want = sys.argv[2] if len(sys.argv) > 2 else ""
rows, order, metrics = {}, [], []
for r in csv.DictReader(lines):
    if want and want not in r["Kernel Name"]:
        continue
    k = r["ID"]
    if k not in rows:
        name = r["Kernel Name"].replace("void ", "").replace("<unnamed>::", "")
        rows[k] = {"kernel": name.split("(")[0][:28], "grid": r["Grid Size"]}
        order.append(k)
    m = r["Metric Name"]
    if m not in metrics:
        metrics.append(m)
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] in ("ns", "nsecond"):
        v /= 1e3
    elif r["Metric Unit"] in ("byte", "Byte"):
        v /= 1e6
    elif r["Metric Unit"] == "Kbyte":
        v /= 1e3
    elif r["Metric Unit"] == "Gbyte":
        v *= 1e3
    rows[k][m] = v
short = [m.replace("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_act_%")
          .replace("gpu__time_duration.sum", "time_us").replace("dram__bytes_read.sum", "dram_rd_MB")
          .replace("dram__bytes_write.sum", "dram_wr_MB").replace("lts__t_bytes.sum", "l2_MB") for m in metrics]
print("%-4s %-28s %-14s " % ("id", "kernel", "grid") + " ".join("%12s" % s for s in short))
for k in order:
    r = rows[k]
    print("%-4s %-28s %-14s " % (k, r["kernel"], r["grid"]) + " ".join("%12.1f" % r.get(m, float("nan")) for m in metrics))
