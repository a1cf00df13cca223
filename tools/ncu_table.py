"""Summarise an `ncu --csv --metrics ...` log: one row per kernel name (launch count, total / mean time, share, DRAM
bytes, DRAM %, tensor-pipe %).   python tools/ncu_table.py gpurun_out/x.csv [skip_launches]"""
import csv, sys, re, collections
rows = collections.OrderedDict()
with open(sys.argv[1], newline="") as fh:
    lines = [l for l in fh if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
for i, r in rows.items():
    if i < skip:
        continue
    n = re.sub(r"\(.*", "", r["name"]).replace("void ", "").replace("b200::", "")
    n = re.sub(r"at::native::", "at::", n)[:90]
    a = agg.setdefault(n, {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0, "dram": [], "tc": [], "grid": r["grid"]})
    a["n"] += 1
    a["ns"] += r.get("gpu__time_duration.sum", 0.0)
    a["rd"] += r.get("dram__bytes_read.sum", 0.0)
    a["wr"] += r.get("dram__bytes_write.sum", 0.0)
    if "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed" in r:
        a["dram"].append(r["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"])
    if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in r:
        a["tc"].append(r["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"])
tot = sum(a["ns"] for a in agg.values())
print(f"{'kernel':90s} {'n':>4s} {'total us':>9s} {'mean us':>8s} {'share':>6s} {'rd MB':>8s} {'wr MB':>8s} {'dram%':>6s} {'tens%':>6s}  grid")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    d = sum(a["dram"]) / len(a["dram"]) if a["dram"] else float("nan")
    t = sum(a["tc"]) / len(a["tc"]) if a["tc"] else float("nan")
    print(f"{n:90s} {a['n']:4d} {a['ns']/1e3:9.1f} {a['ns']/1e3/a['n']:8.1f} {a['ns']/tot*100:5.1f}% {a['rd']/1e6:8.1f} {a['wr']/1e6:8.1f} {d:6.1f} {t:6.1f}  {a['grid']}")
print(f"{'TOTAL':90s} {sum(a['n'] for a in agg.values()):4d} {tot/1e3:9.1f}")
