"""Summarises the `ncu --set full` capture of tools/hbm_kernels_once.py: per kernel (last = warm launch) the duration, DRAM bytes read /
written, DRAM and L2 throughput % and the achieved algorithmic GB/s against MEASURED_PEAKS.json.
usage: ncu -i gpurun_out/prof_hbm.ncu-rep --page raw --csv > raw.csv; python tools/summarize_hbm_ncu.py raw.csv alg.json out.md"""
import csv
import json
import os
import sys

COLS = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "CTAs"), ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram rd MB"),
        ("dram__bytes_write.sum", "dram wr MB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]


def num(v, unit):
  v = float(v.replace(",", ""))
  scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
  return v * scale.get(unit, 1.0)


def main():
  raw, alg_path, md = sys.argv[1:4]
  alg = json.load(open(alg_path))
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  pk = os.path.join(root, "MEASURED_PEAKS.json")
  peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
  rows = list(csv.reader(open(raw)))
  hdr, units, data = rows[0], rows[1], rows[2:]
  col = {h: i for i, h in enumerate(hdr)}
  last = {}
  for r in data:
    name = r[col["Kernel Name"]]
    key = next((k for k in alg if k in name), None)
    if key is None:
      key = name.split("(")[0].split("::")[-1][:40]
    grid = int(float(r[col["launch__grid_size"]].replace(",", "")))
    biggest = max((g for (k, g) in last if k == key), default = 0)
    if grid < biggest:
      continue                                                   # a smaller launch of the same kernel (e.g. the 128-RoI forward before the backward)
    if grid > biggest:
      last.pop((key, biggest), None)
    last[(key, grid)] = r                                        # keep the last (warm) launch of each kernel at its largest grid
  last = {k: v for (k, _), v in last.items()}
  out = ["| kernel | " + " | ".join(c[1] for c in COLS) + " | algorithmic MB | achieved GB/s | of measured peak (%.0f GB/s) | dram traffic / algorithmic |" % peak,
         "|" + "---|" * (len(COLS) + 5)]
  for key, r in last.items():
    vals = [num(r[col[c]], units[col[c]]) if c in col else float("nan") for c, _ in COLS]
    us = vals[0]
    a = alg.get(key)
    tail = "| – | – | – | – |"
    if a:
      gbs = a / (us * 1e-6) / 1e9
      tail = "| %.1f | %.0f | %.3f | %.2f |" % (a / 1e6, gbs, gbs / peak, (vals[3] + vals[4]) * 1e6 / a)
    out.append("| `%s` | " % key + " | ".join("%.1f" % v for v in vals) + " " + tail)
  open(md, "a").write("\n".join(out) + "\n")
  print("\n".join(out))


if __name__ == "__main__":
  main()
