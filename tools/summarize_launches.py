"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel markdown table.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [title]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
  path = sys.argv[1]
  rows = []
  with open(path, newline = "") as f:
    lines = [l for l in f if not l.startswith("==")]
  rd = csv.DictReader(lines)
  for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
      continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    name = re.sub(r"\(.*$", "", r["Kernel Name"])
    rows.append((name, us))
  agg = OrderedDict()
  for n, us in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += us
  tot = sum(a[1] for a in agg.values())
  print("| kernel | launches | total µs | share |")
  print("|---|---:|---:|---:|")
  for n, (c, us) in sorted(agg.items(), key = lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f %% |" % (n[:120], c, us, 100.0 * us / tot))
  print()
  print("Total %.0f µs over %d launches." % (tot, len(rows)))
  tc = sum(a[1] for n, a in agg.items() if "tc_conv_kernel" in n)
  print("tcgen05 implicit-GEMM kernels = %.1f %% of kernel time." % (100.0 * tc / tot))


if __name__ == "__main__":
  main()
