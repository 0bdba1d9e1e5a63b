"""Summarises an `ncu --set full -k regex:tc_conv_kernel` capture of bench.py into profiles/: a markdown table per launch and
profiles/ncu_traffic.json (dram bytes per launch by kernel family, read by bench.py's roofline.traffic).

usage: ncu -i gpurun_out/prof_tc.ncu-rep --page raw --csv > /tmp/raw.csv
       python tools/summarize_ncu.py /tmp/raw.csv gpurun_out/launch_log.txt profiles/r01_tc_conv_ncu_full.md profiles/ncu_traffic.json "<source note>"
The launch log (FRCNN_LAUNCH_LOG, written by ops._gemm in launch order) attributes the i-th captured tc_conv_kernel launch to its
family (conv_fwd, linear_fwd, conv_dgrad ...); the capture must start at the first tc launch of the process (no --launch-skip)."""
import csv
import json
import sys

COLS = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "ctas"), ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "dram active %"),
        ("launch__registers_per_thread", "regs")]


def to_mb(v, unit):
  v = float(v.replace(",", ""))
  return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[unit]


def main():
  raw, log, md_path, json_path = sys.argv[1:5]
  source = sys.argv[5] if len(sys.argv) > 5 else raw
  rows = list(csv.reader(open(raw)))
  hdr, units, data = rows[0], rows[1], rows[2:]
  col = {h: i for i, h in enumerate(hdr)}
  fams = [l.split() for l in open(log)] if log != "-" else []
  out = ["| # | family | GFLOP | kernel | " + " | ".join(c[1] for c in COLS) + " | alg TF/s |", "|" + "---|" * (len(COLS) + 5)]
  agg = {}
  for i, r in enumerate(data):
    name = r[col["Kernel Name"]]
    short = name[name.index("tc_conv_kernel"):name.index(">") + 1] if "tc_conv_kernel" in name else name[:40]
    fam, gflop = (fams[i][0], float(fams[i][2])) if i < len(fams) else ("?", 0.0)
    vals = []
    for c, _ in COLS:
      v, u = r[col[c]], units[col[c]]
      vals.append(to_mb(v, u) if "bytes" in c else float(v.replace(",", "")))
    us = vals[0]
    if units[col["gpu__time_duration.sum"]] in ("ns", "nsecond"):
      us /= 1e3
      vals[0] = us
    tf = gflop / us * 1e-3 * 1e3 if us > 0 else 0.0
    out.append("| %d | %s | %.2f | `%s` | " % (i, fam, gflop, short) + " | ".join("%.1f" % v for v in vals) + " | %.1f |" % (gflop / (us * 1e-6) / 1e3 if us > 0 else 0))
    a = agg.setdefault(fam, dict(launches = 0, dram_mb = 0.0, us = 0.0, gflop = 0.0, tensor = 0.0))
    a["launches"] += 1; a["dram_mb"] += vals[2] + vals[3]; a["us"] += us; a["gflop"] += gflop; a["tensor"] += vals[4] * us
  out.append("")
  out.append("| family | launches | dram MB / launch | us / launch (under ncu, cold) | time-weighted tensor pipe % |")
  out.append("|---|---:|---:|---:|---:|")
  js = {}
  for fam, a in agg.items():
    out.append("| %s | %d | %.1f | %.1f | %.1f |" % (fam, a["launches"], a["dram_mb"] / a["launches"], a["us"] / a["launches"], a["tensor"] / max(a["us"], 1e-9)))
    js[fam] = dict(dram_bytes_per_launch = a["dram_mb"] * 1e6 / a["launches"], launches = a["launches"], source = source)
  open(md_path, "a").write("\n".join(out) + "\n")
  json.dump(js, open(json_path, "w"), indent = 1)
  print("\n".join(out[-(len(agg) + 3):]))


if __name__ == "__main__":
  main()
