"""gpurun_out/parity_margins.jsonl (written by the GPU tests through tests/_margins.py) -> a markdown table per test, latest record per
test name.  python tools/summarize_margins.py [jsonl] > profiles/r02_parity_margins.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fmt(v):
  if isinstance(v, float):
    return "inf" if v == float("inf") else ("%.3g" % v)
  return str(v)


def main():
  import glob
  paths = sys.argv[1:] if len(sys.argv) > 1 else sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "parity_margins_*.jsonl")))
  path = ", ".join(os.path.relpath(p, ROOT) for p in paths)
  latest = {}
  for one in paths:
    with open(one) as f:
      for line in f:
        line = line.strip()
        if line:
          row = json.loads(line)
          if row["test"] not in latest or row.get("when", "") >= latest[row["test"]].get("when", ""):
            latest[row["test"]] = row
  print("# Parity margins measured on the B200 (round 2)\n")
  print("Written by the `-m gpu` tests themselves (`tests/_margins.py`): what each end-to-end comparison MEASURED, next to the oracle's own")
  print("distance from every discontinuous decision (top-N cut, 16-px size filter, IoU 0.7).  The bars asserted in `tests/test_model_gpu.py`")
  print("are set from these numbers.  Source: `%s`.\n" % path)
  for name in sorted(latest):
    row = latest[name]
    print("## %s  (%s)\n" % (name, row.get("when", "")))
    print("| quantity | measured |")
    print("|---|---|")
    for k, v in row.items():
      if k in ("test", "when"):
        continue
      print("| %s | %s |" % (k, fmt(v)))
    print()


if __name__ == "__main__":
  main()
