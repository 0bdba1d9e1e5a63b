"""Development tool: cProfile of the host side of train_step (where does the Python time between kernel launches go).
Usage: python tools/cpu_profile.py [steps]"""
import cProfile
import os
import pstats
import sys
import time

import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def main():
  steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
  step_fn = bench.make_train_step(t.device("cuda:0"))
  for _ in range(5):
    step_fn()
  t.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(steps):
    step_fn()
  t.cuda.synchronize()
  print("wall %.3f ms/step" % ((time.perf_counter() - t0) / steps * 1e3))
  pr = cProfile.Profile()
  pr.enable()
  for _ in range(steps):
    step_fn()
  t.cuda.synchronize()
  pr.disable()
  st = pstats.Stats(pr)
  st.sort_stats("cumulative").print_stats(45)
  st.sort_stats("tottime").print_stats(35)


if __name__ == "__main__":
  main()
