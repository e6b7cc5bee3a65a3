#!/bin/bash
# Runs bench.py with both statistics modes and prints a one-line summary of each (diagnostic helper).
for m in 1 0; do
  VQB_STATS_MODE=$m timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_mode$m.json 2> gpurun_out/bench_mode$m.err
  python - "$m" <<'PY'
import json, sys
m = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/bench_mode{m}.json"))
    print("stats_mode", m, "value %.4g" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e_ms %.3f" % d["e2e"]["ms_per_step"],
          "e2e %.4g" % d["e2e"]["value"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"],
          "launches", d["gpu_launches"], d["clocks"], "cpu %.4g" % d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
except Exception as e:
    print("stats_mode", m, "FAILED", e)
    print(open(f"gpurun_out/bench_mode{m}.err").read()[-1500:])
PY
done
