#!/bin/bash
# run bench.py a few times and print the headline fields (box-to-box and run-to-run variance check)
n=${1:-3}
for i in $(seq 1 $n); do
  timeout 300 python bench.py --steps ${2:-20} --warmup 5 2>/dev/null > /tmp/bench_rep.json
  python - <<'PY'
import json
d = json.load(open("/tmp/bench_rep.json"))
print("step_ms", round(d["ms_per_step"], 4), "host_ms", round(d.get("host_ms_per_step", 0), 4), "with_events",
      round(d["roofline"].get("ms_per_step_with_events", 0), 4), "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "e2e_ms",
      round(d["e2e"]["ms_per_step"], 3), "value", f'{d["value"]:.4g}', "clk", d["clocks"]["sm_mhz"], d["clocks"]["samples"])
PY
done
cp /tmp/bench_rep.json gpurun_out/bench_last.json
