"""Print the headline fields of a bench.py JSON line: show_bench.py <file>"""
import json
import sys

d = json.load(open(sys.argv[1]))
print("step_ms", round(d["ms_per_step"], 4), "host_ms", round(d.get("host_ms_per_step", 0), 4), "kernel_ms",
      round(d["roofline"]["kernel_ms"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 3), "value", f'{d["value"]:.4g}',
      "clocks", d["clocks"]["sm_mhz"], d["clocks"]["samples"])
