#!/bin/bash
# what the driver runs at round end, once more on the final tree: GPU tests, smoke(), both bench arms with default flags
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_reference_final.json 2> $O/bench_reference_final.err; tail -c 700 $O/bench_reference_final.json; echo
timeout 600 python bench.py > $O/bench_final.json 2> $O/bench_final.err; tail -2 $O/bench_final.err | cut -c1-300; python scripts/show_bench.py $O/bench_final.json
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_final.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "steps", "warmup", "gpu_launches")}, d["roofline"]["frac"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["clocks"], (d.get("sustained") or {}).get("ms_per_step"))
PY
