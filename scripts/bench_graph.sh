#!/bin/bash
# graph on/off comparison + clean launch list of the device-resident step (diagnostic helper)
for g in 1 0; do
  VQB_GRAPH=$g VQB_BENCH_SKIP_E2E=1 timeout 200 python bench.py --steps 50 --warmup 10 2>&1 | tail -1 | sed "s/^/graph=$g /"
done
VQB_GRAPH=0 VQB_BENCH_SKIP_E2E=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches3.csv python bench.py --steps 3 --warmup 3 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches3.csv")) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ik][:48], []).append(float(r[iv].replace(",", "")))
tot = 0
for k, v in agg.items():
    print(f"{len(v):4d} x {sum(v)/len(v)/1e3:8.1f} us  {k}")
    if len(v) >= 6: tot += sum(v) / len(v) / 1e3
print("sum of per-step kernels (us):", round(tot, 1))
PY
