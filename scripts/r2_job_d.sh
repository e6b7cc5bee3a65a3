#!/bin/bash
# GPU job: tests, sanitizers (racecheck / synccheck / memcheck), raw PCIe floor, module timings
TAG=${1:-x}
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
grep -E "passed|failed|rc=|^FAILED|^ERROR" $O/pytest_gpu_$TAG.log | tail -30
for tool in synccheck racecheck memcheck; do
  VQB_GRAPH=0 timeout 900 compute-sanitizer --tool $tool --print-limit 10 python scripts/san_small.py > $O/r2_sanitizer_$tool.txt 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|rror" $O/r2_sanitizer_$tool.txt | cut -c1-220 | tail -14
done
python scripts/gpu_pcie.py 2>&1 | tee $O/pcie_$TAG.txt
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; tail -3 $O/bench_$TAG.err; python scripts/show_bench.py $O/bench_$TAG.json 2>/dev/null
timeout 600 python scripts/gpu_configs.py > $O/configs_$TAG.jsonl 2>&1; cat $O/configs_$TAG.jsonl
