#!/bin/bash
# single-GPU job: tests, bench cfg2 + cfg5 at N=1, module timings
TAG=${1:-x}
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
grep -E "passed|failed|rc=|^FAILED|^ERROR" $O/pytest_gpu_$TAG.log | tail -30
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; tail -3 $O/bench_$TAG.err; python scripts/show_bench.py $O/bench_$TAG.json 2>/dev/null
timeout 300 python bench.py --steps 20 --warmup 5 --workload cfg5 --no-sustained > $O/bench_cfg5_n1_$TAG.json 2> $O/bench_cfg5_n1_$TAG.err; tail -c 900 $O/bench_cfg5_n1_$TAG.json | head -c 500; echo
timeout 600 python scripts/gpu_configs.py > $O/configs_$TAG.jsonl 2>&1; cat $O/configs_$TAG.jsonl
