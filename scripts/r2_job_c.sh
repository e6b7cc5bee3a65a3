#!/bin/bash
# lean GPU job: tests, bench line, GroupedResidualVQ host/GPU time, ResidualVQ launch list, module timings of every config
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "passed|failed|rc=|^FAILED|^ERROR" gpurun_out/pytest_gpu_$TAG.log | tail -30
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; python scripts/show_bench.py gpurun_out/bench_$TAG.json 2>/dev/null
python scripts/gpu_cfg.py 5 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cfg3_$TAG.csv python scripts/gpu_cfg.py 3 > /dev/null 2>&1; python scripts/launch_summary.py gpurun_out/launches_cfg3_$TAG.csv | head -8
timeout 600 python scripts/gpu_configs.py > gpurun_out/configs_$TAG.jsonl 2>&1; cat gpurun_out/configs_$TAG.jsonl
