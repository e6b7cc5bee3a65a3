#!/bin/bash
# round-2 job 1: full GPU test suite, a quick bench line, then the per-role cycle accounting of the new epilogue
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2a.log
tail -25 gpurun_out/pytest_gpu_r2a.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -3 gpurun_out/bench_r2a.err; python scripts/show_bench.py gpurun_out/bench_r2a.json 2>/dev/null || cat gpurun_out/bench_r2a.json
bash scripts/roles_job.sh 2,0 1,0 > gpurun_out/roles_r2a.txt 2>&1; cat gpurun_out/roles_r2a.txt
