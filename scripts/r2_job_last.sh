#!/bin/bash
# last GPU call of round 2 (under 2 GPU-minutes left): the whole GPU suite on the build with the in-kernel row mask, then a
# short device-resident bench (search kernel time must not have moved: the hot loops' SASS is unchanged)
O=gpurun_out; mkdir -p $O
timeout 120 python -m pytest tests -q -m gpu -rf > $O/last_all.log 2>&1; tail -12 $O/last_all.log
VQB_BENCH_SKIP_E2E=1 timeout 60 python bench.py --steps 30 --warmup 5 --no-sustained > $O/bench_short.json 2> $O/bench_short.err; tail -c 300 $O/bench_short.json
