#!/bin/bash
# round-2 multi-GPU job (gpurun --gpus 2): peer-memory EMA parity, then the 2-GPU bench line
TAG=${1:-mg}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x -s > gpurun_out/pytest_mg_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mg_$TAG.log
grep -E "RESULT|passed|failed|rc=|Error|error" gpurun_out/pytest_mg_$TAG.log | tail -30
tail -40 gpurun_out/pytest_mg_$TAG.log | head -60
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-sustained > gpurun_out/bench_n2_$TAG.json 2> gpurun_out/bench_n2_$TAG.err; tail -5 gpurun_out/bench_n2_$TAG.err; python scripts/show_bench.py gpurun_out/bench_n2_$TAG.json
VQB_NO_PEER=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-sustained > gpurun_out/bench_n2_nccl_$TAG.json 2> gpurun_out/bench_n2_nccl_$TAG.err; python scripts/show_bench.py gpurun_out/bench_n2_nccl_$TAG.json
