#!/bin/bash
# round-2 multi-GPU job (gpurun --gpus N -- 'bash scripts/r2_job_mg8.sh <tag> N'): the default workload (cfg2) with the
# peer-memory EMA and with the NCCL fallback, --workload cfg5; at N=2 also the multi-GPU parity tests
TAG=${1:-mg8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
run() {  # n workload extra-env
  local n=$1 wl=$2 tag=$3; shift 3
  if [ "$n" = 1 ]; then
    env "$@" timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-sustained --workload $wl > gpurun_out/bench_${wl}_n${n}_${tag}.json 2> gpurun_out/bench_${wl}_n${n}_${tag}.err
  else
    env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 5 --no-sustained --workload $wl > gpurun_out/bench_${wl}_n${n}_${tag}.json 2> gpurun_out/bench_${wl}_n${n}_${tag}.err
  fi
  echo "== $wl N=$n $*"; tail -2 gpurun_out/bench_${wl}_n${n}_${tag}.err | cut -c1-300; python scripts/show_bench.py gpurun_out/bench_${wl}_n${n}_${tag}.json 2>/dev/null || tail -c 600 gpurun_out/bench_${wl}_n${n}_${tag}.json
}
N=${2:-8}
run $N cfg2 $TAG VQB_X=1
run $N cfg2 ${TAG}_nccl VQB_NO_PEER=1 VQB_BENCH_SKIP_E2E=1
run $N cfg5 $TAG VQB_X=1
if [ "$N" = 2 ]; then
  timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -x -s > gpurun_out/pytest_mg_$TAG.log 2>&1; grep -E "RESULT|passed|failed" gpurun_out/pytest_mg_$TAG.log | tail -8
fi
