#!/bin/bash
# round-2 profiling job (ONE GPU): ncu --set full of the kernels as benched, a cuBLAS reference under the same metrics,
# the launch list of the bench command, and the compute-sanitizer runs.  Everything lands in gpurun_out/.
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python scripts/gpu_cfg.py 5 2>&1 | tail -1
python __graft_entry__.py smoke 2>&1 | tail -2
export VQB_GRAPH=0
O=gpurun_out
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:vq_assign_kernel -s 3 -c 1 -f -o $O/r2_assign_benched python scripts/ncu_step.py vq > $O/ncu_assign.log 2>&1
timeout 600 $NCU -k regex:segsum_kernel -s 3 -c 1 -f -o $O/r2_segsum python scripts/ncu_step.py vq > $O/ncu_segsum.log 2>&1
timeout 600 $NCU -k regex:"rvq_accumulate" -s 1 -c 2 -f -o $O/r2_rvq_tail python scripts/ncu_step.py rvq > $O/ncu_rvq.log 2>&1
timeout 600 $NCU -k regex:"gemm|cutlass|nvjet|xmma" -s 2 -c 1 -f -o $O/r2_cublas_gemm python scripts/ncu_step.py gemm > $O/ncu_gemm.log 2>&1
for f in r2_assign_benched r2_segsum r2_rvq_tail r2_cublas_gemm; do python scripts/ncu_summary.py $O/$f.ncu-rep > $O/${f}_summary.txt 2>&1; done
unset VQB_GRAPH
VQB_BENCH_SKIP_E2E=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-sustained > $O/bench_ncu_r2.log 2>&1
python scripts/launch_summary.py $O/r2_launches_bench.csv > $O/r2_launches_bench_summary.txt 2>&1
cat $O/r2_launches_bench_summary.txt | head -30
# sanitizers on small shapes (every kernel of the step, both dtypes, streamed A, RVQ)
for tool in memcheck racecheck synccheck; do
  VQB_GRAPH=0 timeout 900 compute-sanitizer --tool $tool --print-limit 10 python scripts/san_small.py > $O/r2_sanitizer_$tool.txt 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|error" $O/r2_sanitizer_$tool.txt | tail -12
done
