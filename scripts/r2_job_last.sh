#!/bin/bash
# last GPU call of round 2: the whole GPU suite (the tests added last run last), smoke(), the default bench line
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests -q -m gpu -rf > $O/last_all.log 2>&1; tail -15 $O/last_all.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/last_smoke.log 2>&1; tail -2 $O/last_smoke.log
timeout 200 python bench.py > $O/bench_final.json 2> $O/bench_final.err; tail -c 600 $O/bench_final.json
