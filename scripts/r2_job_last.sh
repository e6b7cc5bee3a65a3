#!/bin/bash
# last GPU call of round 2 (5 GPU-minutes left): the tests of the final three features first, then smoke(), then as much of
# the remaining suite as the clamp allows; everything is logged as it goes
O=gpurun_out; mkdir -p $O
timeout 150 python -m pytest tests/test_parity_gpu.py tests/test_features_gpu.py -x -q -m gpu -k "masked or layouts or heads or simvq or codebook_surface or contract" > $O/last_new.log 2>&1; tail -3 $O/last_new.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/last_smoke.log 2>&1; tail -2 $O/last_smoke.log
timeout 600 python -m pytest tests -x -q -m gpu > $O/last_all.log 2>&1; tail -3 $O/last_all.log
