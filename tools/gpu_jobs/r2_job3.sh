#!/bin/bash
# round-2 GPU job 3: tier-B experiment (atomic accumulators x in-flight), suite, C3, C2 variants
mkdir -p gpurun_out
timeout 900 python tools/exp_tierb.py c2_warp c2_kos > gpurun_out/r2_exp_tierb.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | tail -200 > gpurun_out/r2_pytest3.log
timeout 600 python tools/bench_c3.py 20000000 hot,nohot 2 > gpurun_out/r2_c3c.log 2>&1
timeout 600 python tools/tune_warp.py 7,9 20000000 1 > gpurun_out/r2_tune_atomg.log 2>&1
timeout 600 python tools/tune_warp.py 7,9 20000000 0 >> gpurun_out/r2_tune_atomg.log 2>&1
echo job3 done
