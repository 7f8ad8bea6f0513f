#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_refsuite.py > gpurun_out/r2_diag_refsuite.log 2>&1
timeout 600 python tools/bench_c3.py 20000000 hot 2 > gpurun_out/r2_c3e.log 2>&1
echo job6 done
