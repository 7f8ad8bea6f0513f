#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4; do timeout 120 python tools/diag_ranks.py >> gpurun_out/r2_diag_ranks.log 2>&1; done
echo job5 done
