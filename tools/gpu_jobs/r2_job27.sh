#!/bin/bash
# round-2 GPU job 27: feature path with the membership word fetched before the candidate's gather
mkdir -p gpurun_out
timeout 300 python tools/bench_c3.py 20000000 hot,nohot 2 > gpurun_out/r2_c3_job27.log 2>&1
tail -2 gpurun_out/r2_c3_job27.log | cut -c1-260
timeout 600 python -m pytest tests/test_gpu_features.py tests/test_gpu_hogwild.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_pytest27.log
tail -2 gpurun_out/r2_pytest27.log
echo job27 done
