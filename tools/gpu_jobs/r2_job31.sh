#!/bin/bash
# round-2 GPU job 31: last sanity after the sampler change -- reference suite over the shim, API tests, smoke
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_reference_suite.py tests/test_gpu_api.py tests/test_gpu_replay_edge_cases.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_pytest31.log
tail -2 gpurun_out/r2_pytest31.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
echo job31 done
