#!/bin/bash
# round-2 GPU job 16: WARP replay with the positives row's first probe prefetched; ncu of the dataflow replay kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_replay_parity.py tests/test_gpu_replay_edge_cases.py tests/test_gpu_dropin.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_pytest16.log
tail -3 gpurun_out/r2_pytest16.log
timeout 300 python tools/bench_replay.py C2-slice > gpurun_out/r2_replay16.jsonl 2> gpurun_out/r2_replay16.err
cat gpurun_out/r2_replay16.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rdf_kernel -c 1 --launch-skip 1 \
    -o gpurun_out/r2_rdf_c1 -f python tools/bench_replay.py C1 > gpurun_out/r2_rdf_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_rdf_c1.ncu-rep 100000 20 > gpurun_out/r2_ncu_rdf_c1_summary.txt 2>&1
rm -f gpurun_out/r2_rdf_c1.ncu-rep
head -30 gpurun_out/r2_ncu_rdf_c1_summary.txt
echo job16 done
