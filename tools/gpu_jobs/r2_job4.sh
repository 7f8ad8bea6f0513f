#!/bin/bash
# round-2 GPU job 4: suite, rank-kernel layouts (timing + ncu), C3, final-ish bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r2_pytest4.log
for g in 1 3; do timeout 300 python tools/bench_ranks.py 20000 $g >> gpurun_out/r2_ranks.log 2>&1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_ranks_tiled -c 1 \
    -o gpurun_out/r2_ranks_g1 -f python tools/bench_ranks.py 20000 1 > gpurun_out/r2_ranks_ncu.log 2>&1
timeout 600 python tools/bench_c3.py 20000000 hot 2 > gpurun_out/r2_c3d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hogwild_kernel -c 1 --launch-skip 1 \
    -o gpurun_out/r2_c3_hot_v3 -f python tools/bench_c3.py 4000000 hot 1 > gpurun_out/r2_c3_ncu_v3.log 2>&1
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
echo job4 done
