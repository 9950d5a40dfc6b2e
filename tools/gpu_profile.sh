#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + --set full captures of the top kernels
mkdir -p gpurun_out
FN2_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 300 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_fast_kernel -s 3 -c 1 -f -o gpurun_out/prof_corr \
    python tools/profile_ops.py > gpurun_out/prof_corr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_nhwc_kernel -s 2 -c 1 -f -o gpurun_out/prof_conv \
    python tools/profile_ops.py > gpurun_out/prof_conv.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ops.csv \
    python tools/profile_ops.py > /dev/null 2>&1
ls -la gpurun_out | tail -12
