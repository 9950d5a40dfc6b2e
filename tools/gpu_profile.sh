#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + --set full captures of the top kernels.
# Exports the raw / details pages on the box (the .ncu-rep files stay in gpurun_out/, which is scratch).
mkdir -p gpurun_out
O=gpurun_out
# 1. launch list of the bench command (graph replay disabled so that every kernel is a separate launch record)
FN2_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv \
    --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
# 2. tcgen05 convolution: NT=128 (conv3_1: 473->256 3x3, 135 K steps per tile) and NT=16 (fusion interconv0: 82->16)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -f -o $O/prof_tc128 \
    python tools/tc_time.py > $O/prof_tc128.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 29 -c 1 -f -o $O/prof_tc16 \
    python tools/tc_time.py > $O/prof_tc16.log 2>&1
# 3. correlation at the bench shape (4,256,56,128): the tcgen05 path (conv_tc_kernel<128> in correlation mode) and, for
#    reference, the FP32 FMA path it replaced
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -f -o $O/prof_corr \
    python tools/profile_ops.py > $O/prof_corr.log 2>&1
FN2_CORR_NOTC=1 timeout 600 ncu --set full --clock-control none -k regex:corr_fast_kernel -s 1 -c 1 -f -o $O/prof_corr_fp32 \
    python tools/profile_ops.py > $O/prof_corr_fp32.log 2>&1
for r in prof_tc128 prof_tc16 prof_corr prof_corr_fp32; do
    ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
    ncu -i $O/$r.ncu-rep --page details > $O/${r}_details.txt 2>/dev/null
done
ls -la $O | tail -14
