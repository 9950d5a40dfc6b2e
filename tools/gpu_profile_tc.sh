#!/bin/bash
# ncu --set full captures of the tcgen05 conv kernel: NT=128 (473->256, 135 steps/tile) and NT=16 (82->16, 27 steps/tile)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_tc128 \
    python tools/tc_time.py > gpurun_out/prof_tc128.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 29 -c 1 -f -o gpurun_out/prof_tc16 \
    python tools/tc_time.py > gpurun_out/prof_tc16.log 2>&1
tail -3 gpurun_out/prof_tc128.log gpurun_out/prof_tc16.log
