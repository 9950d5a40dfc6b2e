#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_tc \
    python tools/tc_time.py > gpurun_out/prof_tc.log 2>&1
tail -3 gpurun_out/prof_tc.log
