#!/bin/bash
# ncu --set full capture of the taps-on-N kernel on fuse_interconv0 (82->16, 1024x448, 4 pairs): tools/tn_time.py launches it 13 times
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tn_kernel -s 5 -c 1 -f -o gpurun_out/prof_tn16 \
    python tools/tn_time.py > gpurun_out/prof_tn16.log 2>&1
ncu -i gpurun_out/prof_tn16.ncu-rep --page raw --csv > gpurun_out/prof_tn16_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_tn16.ncu-rep --page details > gpurun_out/prof_tn16_details.txt 2>/dev/null
ncu -i gpurun_out/prof_tn16.ncu-rep --page source --csv > gpurun_out/prof_tn16_source.csv 2>/dev/null
tail -3 gpurun_out/prof_tn16.log
