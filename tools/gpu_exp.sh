#!/bin/bash
cd /root/repo
timeout 600 python tools/tc_conv_debug.py 2>&1 | grep "tc err\|rror" | cut -c1-118
timeout 300 python tools/tc_time.py 2>&1 | grep "DBG=" | cut -c1-80
