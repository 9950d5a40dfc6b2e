#!/bin/bash
cd /root/repo
FN2_TC_DBG=16 timeout 300 python tools/corr_time.py 2>&1 | tail -7
timeout 600 python -m pytest tests -m gpu -q -x -k "corr" 2>&1 | tail -2
