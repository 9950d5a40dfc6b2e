#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for env in "A=1" "FN2_TC_PDL=0"; do echo "== $env"; env $env python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c60-175; done
