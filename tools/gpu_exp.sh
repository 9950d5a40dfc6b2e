#!/bin/bash
cd /root/repo
for env in "A=1" "FN2_TC_NOTAIL=1" "FN2_NO_STREAMS=1" "FN2_NO_STREAMS=1 FN2_TC_NOTAIL=1"; do
  echo "== $env"; env $env python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c60-170
done
python tools/layer_times.py --top 12 2>/dev/null | tail -13 | cut -c1-100
