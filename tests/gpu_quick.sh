#!/bin/bash
# one gpurun call: GPU tests + smoke + layer times + short bench; writes logs under gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q ${PYTEST_ARGS} 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 600 python tools/layer_times.py > gpurun_out/layer_times.txt 2>&1; head -${LAYER_LINES:-45} gpurun_out/layer_times.txt | cut -c1-110
timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS} 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-300
tail -5 gpurun_out/bench.err
