#!/bin/bash
# Round-2 ncu evidence for profiles/: launch lists (bench step, training step) and --set full captures of the dominant kernels,
# plus an explicit tensor-pipe counter pass.  The .ncu-rep files stay in gpurun_out/ (scratch).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TM="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32.sum,sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32.sum.per_second,sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32.avg.pct_of_peak_sustained_elapsed,sm__mem_tensor_reads.sum,sm__mem_tensor_writes.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.avg"
FN2_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 660 -c 450 --csv \
    --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_train.csv python tools/train_step_once.py 2 > $O/train_ncu.log 2>&1
cap() {  # name, kernel regex, skip, script, env...
    local name=$1 rx=$2 skip=$3 script=$4
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/$name python $script > $O/$name.log 2>&1
    ncu -i $O/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
    ncu -i $O/$name.ncu-rep --page details > $O/${name}_details.txt 2>/dev/null
}
cap r02_prof_tc128 conv_tc_kernel 3 tools/tc_time.py
cap r02_prof_corr conv_tc_kernel 1 tools/profile_ops.py
cap r02_prof_wgrad128 conv_tc_kernel 1 tools/wg_time.py
cap r02_prof_corr_bwd corr_bwd_fast_kernel 2 tools/wg_time.py
# tensor-pipe counters, explicit list: forward conv3_1 (tc_time.py launch 3), weight gradient conv3_1 and conv2, correlation forward
timeout 600 ncu --metrics $TM --clock-control none -k regex:conv_tc_kernel -s 3 -c 1 --csv --log-file $O/r02_tensor_pipe_tc128.csv python tools/tc_time.py > /dev/null 2>&1
timeout 600 ncu --metrics $TM --clock-control none -k regex:conv_tc_kernel -c 6 --csv --log-file $O/r02_tensor_pipe_wgrad.csv python tools/wg_time.py > /dev/null 2>&1
timeout 600 ncu --metrics $TM --clock-control none -k regex:conv_tc_kernel -s 1 -c 1 --csv --log-file $O/r02_tensor_pipe_corr.csv python tools/profile_ops.py > /dev/null 2>&1
ls -la $O | grep r02_ | awk '{print $5, $9}'
