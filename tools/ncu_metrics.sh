#!/bin/bash
# One metrics pass over every launch of a batch call: tools/ncu_metrics.sh <frames> <levels> <out.csv>
# (duration, warp instructions, issue-slot use, resident warps, DRAM bytes per launch; run under gpurun)
NF=${1:-1000}; LV=${2:-12}; OUT=${3:-gpurun_out/metrics.csv}
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__block_size \
    --clock-control none --csv --log-file $OUT python tools/batch_prof.py $NF $LV noprof > ${OUT%.csv}.log 2>&1
