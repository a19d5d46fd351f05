#!/bin/bash
# ncu --set full over the first <count> launches matching <regex> of a one-chunk batch call, exported as the raw-page csv
# (the .ncu-rep files are too big to travel back): tools/ncu_batch_csv.sh <frames> <levels> <out-prefix> <regex> <count>
NF=${1:-500}; LV=${2:-2}; OUT=${3:-gpurun_out/batch}; RE=$4; CNT=${5:-11}
export XPNGB_PIPE_LANES=1 XPNGB_LAT_MAX_BLOCKS=0
ncu --set full --clock-control none --import-source off --kernel-name "regex:$RE" --launch-count $CNT -f -o /tmp/ncu_batch \
    python tools/batch_prof.py $NF $LV noprof > $OUT.log 2>&1
ncu -i /tmp/ncu_batch.ncu-rep --page raw --csv > ${OUT}_raw.csv 2>/dev/null
rm -f /tmp/ncu_batch.ncu-rep
