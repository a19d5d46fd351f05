#!/bin/bash
# ncu --set full captures of batch kernels: tools/ncu_batch.sh <frames> <levels> <out-prefix> <regex on the demangled kernel name> [skip] [count] [source on|off]
# (run under gpurun; the .ncu-rep lands in gpurun_out/ — keep a call's reports below 64 MiB in total: about 18 MiB per
# launch with source (three per call at most), 1-2 MiB without)
NF=${1:-256}; LV=${2:-1}; OUT=${3:-gpurun_out/batch}; RE=$4; SKIP=${5:-0}; CNT=${6:-3}; SRC=${7:-on}
ncu --set full --clock-control none --import-source $SRC --kernel-name-base demangled --kernel-name "regex:$RE" --launch-skip $SKIP --launch-count $CNT \
    -f -o $OUT python tools/batch_prof.py $NF $LV > $OUT.log 2>&1
