#!/bin/bash
# Profile one build of BASELINE config 2 (run under gpurun).  Usage: tools/ncu_c2.sh <tag>
# 1) plain run must exit 0, 2) launch list with per-launch device time, 3) full capture of the bulk kernels.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python tools/profile_target.py"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_split_uniform|k_split_from_ascii|k_split_from_records|k_filter' \
    -c 6 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
