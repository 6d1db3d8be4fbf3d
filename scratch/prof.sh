#!/bin/bash
# usage: scratch/prof.sh <tag> [kernel-regex] : plain run, launch list, full capture of one kernel
TAG=$1
KRE=${2:-k_gram_syrk}
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile-sweeps 1"
if [ -z "$SKIP_LIST" ]; then
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 240 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
fi
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 2 -f -o gpurun_out/prof_${KRE}_$TAG $CMD > gpurun_out/ncu2_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu2_$TAG.log
