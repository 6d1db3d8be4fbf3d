#!/bin/bash
# usage: scratch/prof.sh <tag> : plain run, ncu launch list, and one --set full capture each of the dominant kernels
# (all with one chain group so that every launch covers the handle's 64 chains, like bench.py's roofline phase)
TAG=$1
CMD="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --profile-sweeps 1 --chain-groups 1"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 260 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
for K in k_gram_syrk k_gamma_gig k_potf2_inv k_bwd_stream; do
  $CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu2_$TAG.log 2>&1
done
tail -n 2 gpurun_out/ncu2_$TAG.log
