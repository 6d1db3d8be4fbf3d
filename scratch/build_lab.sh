#!/bin/bash
# build SYRK lab variants (scratch/syrk_lab.cu).  usage: build_lab.sh NAME[:extra -D flags] ...
cd "$(dirname "$0")/.."
for spec in "$@"; do
  v=${spec%%:*}; extra=""; [ "$spec" != "$v" ] && extra=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DSYRK_LAB_$v $extra -DLABNAME="\"$spec\"" scratch/syrk_lab.cu -o scratch/syrk_lab_$v &
done
wait
