#!/bin/bash
# usage: tools/prof.sh <tag> : plain runs, ncu launch lists (config 3 at 64 and 8 chains, config 2) and one --set full
# capture each of the kernels DESIGN.md discusses (one chain group, so every launch covers all chains of the handle)
cd "$(dirname "$0")/.."
TAG=$1
BASE="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --profile-sweeps 1 --chain-groups 1"
for spec in "c3:--config c3" "c3x8:--config c3 --chains 8" "c2:--config c2" "c4:--config c4" "c5:--config c5"; do
  name=${spec%%:*}; extra=${spec#*:}
  CMD="$BASE $extra"
  $CMD > gpurun_out/plain_${name}_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 330 --csv --log-file gpurun_out/launches_${name}_$TAG.csv $CMD > gpurun_out/ncu1_${name}_$TAG.log 2>&1
done
CMD="$BASE --config c3"
for K in k_gram_syrk k_potf2_inv k_chol_update k_trsm_dmma k_bwd_stream k_gamma_gig k_uxi k_xmma; do
  $CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu2_${K}_$TAG.log 2>&1
  python tools/ncu_raw.py gpurun_out/prof_${K}_$TAG.ncu-rep > gpurun_out/raw_${K}_$TAG.txt 2>&1
  # (gpurun brings back at most 64 MiB: only the summaries travel, and the full report of the two headline kernels)
  case $K in k_gram_syrk|k_potf2_inv) ;; *) rm -f gpurun_out/prof_${K}_$TAG.ncu-rep ;; esac
done
CMD="$BASE --config c3 --chains 8"
for K in k_potf2_inv k_small_tile k_chol_update; do
  $CMD > gpurun_out/plain3_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 -f -o gpurun_out/prof_x8_${K}_$TAG $CMD > gpurun_out/ncu3_${K}_$TAG.log 2>&1
  python tools/ncu_raw.py gpurun_out/prof_x8_${K}_$TAG.ncu-rep > gpurun_out/raw_x8_${K}_$TAG.txt 2>&1
  case $K in k_potf2_inv) ;; *) rm -f gpurun_out/prof_x8_${K}_$TAG.ncu-rep ;; esac
done
ls -la gpurun_out/*_$TAG* | head -40
