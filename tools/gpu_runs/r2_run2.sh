#!/bin/bash
cd "$(dirname "$0")/../.."
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/r2b_$tag.json 2> gpurun_out/r2b_$tag.err; }
for g in 1 2 4; do run c3x8_g$g $B --config c3 --chains 8 --chain-groups $g; done
for s in 1 2 4 8; do BNR_SYRK_SPLITS=$s run c3x8_g4_s$s $B --config c3 --chains 8 --chain-groups 4; done
BNR_SYRK_SPLITS=8 run c3x8_g2_s8 $B --config c3 --chains 8 --chain-groups 2
BNR_CHOL_SERIAL=1 run c3x8_g4_serial $B --config c3 --chains 8 --chain-groups 4
run c3 $B --config c3
run c3_g3 $B --config c3 --chain-groups 3
for c in c2 c4 c5; do run $c $B --config $c; done
for g in 1 2; do run c2_g$g $B --config c2 --chain-groups $g; done
for s in 2 4; do BNR_SYRK_SPLITS=$s run c4_s$s $B --config c4; done
run c4_g2 $B --config c4 --chain-groups 2
run c5_g4 $B --config c5 --chain-groups 4
CMD="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 8 --warmup 3 --profile-sweeps 1 --chain-groups 1"
for spec in "c3x8:--config c3 --chains 8" "c2:--config c2"; do
  tag=${spec%%:*}; extra=${spec#*:}
  $CMD $extra > gpurun_out/r2b_plain_$tag.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 120 -c 330 --csv --log-file gpurun_out/r2b_launches_$tag.csv $CMD $extra > gpurun_out/r2b_ncu_$tag.log 2>&1
done
