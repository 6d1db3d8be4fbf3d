#!/bin/bash
cd "$(dirname "$0")/../.."
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/r2k_$tag.json 2> gpurun_out/r2k_$tag.err; }
for g in 2 3 4 5 6 8; do run c3_g$g $B --config c3 --chain-groups $g; done
for g in 2 3 4; do run c5_g$g $B --config c5 --chain-groups $g; done
run c3x32_g2 $B --config c3 --chains 32 --chain-groups 2
run c3x32_g3 $B --config c3 --chains 32 --chain-groups 3
