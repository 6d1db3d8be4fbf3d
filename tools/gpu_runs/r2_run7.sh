#!/bin/bash
cd "$(dirname "$0")/../.."
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 200"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/r2g_$tag.json 2> gpurun_out/r2g_$tag.err; }
run c3x8 $B --config c3 --chains 8
for us in 200 400 500 700; do BNR_STAGGER_US=$us run c3x8_stag$us $B --config c3 --chains 8; done
BNR_STAGGER_US=900 run c3x8_g2_stag900 $B --config c3 --chains 8 --chain-groups 2
BNR_STAGGER_US=250 run c3x8_g8_stag250 $B --config c3 --chains 8 --chain-groups 8
run c3x16 $B --config c3 --chains 16
BNR_STAGGER_US=800 run c3x16_stag800 $B --config c3 --chains 16
run c4 $B --config c4
BNR_STAGGER_US=800 run c4_stag800 $B --config c4
run c2 $B --config c2
BNR_STAGGER_US=100 run c2_stag100 $B --config c2
run c3 $B --config c3 --steps 100
BNR_STAGGER_US=3600 run c3_stag3600 $B --config c3 --steps 100
for i in 1 2 3; do BNR_FIT_TIMING=1 run c3_s20_$i python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 20 --warmup 3; done
