#!/bin/bash
cd "$(dirname "$0")/../.."
rm -f gpurun_out/parity_constants.jsonl
timeout 120 tools/lab/potf2_lab 8 1024 > gpurun_out/r2d_lab_8_1024.txt 2>&1
timeout 120 tools/lab/potf2_lab 16 512 > gpurun_out/r2d_lab_16_512.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/r2d_$tag.json 2> gpurun_out/r2d_$tag.err; }
run c3x8_g4 $B --config c3 --chains 8 --chain-groups 4
BNR_STRIPS=1 run c3x8_g4_strips1 $B --config c3 --chains 8 --chain-groups 4
BNR_STRIPS=2 run c3x8_g4_strips2 $B --config c3 --chains 8 --chain-groups 4
run c3x8_g2 $B --config c3 --chains 8 --chain-groups 2
run c3x8_g8 $B --config c3 --chains 8 --chain-groups 8
run c3x16_g4 $B --config c3 --chains 16 --chain-groups 4
run c3x32_g4 $B --config c3 --chains 32 --chain-groups 4
run c3 $B --config c3
run c3_g3 $B --config c3 --chain-groups 3
run c2 $B --config c2
BNR_STRIPS=1 run c2_strips1 $B --config c2
run c4 $B --config c4
run c4_g2 $B --config c4 --chain-groups 2
run c5 $B --config c5
run c5_g4 $B --config c5 --chain-groups 4
tail -5 gpurun_out/r2d_pytest.log
