#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 120 tools/lab/potf2_lab 8 1024 > gpurun_out/r2c_lab_8_1024.txt 2>&1
timeout 120 tools/lab/potf2_lab 16 512 > gpurun_out/r2c_lab_16_512.txt 2>&1
timeout 120 tools/lab/potf2_lab 64 1024 > gpurun_out/r2c_lab_64_1024.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/r2c_$tag.json 2> gpurun_out/r2c_$tag.err; }
run c3x8_g4 $B --config c3 --chains 8 --chain-groups 4
BNR_SYRK_STAGGER=1 run c3x8_g4_stag $B --config c3 --chains 8 --chain-groups 4
BNR_SYRK_STAGGER=1 BNR_SYRK_SPLITS=2 run c3x8_g4_stag_s2 $B --config c3 --chains 8 --chain-groups 4
BNR_SYRK_STAGGER=1 BNR_SYRK_SPLITS=4 run c3x8_g4_stag_s4 $B --config c3 --chains 8 --chain-groups 4
BNR_SYRK_STAGGER=1 run c3x8_g8_stag $B --config c3 --chains 8 --chain-groups 8
BNR_SYRK_STAGGER=1 BNR_SYRK_SPLITS=4 run c3x8_g8_stag_s4 $B --config c3 --chains 8 --chain-groups 8
BNR_SYRK_STAGGER=1 run c3x16_g4_stag $B --config c3 --chains 16 --chain-groups 4
run c3x16_g4 $B --config c3 --chains 16 --chain-groups 4
BNR_SYRK_STAGGER=1 run c3_g4_stag $B --config c3 --chain-groups 4
BNR_SYRK_STAGGER=1 run c3_g2_stag $B --config c3 --chain-groups 2
run c3_s20 python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 20 --warmup 3
run c4 $B --config c4
BNR_SYRK_STAGGER=1 run c4_stag $B --config c4
run c4_g2 $B --config c4 --chain-groups 2
tail -3 gpurun_out/r2c_pytest.log
