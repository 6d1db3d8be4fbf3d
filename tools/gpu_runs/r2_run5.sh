#!/bin/bash
cd "$(dirname "$0")/../.."
rm -f gpurun_out/parity_constants.jsonl
timeout 120 tools/lab/potf2_lab 8 1024 > gpurun_out/r2e_lab_8_1024.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/r2e_$tag.json 2> gpurun_out/r2e_$tag.err; }
run c3 $B --config c3
run c3_g2 $B --config c3 --chain-groups 2
run c3x8 $B --config c3 --chains 8
BNR_CHOL_SCHEDULE=1 run c3x8_schedA $B --config c3 --chains 8
run c3x16 $B --config c3 --chains 16
BNR_CHOL_SCHEDULE=2 run c3x16_schedB $B --config c3 --chains 16
run c3x32 $B --config c3 --chains 32
run c2 $B --config c2
BNR_CHOL_SCHEDULE=1 run c2_schedA $B --config c2
run c4 $B --config c4
run c4_g2 $B --config c4 --chain-groups 2
BNR_CHOL_SCHEDULE=1 run c4_g2_schedA $B --config c4 --chain-groups 2
run c5 $B --config c5
run c5_g3 $B --config c5 --chain-groups 3
for i in 1 2 3; do BNR_TIMING=1 BNR_FIT_TIMING=1 run c3_s20_$i python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 20 --warmup 3; done
tail -5 gpurun_out/r2e_pytest.log
