#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_edges.py tests/test_gpu_engine.py -m gpu -q -x > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/r2j_$tag.json 2> gpurun_out/r2j_$tag.err; }
run c3 $B --config c3
run c3_g2 $B --config c3 --chain-groups 2
run c3x8 $B --config c3 --chains 8
run c3x16 $B --config c3 --chains 16
run c3x32 $B --config c3 --chains 32
run c5 $B --config c5
run c4 $B --config c4
run c2 $B --config c2
tail -3 gpurun_out/r2j_pytest.log
