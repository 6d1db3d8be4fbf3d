#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_edges.py tests/test_gpu_engine.py -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 200"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/r2l_$tag.json 2> gpurun_out/r2l_$tag.err; }
run c3x8 $B --config c3 --chains 8
BNR_NO_SMALL_TILES=1 run c3x8_ring $B --config c3 --chains 8
run c3x16 $B --config c3 --chains 16
BNR_NO_SMALL_TILES=1 run c3x16_ring $B --config c3 --chains 16
run c2 $B --config c2
BNR_NO_SMALL_TILES=1 run c2_ring $B --config c2
run c4 $B --config c4
BNR_NO_SMALL_TILES=1 run c4_ring $B --config c4
tail -3 gpurun_out/r2l_pytest.log
