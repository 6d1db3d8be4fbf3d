#!/bin/bash
# round-2 GPU check #1: parity tests, then the new bench line and few-chain variants
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 100 --warmup 6 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
for g in 4 8; do
  timeout 200 python bench.py --config c3 --chains 8 --chain-groups $g --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 > gpurun_out/r2a_c3x8_g$g.json 2>&1
done
timeout 200 python bench.py --config c3 --chains 16 --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 > gpurun_out/r2a_c3x16.json 2>&1
timeout 200 python bench.py --config c3 --chains 32 --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 > gpurun_out/r2a_c3x32.json 2>&1
tail -3 gpurun_out/r2a_pytest.log
