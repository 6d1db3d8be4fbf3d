#!/bin/bash
# single-GPU validation of the round's final code: GPU test-suite, smoke(), the default bench line, the reference arm
cd "$(dirname "$0")/../.."
rm -f gpurun_out/parity_constants.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2z_smoke.log
timeout 300 bash tools/gpu_runs/r2_det.sh > gpurun_out/r2z_det.log 2>&1
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2z_bench_s20.json 2> gpurun_out/r2z_bench_s20.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2z_ref_s20.json 2> gpurun_out/r2z_ref_s20.err
tail -3 gpurun_out/r2z_pytest.log; tail -2 gpurun_out/r2z_smoke.log
