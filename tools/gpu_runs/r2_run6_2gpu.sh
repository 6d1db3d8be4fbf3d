#!/bin/bash
# two GPUs: library-side multi-device bnr_fit (NCCL / peer copies), the plain-C client, torchrun bench at N = 2
cd "$(dirname "$0")/../.."
nvidia-smi -L > gpurun_out/r2y_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_engine.py tests/test_gpu_multirank.py -m gpu -q -x -k "two_devices or c_client or two_ranks or fit_ess" > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --ess-burn 100 --ess-draws 200 > gpurun_out/r2y_bench_n2.json 2> gpurun_out/r2y_bench_n2.err
timeout 300 python bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2y_ref.json 2> gpurun_out/r2y_ref.err
tail -5 gpurun_out/r2y_pytest.log; tail -c 600 gpurun_out/r2y_bench_n2.json
