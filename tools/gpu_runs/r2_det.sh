#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 200 python tools/determinism_check.py --config c3 --sweeps 100 --groups 0,0,0,1 2>&1 | tail -3
timeout 200 python tools/determinism_check.py --config c3 --chains 8 --sweeps 300 --groups 0,0,1 2>&1 | tail -2
timeout 200 python tools/determinism_check.py --config c3 --chains 16 --sweeps 200 --groups 0,0,1 2>&1 | tail -2
timeout 200 python tools/determinism_check.py --config c4 --sweeps 200 --groups 0,0,1 2>&1 | tail -2
timeout 200 python tools/determinism_check.py --config c5 --sweeps 100 --groups 0,0,1 2>&1 | tail -2
timeout 200 python tools/determinism_check.py --config c2 --sweeps 400 --groups 0,0,1 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_engine.py -q -x -k "reproducible" 2>&1 | tail -2
