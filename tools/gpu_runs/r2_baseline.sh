#!/bin/bash
# round-2 baseline of the round-1 kernels: per-config phases + launch lists, few-chain C3, honest CPU baseline
cd "$(dirname "$0")/../.."
B="python tools/gpu_runs/bench_r1.py --no-cpu-baseline --steps 100 --warmup 6"
for c in c2 c4 c5; do
  $B --config $c > gpurun_out/r2base_$c.json 2> gpurun_out/r2base_$c.err
done
for g in 1 2 4; do
  $B --config c3 --chains 8 --chain-groups $g > gpurun_out/r2base_c3x8_g$g.json 2> gpurun_out/r2base_c3x8_g$g.err
done
$B --config c3 --chains 16 --chain-groups 4 > gpurun_out/r2base_c3x16_g4.json 2>&1
$B --config c3 --chains 32 --chain-groups 4 > gpurun_out/r2base_c3x32_g4.json 2>&1
$B --config c3 --chain-groups 3 > gpurun_out/r2base_c3_g3.json 2>&1
$B --config c3 --chain-groups 4 > gpurun_out/r2base_c3_g4.json 2>&1
python - <<'PY' > gpurun_out/r2base_cpu.json 2>&1
import sys, json, numpy as np
sys.path.insert(0, '.')
from bench import synth
from oracle import cpu_baseline as B
out = {}
for cfg, ch, sw in (("c3", 64, 10), ("c2", 16, 60)):
    X, y, d = synth(cfg)
    out[cfg] = B.time_port(np.ascontiguousarray(X), y, d["R"], ch, sw, warm=2)
print(json.dumps(out))
PY
for c in c4 c5; do
  CMD="python tools/gpu_runs/bench_r1.py --no-cpu-baseline --steps 12 --warmup 3 --profile-sweeps 1 --chain-groups 1 --config $c"
  $CMD > gpurun_out/r2base_plain_$c.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/r2base_launches_$c.csv $CMD > gpurun_out/r2base_ncu_$c.log 2>&1
done
nproc > gpurun_out/r2base_nproc.txt
