#!/bin/bash
# N-GPU bench line exactly as the driver launches it (weak scaling + the strong_scaling leg + e2e)
cd "$(dirname "$0")/../.."
N=${N:-4}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 --ess-burn 200 --ess-draws 400 --no-cpu-baseline > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_scale_n$N.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "weak", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1))
print("strong", d.get("strong_scaling"))
print("ess median", d["gamma_ess_per_sec"]["median"] if d.get("gamma_ess_per_sec") else None)
PY
tail -2 gpurun_out/r2_scale_n$N.err
