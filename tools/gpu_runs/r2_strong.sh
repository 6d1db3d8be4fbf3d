#!/bin/bash
# per-GPU rates of the strong-scaling shards of config 3 (32 / 16 / 8 chains) on one B200
cd "$(dirname "$0")/../.."
T=${TAG:-r2s}
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100 --config c3"
for ch in 32 16 8; do
  timeout 200 $B --chains $ch > gpurun_out/${T}_x$ch.json 2> gpurun_out/${T}_x$ch.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_x$ch.json").read().strip().splitlines()[-1])
print("x$ch", round(d["ms_per_step"],4), round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "groups", d.get("chain_groups"))
PY
done
