#!/bin/bash
# strong-scaling shards (8 / 16 chains of config 3): SYRK k-split x chain groups
cd "$(dirname "$0")/../.."
T=${TAG:-r2h}
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100 --config c3"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],4), round(d["value"],1))
except Exception as ex: print("$tag", "failed", ex)
PY
}
for ch in 8 16; do
for sp in 1 2 4; do
for g in 2 4 8; do
  BNR_SYRK_SPLITS=$sp run x${ch}_s${sp}_g${g} $B --chains $ch --chain-groups $g
done; done; done
