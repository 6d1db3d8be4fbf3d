#!/bin/bash
# chain-group sweep with the look-ahead panel kernel
cd "$(dirname "$0")/../.."
T=${TAG:-r2i}
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],4), round(d["value"],1))
except Exception as ex: print("$tag", "failed", ex)
PY
}
run x8_g3 $B --config c3 --chains 8 --chain-groups 3
run x16_g3 $B --config c3 --chains 16 --chain-groups 3
for g in 1 2 4 8; do run c2_g$g $B --config c2 --chain-groups $g; done
for g in 1 2 4; do run c4_g$g $B --config c4 --chain-groups $g; done
for g in 2 3 4; do run c5_g$g $B --config c5 --chain-groups $g; done
for g in 2 3 4; do run c3_g$g $B --config c3 --chain-groups $g; done
