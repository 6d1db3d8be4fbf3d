#!/bin/bash
cd "$(dirname "$0")/../.."
T=${TAG:-r2o}
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],4), round(d["value"],1))
except Exception as ex: print("$tag", "failed", ex)
PY
}
export BNR_CHOL_SCHEDULE=2
for g in 2 3 4; do run c3_B_g$g $B --config c3 --chain-groups $g; done
run x32_B_g2 $B --config c3 --chains 32 --chain-groups 2
run x32_B_g3 $B --config c3 --chains 32 --chain-groups 3
run c5_B_g2 $B --config c5 --chain-groups 2
run c5_B_g3 $B --config c5 --chain-groups 3
unset BNR_CHOL_SCHEDULE
export BNR_CHOL_A_SIDE_HI=1
run x32_Ahi_g3 $B --config c3 --chains 32 --chain-groups 3
run c5_Ahi_g3 $B --config c5 --chain-groups 3
run c3_Ahi_g5 $B --config c3 --chain-groups 5
run c3_Ahi_g6 $B --config c3 --chain-groups 6
