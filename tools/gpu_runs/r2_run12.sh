#!/bin/bash
cd "$(dirname "$0")/../.."
T=${TAG:-r2m}
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],4), round(d["value"],1))
except Exception as ex: print("$tag", "failed", ex)
PY
}
for same in 0 1; do
if [ $same = 1 ]; then export BNR_SIDE_HI_SAME=1; fi
run c2_same$same $B --config c2
run c4_same$same $B --config c4
run x8_same$same $B --config c3 --chains 8
run x16_same$same $B --config c3 --chains 16
run c5_same$same $B --config c5
run c3_same$same $B --config c3
done
