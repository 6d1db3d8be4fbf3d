#!/bin/bash
cd "$(dirname "$0")/../.."
T=${TAG:-r2n}
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],4), round(d["value"],1))
except Exception as ex: print("$tag", "failed", ex)
PY
}
for hi in 0 1; do
if [ $hi = 1 ]; then export BNR_CHOL_A_SIDE_HI=1; fi
run c3_hi$hi $B --config c3
run c3_g3_hi$hi $B --config c3 --chain-groups 3
run c3_g4_hi$hi $B --config c3 --chain-groups 4
run x32_hi$hi $B --config c3 --chains 32
run c5_hi$hi $B --config c5
done
python tools/timeline.py --config c3 --sweeps 2 --out gpurun_out/${T}_tl_c3_hi.txt 2>/dev/null
