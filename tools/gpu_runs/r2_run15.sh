#!/bin/bash
cd "$(dirname "$0")/../.."
T=${TAG:-r2q}
echo skip syrk_lab
echo skip potf2_lab
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],4), round(d["value"],1), d["roofline"]["frac"])
except Exception as ex: print("$tag", "failed", ex)
PY
}
run c3 $B --config c3
run x32 $B --config c3 --chains 32
run x16 $B --config c3 --chains 16
run x8 $B --config c3 --chains 8
run c2 $B --config c2
run c4 $B --config c4
run c5 $B --config c5
