#!/bin/bash
cd "$(dirname "$0")/../.."
T=${TAG:-r2j}
for v in potf2; do
  timeout 120 tools/lab/${v}_lab 8 1024 > gpurun_out/${T}_${v}.txt 2>&1
  echo "== $v"; head -22 gpurun_out/${T}_${v}.txt | cat
done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],4), round(d["value"],1), d["config"].get("chain_groups"))
except Exception as ex: print("$tag", "failed", ex)
PY
}
run x8 $B --config c3 --chains 8
run x16 $B --config c3 --chains 16
run c2 $B --config c2
