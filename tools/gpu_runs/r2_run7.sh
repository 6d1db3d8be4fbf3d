#!/bin/bash
# look-ahead panel kernel: lab timings + host check, parity suite, small-config bench lines
cd "$(dirname "$0")/../.."
T=${TAG:-r2g}
timeout 120 tools/lab/potf2_lab 8 1024 > gpurun_out/${T}_lab_8_1024.txt 2>&1
timeout 120 tools/lab/potf2_lab 16 512 > gpurun_out/${T}_lab_16_512.txt 2>&1
head -60 gpurun_out/${T}_lab_8_1024.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; }
run c3 $B --config c3
run c3x8 $B --config c3 --chains 8
run c3x16 $B --config c3 --chains 16
run c2 $B --config c2
run c4 $B --config c4
run c5 $B --config c5
tail -5 gpurun_out/${T}_pytest.log
for f in c3 c3x8 c3x16 c2 c4 c5; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_$f.json").read().strip().splitlines()[-1])
    print("$f", d["ms_per_step"], d["value"], d.get("e2e",{}).get("value"))
except Exception as ex: print("$f", "failed", ex)
PY
done
