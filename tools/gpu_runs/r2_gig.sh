#!/bin/bash
cd "$(dirname "$0")/../.."
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100"
for v in base gig3 gig4; do
  if [ $v = base ]; then unset LIBBNR; else export LIBBNR=$PWD/variants/libbnr_$v.so; fi
  for cfg in "c3" "c4" "c2" "c3 --chains 8"; do
    timeout 200 $B --config $cfg > gpurun_out/_g.json 2> gpurun_out/_g.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/_g.json").read().strip().splitlines()[-1])
print("$v", "$cfg", round(d["ms_per_step"],4), "gig phase", round(d["phases_ms"]["xt_gamma_gig"],4))
PY
  done
done
