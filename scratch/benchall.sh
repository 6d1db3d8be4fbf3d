# usage: scratch/benchall.sh [steps]  -- short bench lines for every config
S=${1:-60}
for c in c3 c2 c4 c5; do python bench.py --config $c --steps $S --warmup 6 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['phases_ms'].items()}, round(d['roofline']['frac'],3), d['status_or'])"; done
