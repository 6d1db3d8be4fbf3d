# usage: scratch/quick.sh [bench args...]  -- one short C3 line
python bench.py --steps 60 --warmup 6 --no-cpu-baseline "$@" | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms'].items()}, round(d['roofline']['frac'],3), d['status_or'])"
