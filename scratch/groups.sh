for g in 2 3; do python bench.py --steps 60 --warmup 6 --no-cpu-baseline --chain-groups $g | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prio groups', $g, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; done
BNR_NO_PRIO=1 python bench.py --steps 60 --warmup 6 --no-cpu-baseline --chain-groups 2 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('noprio groups 2', d['value'], d['ms_per_step'])"
