# C2 step time against the SMs reserved for the upper layers (green-context partition), 1000 and 20 steps
cd /root/repo
for r in 8 12 16 20 24; do
  for st in "1000 5" "20 5"; do set -- $st
    python bench.py --steps $1 --warmup $2 --no-extras --no-cpu-baseline --pipeline-reserve $r 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('reserve', $r, 'steps', d['steps'], round(d['ms_per_step']*1e3,1), 'e2e', round(d['e2e']['ms_per_step']*1e3,1))"
  done
done
