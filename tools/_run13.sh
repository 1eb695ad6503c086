for dbg in 0 16 32 48 49; do
SMELTER_MEGA_DEBUG=$dbg timeout 300 python bench.py --no-cpu --steps 30 2>&1 | tail -1 > /tmp/b.json; python -c "
import json
d=json.load(open('/tmp/b.json')); print('dbg $dbg: ms',round(d['ms_per_step'],4))"
done
