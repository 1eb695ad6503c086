timeout 300 python -m pytest tests/test_gpu_models.py -x -q 2>&1 | tail -3
for dbg in 0 4 8; do
SMELTER_MEGA_DEBUG=$dbg timeout 300 python bench.py --no-cpu --steps 30 2>&1 | tail -1 > /tmp/b.json; python -c "
import json
d=json.load(open('/tmp/b.json')); print('dbg $dbg: ms',round(d['ms_per_step'],4))"
done
