for h in 0 2 x 0 x; do if [ $h = x ]; then unset SMELTER_L2_HINTS; else export SMELTER_L2_HINTS=$h; fi; timeout 300 python bench.py --no-cpu --steps 300 2>&1 | tail -1 > /tmp/b.json; python -c "
import json
d=json.load(open('/tmp/b.json')); print('hints $h: ms',round(d['ms_per_step'],4), 'img/s', round(d['value']))"; done
