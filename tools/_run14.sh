timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --no-cpu --steps 50 2>&1 | tail -1 > gpurun_out/bench_s2d.json; python -c "
import json
d=json.load(open('gpurun_out/bench_s2d.json')); print('ms',round(d['ms_per_step'],4), 'img/s', round(d['value']), 'e2e', round(d['e2e']['value']))
p=json.load(open('gpurun_out/bench_profile_n1.json'))
for s in p['steps'][:4]+p['steps'][-2:]: print(round(s['ms']*1e3,1), s['desc'][:70])"
