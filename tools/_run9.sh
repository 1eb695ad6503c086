timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --no-cpu --steps 50 2>&1 | tail -1 > gpurun_out/bench_mega.json; python -c "
import json
d=json.load(open('gpurun_out/bench_mega.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'], d['launches_per_step'])"
