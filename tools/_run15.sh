timeout 300 python bench.py --steps 100 2>&1 | tail -1 > gpurun_out/bench_r1b.json; python -c "
import json
d=json.load(open('gpurun_out/bench_r1b.json')); print('ms',round(d['ms_per_step'],4), 'img/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), d['clocks'])"
timeout 200 python -m pytest tests/test_gpu_models.py -q -k "mobilenet or transformer" 2>&1 | tail -2
