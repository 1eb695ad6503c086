timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/conv_layers.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:100]); continue
    print(f\"{d['layer']:22s} {d['us']:7.2f}us {d['tflops']:7.1f}\")
"
for i in 1 2; do timeout 300 python bench.py --no-cpu --steps 300 2>&1 | tail -1 > /tmp/b.json; python -c "
import json
d=json.load(open('/tmp/b.json')); print('ms',round(d['ms_per_step'],4), 'img/s', round(d['value']), 'e2e', round(d['e2e']['value']))"; done
