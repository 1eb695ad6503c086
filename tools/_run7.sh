timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py -x -q 2>&1 | tail -5
python tools/conv_layers.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(f\"{d['layer']:22s} {d['us']:7.2f}us {d['tflops']:7.1f}\")
"
python bench.py --no-cpu --steps 50 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])"
