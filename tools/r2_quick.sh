#!/bin/bash
# quick A/B: isolated layers + bench for pair (default) and duo
echo "== layers pair"; timeout 300 python tools/conv_layers.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print(d['layer'].ljust(22), d['us'], d['tflops'])
    except Exception as e: print(l.strip()[:200])
"
echo "== bench pair"; timeout 300 python bench.py --steps 300 --no-cpu 2>/dev/null | cut -c1-200
echo "== bench duo"; SMELTER_DUO=1 timeout 300 python bench.py --steps 300 --no-cpu 2>/dev/null | cut -c1-200
echo "== dual duo"; SMELTER_DUO=1 python tools/dual_encode_probe.py 2>&1 | grep stream | head -3
echo "== dual pair"; python tools/dual_encode_probe.py 2>&1 | grep stream | head -3
