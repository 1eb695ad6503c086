#!/bin/bash
# Build a variant of the library with compile-time switches (env) into a side directory and run layers + bench with it.
# usage: tools/r2_variant.sh NAME VAR=VAL ...   (run on the GPU box; nvcc is there too)
name=$1; shift
env "$@" python -m smelter_b200.build --force > /dev/null 2>&1 || { echo "build failed"; exit 1; }
echo "== variant $name ($*)"
timeout 300 python tools/conv_layers.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print(d['layer'].ljust(22), d['us'], d['tflops'])
    except Exception as e: print(l.strip()[:200])
"
timeout 300 python bench.py --steps 300 --no-cpu 2>/dev/null | cut -c1-200
