export SMELTER_CONV_INSTRUMENT=1
for v in 256; do for base in 32 1056; do dbg=$((v+base)); echo "== variant $((v/256)) dbg $base"; SMELTER_CONV_DEBUG=$dbg python tools/conv_layers.py 3x3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(f\"{d['layer']:22s} {d['us']:7.2f}us {d['tflops']:7.1f}\")
"; done; done
