for bn in 0 64 128 256; do echo "== FORCE_BN $bn"; if [ $bn = 0 ]; then unset SMELTER_FORCE_BN; else export SMELTER_FORCE_BN=$bn; fi; python tools/conv_layers.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:100]); continue
    print(f\"{d['layer']:22s} {d['us']:7.2f}us {d['tflops']:7.1f}\")
"; done
