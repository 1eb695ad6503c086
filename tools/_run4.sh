for dbg in 0 16 64 80; do echo "== dbg $dbg"; SMELTER_CONV_DEBUG=$dbg python tools/conv_layers.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(f\"{d['layer']:22s} {d['us']:7.2f}us\")
"; done
echo "== NO_PDL"; SMELTER_NO_PDL=1 python tools/conv_layers.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(f\"{d['layer']:22s} {d['us']:7.2f}us\")
"
