#!/bin/bash
# A/B of environment switches on one box, default bench plan: usage tools/r2_env_ab.sh VAR=VAL ... (each argument one arm, "A=0" = default)
for arm in "$@"; do
  for rep in 1 2; do
    env $arm timeout 300 python bench.py --steps 400 --no-cpu --no-extra 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('%-24s value=%8.0f img/s  ms_per_step=%.4f  one_in_flight_ms=%.4f  e2e=%8.0f' % ('$arm', d['value'], d['ms_per_step'], d['one_in_flight']['ms_per_step'], d['e2e']['value']))
"
  done
done
