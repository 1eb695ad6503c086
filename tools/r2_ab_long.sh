#!/bin/bash
# A/B of prebuilt library variants over the default 1000-step run (the power cap bites after a few hundred steps): images/s and SM clock
for name in "$@"; do
  if [ "$name" = default ]; then unset SMELTER_LIB_PATH; else export SMELTER_LIB_PATH=$PWD/smelter_b200/_variants/$name.so; fi
  timeout 300 python bench.py --steps 1000 --no-cpu --no-extra 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('%-10s value=%8.0f img/s  ms_per_step=%.4f  one_in_flight_ms=%.4f  e2e=%8.0f  clocks=%s' % ('$name', d['value'], d['ms_per_step'], d['one_in_flight']['ms_per_step'], d['e2e']['value'], d['clocks']))
"
done
