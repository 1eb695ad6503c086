#!/bin/bash
# A/B of prebuilt library variants (tools/build_variant.sh) on one box: bench.py device-resident leg with the default plan
# (smShare 2, three in flight) and one encode at a time on the whole chip.
# usage: tools/r2_ab.sh NAME ...   ("default" = the in-tree library)
for name in "$@"; do
  if [ "$name" = default ]; then unset SMELTER_LIB_PATH; else export SMELTER_LIB_PATH=$PWD/smelter_b200/_variants/$name.so; fi
  for rep in 1 2; do
    timeout 300 python bench.py --steps 400 --no-cpu --no-extra 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('%-10s value=%8.0f img/s  ms_per_step=%.4f  one_in_flight_ms=%.4f  e2e=%8.0f' % ('$name', d['value'], d['ms_per_step'], d['one_in_flight']['ms_per_step'], d['e2e']['value']))
"
  done
done
