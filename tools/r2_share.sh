#!/bin/bash
# Several batch-32 encodes in flight, each confined to a share of the SMs (SMELTER_SM_LIMIT): images/s of bench.py's device-resident leg.
# usage (GPU box): tools/r2_share.sh > gpurun_out/share_sweep.txt
run() {  # $1 = SM limit, $2 = encodes in flight
  out=$(SMELTER_SM_LIMIT=$1 timeout 300 python bench.py --steps 400 --no-cpu --no-extra --in-flight $2 2>/dev/null | tail -1)
  echo "$out" | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('sm_limit=%4s in_flight=%s  value=%8.0f img/s  ms_per_step=%.4f  one_in_flight_ms=%.4f  e2e=%8.0f' % ('$1', '$2', d['value'], d['ms_per_step'], d.get('one_in_flight', {}).get('ms_per_step', float('nan')), d['e2e']['value']))
"
}
run 148 1; run 148 2; run 148 3
run 74 2; run 74 3; run 74 4
run 48 3; run 48 4; run 48 6
run 36 4; run 36 6
