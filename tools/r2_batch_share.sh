#!/bin/bash
# which plan for which per-GPU batch: whole chip with 2 in flight vs half-chip plans with 3 in flight
run() {
  timeout 300 python bench.py --steps 150 --no-cpu --no-extra --batch $1 --sm-share $2 --in-flight $3 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('batch=%4s share=$2 in_flight=$3 value=%8.0f img/s  ms_per_step=%.4f  e2e=%8.0f' % ('$1', d['value'], d['ms_per_step'], d['e2e']['value']))
"
}
for b in 16 64 128; do run $b 1 2; run $b 2 3; run $b 2 4; done
