#!/bin/bash
run() {  # env... -- share in_flight
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --steps 400 --no-cpu --no-extra --sm-share $1 --in-flight $2 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('%-22s share=$1 in_flight=$2 value=%8.0f img/s  ms_per_step=%.4f  e2e=%8.0f' % ('${envs[*]}', d['value'], d['ms_per_step'], d['e2e']['value']))
"
}
run A=pdl -- 2 3
run SMELTER_NO_PDL=1 -- 2 3
run SMELTER_NO_PDL=1 -- 2 4
run SMELTER_NO_PDL=1 -- 2 6
run SMELTER_NO_PDL=1 -- 3 5
run SMELTER_NO_PDL=1 -- 3 6
run SMELTER_NO_PDL=1 -- 1 3
run A=pdl -- 2 4
