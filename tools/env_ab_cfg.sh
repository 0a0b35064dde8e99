#!/bin/bash
# like env_ab.sh for any bench invocation; prints only ms/step: tools/env_ab_cfg.sh VAR "v1 v2" reps <bench flags...>
var=$1; vals=$2; reps=$3; shift 3
for rep in $(seq $reps); do
  for v in $vals; do
    env $var=$v timeout 300 python bench.py --no-extras --no-e2e --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$var=$v', '$*', 'ms', round(d['ms_per_step'],4), 'mhz', d['clocks']['sm_mhz'])"
  done
done
