#!/bin/bash
# A/B an environment switch of the library on the same box: tools/env_ab.sh VAR "v1 v2 ..." [reps] [bench flags...]
# Alternates the values so the power-cap clock drift hits all of them equally; prints ms/step, clocks and the phases.
var=$1; vals=$2; reps=${3:-2}; shift 3
for rep in $(seq $reps); do
  for v in $vals; do
    env $var=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-e2e --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$var=$v', 'ms', round(d['ms_per_step'],3), 'mhz', d['clocks']['sm_mhz'], {k: round(x,2) for k,x in d['phases_ms'].items() if x > 0.3})"
  done
done
