#!/bin/bash
# A/B two builds of the library on the same box: tools/ab_bench.sh ab/lib_A.so ab/lib_B.so [phase-regex] [reps]
# Alternates the builds so the power-cap clock drift hits both equally.  Leaves the SECOND build installed.
phase=${3:-gru_unroll}
reps=${4:-3}
for rep in $(seq $reps); do
  for v in "$1" "$2"; do
    cp "$v" pymarl_b200/libpymarl_b200.so
    timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys,re; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], {k: round(x,3) for k,x in d['phases_ms'].items() if re.search('$phase', k)})"
  done
done
