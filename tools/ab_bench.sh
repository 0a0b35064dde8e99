#!/bin/bash
# A/B two builds of the library on the same box: tools/ab_bench.sh ab/lib_A.so ab/lib_B.so [phase]
# Alternates the builds so the power-cap clock drift hits both equally.
phase=${3:-gru_unroll_bwd_tc}
for rep in 1 2 3; do
  for v in "$1" "$2"; do
    cp "$v" pymarl_b200/libpymarl_b200.so
    timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],2), round(d['phases_ms']['$phase'],3), d['clocks']['sm_mhz'])"
  done
done
