for dbg in 0 1 2 4 7; do
PMB_FC1_DBG=$dbg timeout 200 python bench.py --precision bf16 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline --batch 1400 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dbg',$dbg, d['phases_ms']['fc1_fwd_both_tc'])"
done
timeout 300 python -m pytest tests -m gpu -x -q -k "bf16" 2>&1 | tail -3
