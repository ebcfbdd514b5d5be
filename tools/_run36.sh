timeout 200 python tools/probe_tma.py 2>&1 | tail -9
echo "=== layers auto"; IMGS=16 timeout 300 python tools/bench_layers.py l1c 2>&1 | tail -4
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
