timeout 200 python tools/probe_tma.py 2>&1 | tail -6
timeout 400 python -m pytest tests/test_umma_gpu.py tests/test_loftr_gpu.py -x -q 2>&1 | tail -8
echo "=== layers"; IMGS=16 timeout 300 python tools/bench_layers.py q mlp 2>&1 | tail -8
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
