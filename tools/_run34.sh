timeout 400 python -m pytest tests/test_umma_gpu.py tests/test_loftr_gpu.py -x -q 2>&1 | tail -12
echo "=== layers auto"; IMGS=16 timeout 300 python tools/bench_layers.py 2>&1 | tail -26
echo "=== layers bk32"; GIMB_BK=32 IMGS=16 timeout 300 python tools/bench_layers.py c2 l2c1 l1c c1 mlp 2>&1 | tail -16
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
