timeout 400 python -m pytest tests/test_umma_gpu.py -x -q 2>&1 | tail -15
echo "=== layers TMA"; IMGS=16 timeout 300 python tools/bench_layers.py 2>&1 | tail -30
echo "=== layers legacy"; GIMB_EPI=legacy IMGS=16 timeout 300 python tools/bench_layers.py c3 c1 "l1.0" q mlp 2>&1 | tail -30
timeout 600 python -m pytest tests/test_loftr_gpu.py -x -q 2>&1 | tail -8
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('TMA', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
