#!/bin/bash
# compute-sanitizer passes over the hot path (run on the GPU box through gpurun; logs land in gpurun_out/).
#   memcheck : smoke() (tiny pair, checked against the golden) + one 480x640 pair through the C ABI
#   racecheck: shared-memory hazards on the tiny pair (the warp-specialised mbarrier/TMEM protocol of umma_gemm.cu)
#   synccheck: barrier misuse on the tiny pair
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
FWD='import torch, sys; sys.path.insert(0, "."); from gim_b200 import LoFTR, get_default_config, load_default_weights; from tests.goldens import load_case, assert_matches_equal
m = LoFTR(get_default_config()); m.load_state_dict(load_default_weights()); m = m.eval().cuda()
data, gold = load_case(CASE); d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in data.items()}; m(d); torch.cuda.synchronize()
print(CASE, "M =", d["b_ids"].numel(), assert_matches_equal(d, gold))'
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $SAN --tool memcheck --print-limit 30 --log-file "$OUT/sanitizer_memcheck_tiny.log" python -c "CASE='tiny_64x96'; $FWD" > "$OUT/sanitizer_memcheck_tiny.out" 2>&1; echo "memcheck tiny rc=$?"
timeout 1500 $SAN --tool memcheck --print-limit 30 --log-file "$OUT/sanitizer_memcheck_480x640.log" python -c "CASE='demo_a_480x640'; $FWD" > "$OUT/sanitizer_memcheck_480x640.out" 2>&1; echo "memcheck 480x640 rc=$?"
timeout 1200 $SAN --tool racecheck --print-limit 30 --log-file "$OUT/sanitizer_racecheck_tiny.log" python -c "CASE='tiny_64x96'; $FWD" > "$OUT/sanitizer_racecheck_tiny.out" 2>&1; echo "racecheck tiny rc=$?"
timeout 900 $SAN --tool synccheck --print-limit 30 --log-file "$OUT/sanitizer_synccheck_tiny.log" python -c "CASE='tiny_64x96'; $FWD" > "$OUT/sanitizer_synccheck_tiny.out" 2>&1; echo "synccheck tiny rc=$?"
tail -n 3 "$OUT"/sanitizer_*.log "$OUT"/sanitizer_*.out
