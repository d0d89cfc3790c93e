#!/bin/bash
# Precision / speed of the geometry kernel as a function of which stages get the W_lo (bits 0..5) and A_lo (bits 6,7) passes.
out=gpurun_out/lo_sweep.log
: > $out
for m in "$@"; do
  echo "=== KPN_LO_MASK=$m" >> $out
  KPN_LO_MASK=$m timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "test_tile_coarse or test_cfg1 or test_cfg3 or test_tile_fine_with" 2>&1 | grep -E "engine 0|passed|failed" | cut -c1-200 >> $out
  KPN_LO_MASK=$m timeout 300 python bench.py --steps 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'])" >> $out
done
cat $out
