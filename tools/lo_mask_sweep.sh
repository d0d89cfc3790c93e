#!/bin/bash
# Precision / speed of the geometry kernel as a function of which stages get the W_lo (bits 0..5) and A_lo (bits 6,7) passes.
# usage: tools/lo_mask_sweep.sh 0xFF 0xFE ...   (environment knob KPN_LO_MASK of the library)
out=gpurun_out/lo_sweep.log
: > $out
for m in "$@"; do
  echo "=== KPN_LO_MASK=$m" >> $out
  KPN_LO_MASK=$m timeout 300 python -m pytest tests/test_gpu_vseq.py -m gpu -q -s 2>&1 | grep -E "vseq|passed|failed" | cut -c1-160 >> $out
  KPN_LO_MASK=$m timeout 300 python bench.py --steps 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step %.2f geo %.2f' % (d['ms_per_step'], d['roofline']['geo_ms_per_step']))" >> $out
done
cat $out
