#!/bin/bash
# A/B of debug knobs: for each "VAR=val[,VAR=val]" argument (or "base") run the cfg parity tests and a short bench.
out=gpurun_out/ab.log
: > $out
for cfg in "$@"; do
  echo "=== $cfg" >> $out
  envs=""
  if [ "$cfg" != "base" ]; then envs=$(echo $cfg | tr ',' ' '); fi
  env $envs timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "test_cfg1 or test_cfg3 or full_frame" 2>&1 | grep -E "engine 0|frame|passed|failed" | cut -c1-220 >> $out
  env $envs timeout 300 python bench.py --steps 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'shade_ms', d['roofline']['shade_ms_per_step'])" >> $out
done
cat $out
