#!/bin/bash
# check -> chunk-size A/B (bench only) -> evidence profile; stops at the first failure.  Usage: tools/gpu_round.sh <tag>
tag=${1:-r}
tools/gpu_check.sh $tag || exit 1
for cs in 4194304 33554432; do
  KPN_CHUNK_SAMPLES=$cs timeout 120 python bench.py --steps 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk $cs ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'])" >> gpurun_out/${tag}_chunks.log
done
cat gpurun_out/${tag}_chunks.log
timeout 400 tools/gpu_profile.sh ${tag}p > /dev/null 2>&1
