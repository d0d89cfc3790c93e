#!/bin/bash
# One GPU-box pass: tensor-core primitive self-tests, parity tests, bench.  Everything under `timeout` so that a hung
# kernel cannot hold the box.  Usage: tools/gpu_check.sh <tag>
tag=${1:-run}
out=gpurun_out
mkdir -p $out
timeout 120 python -m pytest tests/test_gpu_umma.py -m gpu -x -q 2>&1 | tail -8 > $out/${tag}_umma.log
timeout 180 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_umma.py -s 2>&1 | tail -40 > $out/${tag}_tests.log
if ! grep -q " passed" $out/${tag}_tests.log || grep -q "failed" $out/${tag}_tests.log; then
  echo "parity tests did not pass: skipping the bench" > $out/${tag}_bench.err; tail -30 $out/${tag}_tests.log; exit 1
fi
timeout 180 python bench.py --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err
cat $out/${tag}_umma.log; tail -5 $out/${tag}_tests.log; cat $out/${tag}_bench.json
