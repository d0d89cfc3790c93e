#!/bin/bash
# One GPU-box pass: tensor-core primitive self-tests, parity tests, bench.  Everything under `timeout` so that a hung
# kernel cannot hold the box.  Usage: tools/gpu_check.sh <tag>
tag=${1:-run}
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_umma.py -m gpu -x -q 2>&1 | tail -8 > $out/${tag}_umma.log
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_umma.py -s 2>&1 | tail -40 > $out/${tag}_tests.log
if ! grep -q " passed" $out/${tag}_tests.log || grep -q "failed" $out/${tag}_tests.log; then
  KPN_ISSUE_BRANCH=1 timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_umma.py -s 2>&1 | tail -40 > $out/${tag}_tests_branch.log
fi
timeout 600 python bench.py --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err
cat $out/${tag}_umma.log; tail -5 $out/${tag}_tests.log; cat $out/${tag}_bench.json
