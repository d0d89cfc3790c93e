#!/bin/bash
# One GPU-box pass: tensor-core primitive self-tests, parity tests, smoke, bench.  Everything under `timeout` so that a hung
# kernel cannot hold the box.  Usage: tools/gpu_check.sh <tag> [extra bench args]
tag=${1:-run}
shift
out=gpurun_out
mkdir -p $out
timeout 120 python -m pytest tests/test_gpu_umma.py -m gpu -x -q 2>&1 | tail -8 > $out/${tag}_umma.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_umma.py -s 2>&1 | tail -150 > $out/${tag}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
if ! grep -q " passed" $out/${tag}_tests.log || grep -q "failed" $out/${tag}_tests.log; then
  echo "parity tests did not pass" > $out/${tag}_bench.err; tail -60 $out/${tag}_tests.log
fi
timeout 300 python bench.py --no-cpu-baseline "$@" > $out/${tag}_bench.json 2>> $out/${tag}_bench.err
cat $out/${tag}_umma.log; tail -8 $out/${tag}_tests.log; tail -3 $out/${tag}_smoke.log; cat $out/${tag}_bench.json
