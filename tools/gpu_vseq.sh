#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=${1:-v}
timeout 300 python -m pytest tests/test_gpu_vseq.py -m gpu -q -x -s 2>&1 | tail -25 > $out/${tag}_vseq_tests.log; cat $out/${tag}_vseq_tests.log
for e in 0 3; do timeout 200 python bench.py --no-cpu-baseline --engine $e > $out/${tag}_bench_e$e.json 2> $out/${tag}_bench_e$e.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_bench_e$e.json")); r=d["roofline"]; print("engine $e", "ms/step %.2f" % d["ms_per_step"], "geo %.2f" % r["geo_ms_per_step"], "pair %.2f" % r["pair"]["ms_per_step"], "frac %.3f" % r["frac"])
except Exception as ex: print("engine $e FAILED", ex); print(open("$out/${tag}_bench_e$e.err").read()[-1500:])
PY
done
