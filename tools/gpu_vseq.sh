#!/bin/bash
# view-sequential geometry kernel: parity tests, engine 4 vs engine 0 bench, stage timing of the instrumented build (if built)
out=gpurun_out; mkdir -p $out; tag=${1:-v}
timeout 300 python -m pytest tests/test_gpu_vseq.py -m gpu -q -x -s 2>&1 | tail -25 > $out/${tag}_vseq_tests.log; cat $out/${tag}_vseq_tests.log
for e in ${ENGINES:-4 0}; do timeout 200 python bench.py --no-cpu-baseline --engine $e > $out/${tag}_bench_e$e.json 2> $out/${tag}_bench_e$e.err; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_bench_e$e.json")); r=d["roofline"]; print("engine $e", "ms/step %.2f" % d["ms_per_step"], "geo %.2f" % r["geo_ms_per_step"], "pair %.2f" % r["pair"]["ms_per_step"], "frac %.3f" % r["frac"])
except Exception as ex: print("engine $e FAILED", ex); print(open("$out/${tag}_bench_e$e.err").read()[-1500:])
PY
done
if [ -f keypointnerf_b200/lib/libkpnerf_b200_timing.so ]; then timeout 200 python tools/stage_times.py --vseq > $out/${tag}_stage.txt 2>&1; cat $out/${tag}_stage.txt; fi
