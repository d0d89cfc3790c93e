#!/bin/bash
# Bench lines of the BASELINE configs other than the default one, on ONE GPU: config 3 (64+64, ERT), config 4 (1024^2 frame),
# the 24-keypoint (reference default n_kpt) and silhouette-scene variants of config 2.  Usage: tools/gpu_configs.sh <tag>
tag=${1:-cfg}
out=gpurun_out
mkdir -p $out
for spec in "c3:--config 3" "c4:--config 4" "k24:--config 2 --n-kpt 24" "hull:--config 2 --scene hull" "c3hull:--config 3 --scene hull"; do
  name=${spec%%:*}; args=${spec#*:}
  timeout 300 python bench.py --no-cpu-baseline --steps 5 $args > $out/${tag}_${name}.json 2> $out/${tag}_${name}.err
  python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_${name}.json")); r=d["roofline"]
    print("$name", "ms/step %.2f" % d["ms_per_step"], "rays/s %.3g" % d["value"], "e2e %.3g" % d["e2e"]["value"], "geo_ms %.2f" % r["geo_ms_per_step"],
          "pair_ms %.2f" % r["pair"]["ms_per_step"], "valid %.3f" % r["valid_frac"], "frac %.3f" % (r["frac"] or 0), "launches", d["gpu_launches"])
except Exception as e:
    print("$name FAILED", e); print(open("$out/${tag}_${name}.err").read()[-1500:])
PY
done
