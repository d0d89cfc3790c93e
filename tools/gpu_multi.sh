#!/bin/bash
# Multi-GPU pass on N GPUs of one box: the N-GPU == 1-GPU bit-identity check, then the bench at N for configs 2 (= 5: one view per
# GPU, weak scaling) and 4 (one 1024^2 frame in lattice phases, strong scaling) and the reference arm.  Usage: tools/gpu_multi.sh <tag> <N> [noref]
tag=${1:-mg}; N=${2:-2}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 tools/multi_gpu_check.py 256 > $out/${tag}_check.log 2>&1; tail -2 $out/${tag}_check.log
for cfg in 2 4; do
  timeout 400 $TR --master-port 2962$cfg bench.py --gpus $N --steps 10 --warmup 3 --config $cfg > $out/${tag}_n${N}_c${cfg}.json 2> $out/${tag}_n${N}_c${cfg}.err
  python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_n${N}_c${cfg}.json"))
    print("N=$N cfg $cfg", "ms/step %.2f" % d["ms_per_step"], "rays/s %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d["scaling"],
          "per-rank ms", [round(r["ms"], 2) for r in d["per_rank"]], "gather ms", [round(r["all_gather_ms"], 3) for r in d["per_rank"]],
          "shade ms", [round(r["shade_ms"], 2) for r in d["per_rank"]])
except Exception as e:
    print("N=$N cfg $cfg FAILED", e); print(open("$out/${tag}_n${N}_c${cfg}.err").read()[-2000:])
PY
done
[ "$3" = noref ] || timeout 400 $TR --master-port 29630 bench.py --gpus $N --steps 5 --warmup 1 --impl reference > $out/${tag}_n${N}_ref.json 2> $out/${tag}_n${N}_ref.err; cat $out/${tag}_n${N}_ref.json | cut -c1-300
