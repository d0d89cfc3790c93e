#!/bin/bash
# Evidence pass for profiles/: launch list of one bench run + one full ncu capture of each shading kernel.
tag=${1:-prof}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shade_geo -s 3 -c 1 -o $out/${tag}_geo python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_geo.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shade_color -s 3 -c 1 -o $out/${tag}_color python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_color.log 2>&1
cat $out/${tag}_bench.json
