#!/bin/bash
# Evidence pass for profiles/: launch list of one bench run, DRAM bytes of every kernel of one frame, one full ncu capture of
# each shading kernel.  Usage: tools/gpu_profile.sh <tag>
tag=${1:-prof}
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_bench.log 2>&1
# one frame = front, shade_geo, colour_list, shade_color, composite: skip the 3 warm-up frames (15 matching launches)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'front_kernel|shade_|colour_list|composite|pack_nhwc' -s 27 -c 9 --csv --log-file $out/${tag}_frame_dram.csv $B > $out/${tag}_ncu_dram.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shade_geo -s 3 -c 1 -o $out/${tag}_geo $B > $out/${tag}_ncu_geo.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shade_color -s 3 -c 1 -o $out/${tag}_color $B > $out/${tag}_ncu_color.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:front_kernel -s 3 -c 1 -o $out/${tag}_front $B > $out/${tag}_ncu_front.log 2>&1
cat $out/${tag}_bench.json
