#!/usr/bin/env bash
# ncu evidence for the round: the dominant kernel on C5 (whole frame, one pipeline), the chain kernel, and the launch list of bench.py
set -x
python tools/prof_render.py c5_dragon 3 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wf_vis_regroup_kernel -s 5 -c 1 -o gpurun_out/r03_wf_vis_regroup_c5 python tools/prof_render.py c5_dragon 3 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_chain_kernel -s 2 -c 1 -o gpurun_out/r03_wf_chain_c5 python tools/prof_render.py c5_dragon 3 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_shade_kernel -s 2 -c 1 -o gpurun_out/r03_wf_shade_c5 python tools/prof_render.py c5_dragon 3 > gpurun_out/ncu3.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r03_launches_bench_c5_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu4.log 2>&1
tail -2 gpurun_out/ncu1.log gpurun_out/ncu4.log
