#!/usr/bin/env bash
# ncu evidence for the round (R = file prefix): the shadow-ray kernel and the light-hull pre-pass on C5 (whole frame, one pipeline), the
# chain and shading kernels, the same shadow-ray kernel without the pre-pass, and the launch list of bench.py
R=${R:-r04}
set -x
python tools/prof_render.py c5_dragon 3 > gpurun_out/plain.log 2>&1 || exit 1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:wf_vis_regroup_kernel -s 5 -c 1 -o gpurun_out/${R}_wf_vis_regroup_c5 python tools/prof_render.py c5_dragon 3 > gpurun_out/ncu1.log 2>&1
$NCU -k regex:wf_vis_cull_kernel -s 2 -c 1 -o gpurun_out/${R}_wf_vis_cull_c5 python tools/prof_render.py c5_dragon 3 > gpurun_out/ncu1b.log 2>&1
CGE_VIS_CULL=0 $NCU -k regex:wf_vis_regroup_kernel -s 5 -c 1 -o gpurun_out/${R}_wf_vis_regroup_nocull_c5 python tools/prof_render.py c5_dragon 3 > gpurun_out/ncu1c.log 2>&1
$NCU -k regex:wf_chain_kernel -s 2 -c 1 -o gpurun_out/${R}_wf_chain_c5 python tools/prof_render.py c5_dragon 3 > gpurun_out/ncu2.log 2>&1
$NCU -k regex:wf_shade_kernel -s 2 -c 1 -o gpurun_out/${R}_wf_shade_c5 python tools/prof_render.py c5_dragon 3 > gpurun_out/ncu3.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_launches_bench_c5_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu4.log 2>&1
tail -n 2 gpurun_out/ncu1.log; tail -n 2 gpurun_out/ncu4.log
# summaries travel back, the reports (18 MB each) stay on the box except the dominant kernel's
for f in gpurun_out/${R}_*.ncu-rep; do python tools/ncu_summary.py $f > ${f%.ncu-rep}.txt; done
ncu -i gpurun_out/${R}_wf_vis_regroup_c5.ncu-rep --page source --csv > gpurun_out/${R}_wf_vis_regroup_c5_source.csv 2>/dev/null
ls gpurun_out/${R}_*.ncu-rep | grep -v "${R}_wf_vis_regroup_c5.ncu-rep" | xargs rm -f
