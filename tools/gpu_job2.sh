#!/usr/bin/env bash
V="CGE_PACKET=0 CGE_PACKET=16,CGE_PACKET_BUDGET=24 CGE_PACKET=16,CGE_PACKET_BUDGET=48 CGE_PACKET=16,CGE_PACKET_BUDGET=96 CGE_PACKET=16,CGE_PACKET_BUDGET=48,CGE_PACKET_FAT_PCT=400 CGE_PACKET=8,CGE_PACKET_BUDGET=24,CGE_PACKET_FAT_PCT=100 CGE_PACKET=8,CGE_PACKET_BUDGET=48 CGE_PACKET=4,CGE_PACKET_BUDGET=32"
export CGE_BANDS=1
python tools/sweep_vis.py c5_dragon $V 2>&1 | tee gpurun_out/sweep_c5.log
SWEEP_PART=8 python tools/sweep_vis.py c5_dragon $V 2>&1 | tee gpurun_out/sweep_c5_p8.log
python tools/sweep_vis.py c3_teapot_soft $V 2>&1 | tee gpurun_out/sweep_c3.log
export CGE_LIB=$PWD/computer-graphics-engine_b200/libcge_stats.so
S="CGE_PACKET=16,CGE_PACKET_BUDGET=24 CGE_PACKET=16,CGE_PACKET_BUDGET=48 CGE_PACKET=16,CGE_PACKET_BUDGET=96 CGE_PACKET=8,CGE_PACKET_BUDGET=48"
python tools/packet_stats.py c5_dragon:0.5 $S; python tools/packet_stats.py c3_teapot_soft $S
