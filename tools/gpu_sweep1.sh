#!/usr/bin/env bash
# development aid: A/B of the packet shadow pass against the per-ray pass (run under gpurun)
V="CGE_PACKET=0 CGE_PACKET=16,CGE_PACKET_FAT_PCT=25 CGE_PACKET=16,CGE_PACKET_FAT_PCT=50 CGE_PACKET=16,CGE_PACKET_FAT_PCT=100 CGE_PACKET=16,CGE_PACKET_FAT_PCT=200 CGE_PACKET=16,CGE_PACKET_FAT_PCT=400 CGE_PACKET=8,CGE_PACKET_FAT_PCT=100 CGE_PACKET=8,CGE_PACKET_FAT_PCT=200 CGE_PACKET=4,CGE_PACKET_FAT_PCT=100"
python tools/sweep_vis.py c5_dragon $V > gpurun_out/sweep_c5.log 2>&1
SWEEP_PART=8 python tools/sweep_vis.py c5_dragon $V > gpurun_out/sweep_c5_p8.log 2>&1
python tools/sweep_vis.py c3_teapot_soft $V > gpurun_out/sweep_c3.log 2>&1
cat gpurun_out/sweep_c5.log gpurun_out/sweep_c5_p8.log gpurun_out/sweep_c3.log
