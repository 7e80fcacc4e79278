#!/usr/bin/env bash
export CGE_BANDS=1 CGE_PACKET=0
for L in libcge.so libcge_far.so; do
  export CGE_LIB=$PWD/computer-graphics-engine_b200/$L; echo $L
  python tools/sweep_vis.py c5_dragon "CGE_PACKET=0"
  SWEEP_PART=8 python tools/sweep_vis.py c5_dragon "CGE_PACKET=0"
  python tools/sweep_vis.py c3_teapot_soft "CGE_PACKET=0"
done
