#!/usr/bin/env bash
# development aid: A/B of library builds (computer-graphics-engine_b200/libcge*.so) on the shadow pass, frames compared bit for bit
export CGE_BANDS=1
python - <<'PY'
import importlib, os, sys, json, glob, subprocess
PY
for L in ${LIBS:-libcge_H.so libcge_df69fcf.so}; do
  export CGE_LIB=$PWD/computer-graphics-engine_b200/$L; echo $L
  python tools/sweep_vis.py c5_dragon "X=0" | cut -c1-200
  SWEEP_PART=8 python tools/sweep_vis.py c5_dragon "X=0" | cut -c1-200
  python tools/sweep_vis.py c3_teapot_soft "X=0" | cut -c1-200
  true
done
