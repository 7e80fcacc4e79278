#!/usr/bin/env bash
# development aid: A/B of library builds (computer-graphics-engine_b200/libcge_<name>.so, built with CGE_NVCC_EXTRA / CGE_OUT) on the
# shadow pass: C5 whole frame, a 1/8 share, C3, and the point-light configs; frames compared by hash.
# usage: LIBS="libcge_a.so libcge_b.so" tools/ab_libs.sh
for L in ${LIBS:-libcge.so}; do
  export CGE_LIB=$PWD/computer-graphics-engine_b200/$L; echo "== $L"
  export CGE_BANDS=1
  python tools/sweep_vis.py c5_dragon "X=0" | cut -c1-260
  SWEEP_PART=8 python tools/sweep_vis.py c5_dragon "X=0" | cut -c1-260
  python tools/sweep_vis.py c3_teapot_soft "X=0" | cut -c1-260
  unset CGE_BANDS
  python tools/quick_bench.py c1_cornell c2_cube_textured c4_monkey_mirror 2>&1 | grep "fast-default" | cut -c1-110
  true
done
