export CGE_BANDS=1
for L in libcge.so libcge_B.so libcge_C.so libcge_D.so; do
  export CGE_LIB=$PWD/computer-graphics-engine_b200/$L; echo $L
  python tools/sweep_vis.py c5_dragon "X=0" | cut -c1-200
  SWEEP_PART=8 python tools/sweep_vis.py c5_dragon "X=0" | cut -c1-200
  python tools/sweep_vis.py c3_teapot_soft "X=0" | cut -c1-200
done
