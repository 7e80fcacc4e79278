export CGE_BANDS=1
python tools/sweep_vis.py c5_dragon "CGE_VIS_CULL=0" "CGE_VIS_CULL=1,CGE_CULL_SEED=0" "CGE_VIS_CULL=1,CGE_CULL_SEED=1" "CGE_VIS_CULL=1,CGE_CULL_BUDGET=32" | cut -c1-300
SWEEP_PART=8 python tools/sweep_vis.py c5_dragon "CGE_VIS_CULL=0" "CGE_VIS_CULL=1,CGE_CULL_SEED=0" "CGE_VIS_CULL=1,CGE_CULL_SEED=1" | cut -c1-300
python tools/sweep_vis.py c3_teapot_soft "CGE_VIS_CULL=0" "CGE_VIS_CULL=1,CGE_CULL_SEED=0" "CGE_VIS_CULL=1,CGE_CULL_SEED=1" | cut -c1-300
unset CGE_BANDS
python tools/quick_bench.py c1_cornell c2_cube_textured c4_monkey_mirror 2>&1 | grep -v reference | cut -c1-200
CGE_LIB=$PWD/computer-graphics-engine_b200/libcge_base.so python tools/quick_bench.py c1_cornell c2_cube_textured c4_monkey_mirror 2>&1 | grep -v reference | cut -c1-200
