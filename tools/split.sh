export CGE_BANDS=1
SWEEP_PART=8 python tools/sweep_vis.py c5_dragon CGE_CHAIN_SPLIT=1 CGE_CHAIN_SPLIT=0 CGE_CHAIN_SPLIT=1 CGE_CHAIN_SPLIT=0 | cut -c1-200
SWEEP_PART=16 python tools/sweep_vis.py c5_dragon CGE_CHAIN_SPLIT=1 CGE_CHAIN_SPLIT=0 | cut -c1-200
python tools/sweep_vis.py c3_teapot_soft:0.7 CGE_CHAIN_SPLIT=1 CGE_CHAIN_SPLIT=0 | cut -c1-200
