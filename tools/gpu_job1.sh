#!/usr/bin/env bash
bash tools/gpu_sweep1.sh
python -m pytest tests/test_gpu_robustness.py tests/test_gpu_render.py -x -q -m gpu 2>&1 | tail -15
