#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_full.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python tools/prof_full.py 1048576 5; timeout 200 python tools/prof_full.py 1048576 5 shuffle; SP_NNUE_FT=warp timeout 200 python tools/prof_full.py 1048576 5 shuffle
