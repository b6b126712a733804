#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_full.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/prof_full_mixed.py 2>&1 | tail -32
