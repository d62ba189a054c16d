#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_train.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5
