#!/bin/bash
for i in 1 2 3; do timeout 200 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s -k psnr 2>&1 | grep -i "volume PSNR\|passed\|failed\|assert" | head -5; done
