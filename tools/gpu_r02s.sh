#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_render.py tests/test_gpu_comm.py tests/test_gpu_distributed.py tests/test_api_app.py tests/test_gpu_modes.py tests/test_gpu_pathtracing.py -m gpu -x -q 2>&1 | tail -4
