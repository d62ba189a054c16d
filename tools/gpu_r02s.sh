#!/bin/bash
timeout 400 python -m pytest tests/test_api_app.py -m gpu -x -q 2>&1 | tail -8
