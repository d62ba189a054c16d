#!/bin/bash
REPS=150 timeout 800 python tools/exp_var2_fp32.py 2>&1 | tail -24
