#!/bin/bash
for v in 1 2; do echo "== variant $v"; VNR_TRAIN_VARIANT=$v timeout 120 python tools/exp_dp_vs_accum.py 2>&1 | tail -5; done
