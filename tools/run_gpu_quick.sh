#!/bin/bash
# scratch script for the A/B experiment of the day
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -s 2>&1 | tail -45 | tee gpurun_out/test_fullsize.log
