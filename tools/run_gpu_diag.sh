#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_determinism.py 2>&1 | tail -60 | tee gpurun_out/diag_determinism.log
timeout 300 python tools/diag_engine.py 2>&1 | tail -40 | tee gpurun_out/diag_engine.log
