#!/bin/bash
# scratch script for the A/B experiment of the day
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:attn_fwd" -s 2 -c 1 -o gpurun_out/prof_attn_new -f python tools/ncu_attn.py > gpurun_out/ncu_attn_new.log 2>&1
tail -3 gpurun_out/ncu_attn_new.log
