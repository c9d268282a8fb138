#!/bin/bash
# Two-GPU visit: the torchrun launch the driver uses, the sharded cfg4 record over NCCL, the reference arm under torchrun.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | tee gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
cut -c1-300 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
cut -c1-300 gpurun_out/bench_ref_n2.json; tail -2 gpurun_out/bench_ref_n2.err | cut -c1-200
