#!/bin/bash
# Round-2 visit after X (1 GPU, < 80 s; first run under the name gpu_r2_g.sh): the training step (forward + backward + AdamW) captured in ONE CUDA graph — one shot, experimental
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 75 python bench.py --mode train --train-graph --steps 5 --warmup 2 > gpurun_out/g_bench_train_graph.json 2> gpurun_out/g_bench_train_graph.err
echo "graph bench rc=$?"; cat gpurun_out/g_bench_train_graph.json | cut -c1-900; grep -v "^\s*$" gpurun_out/g_bench_train_graph.err | tail -12 | cut -c1-300
