#!/bin/bash
# Round-2 visit W (1 GPU, the last ~4 GPU-minutes): roofline of the training-side streaming kernels (csrc/train.cu) and a last bench
# line of the committed build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 100 python tools/train_kernels_bench.py > gpurun_out/w_train_kernels_bench.json 2> gpurun_out/w_train_kernels_bench.err; echo "train kernels rc=$?"; cat gpurun_out/w_train_kernels_bench.err | grep kernel | cut -c1-230
timeout 150 python bench.py --no-cpu-baseline --reps 3 > gpurun_out/w_bench_v8.json 2> gpurun_out/w_bench_v8.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/w_bench_v8.json
