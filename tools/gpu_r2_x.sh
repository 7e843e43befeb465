#!/bin/bash
# Round-2 visit X / Y (1 GPU, ~1 minute each): the ABI-16 bilinear gather kernels, then the ABI-17 two-launch GroupNorm, on hardware — pointwise + training-gradient tests, micro-benchmark
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_training.py tests/test_lib_abi.py -q -m gpu -s > gpurun_out/x_train.log 2>&1
echo "train tests rc=$?"; grep -E "^pointwise gather|passed|failed|Error" gpurun_out/x_train.log | tail -8
timeout 60 python tools/train_kernels_bench.py > gpurun_out/x_train_kernels_bench.json 2> gpurun_out/x_train_kernels_bench.err; echo "bench rc=$?"; grep -E "gather|groupnorm" gpurun_out/x_train_kernels_bench.err | cut -c1-200
