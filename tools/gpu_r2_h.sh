#!/bin/bash
# Round-2 visit H (2 GPUs): peer-memory probe (symmetric memory / CUDA IPC); reference bytecode archive present?; strided implicit conv tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ls -la oracle/_ref/ | head -5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/probe_symm.py > gpurun_out/probe_symm.log 2>&1; echo "probe rc=$?"; grep -v "^W\|OMP_NUM\|^\*\*\*" gpurun_out/probe_symm.log | cut -c1-250 | head -40
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "small" > gpurun_out/t_small.log 2>&1; echo "small parity rc=$?"; tail -3 gpurun_out/t_small.log
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-500 gpurun_out/bench_ref.json
