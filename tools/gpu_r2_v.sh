#!/bin/bash
# Round-2 visit V (1 GPU, the round's last ~8 GPU-minutes): training second slice (ABI 15, csrc/train.cu) — the pointwise kernels
# against torch.autograd in float64, the training-gradient parity tests on them, the training bench line (native vs ATen pointwise),
# then as much of the full GPU suite as the remaining budget allows.  Most important first: the call may be cut off at any point.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_training.py -q -m gpu -k pointwise -s > gpurun_out/v_pointwise.log 2>&1
echo "pointwise rc=$?"; grep -E "^pointwise|passed|failed|Error|error" gpurun_out/v_pointwise.log | tail -25
timeout 240 python -m pytest tests/test_training.py -q -m gpu -k "not pointwise" > gpurun_out/v_train.log 2>&1
echo "train rc=$?"; tail -5 gpurun_out/v_train.log
timeout 150 python bench.py --mode train --steps 5 --warmup 2 > gpurun_out/v_bench_train_native.json 2> gpurun_out/v_bench_train_native.err; echo "bench native rc=$?"; cat gpurun_out/v_bench_train_native.json
MVD_TRAIN_ATEN_POINTWISE=1 timeout 150 python bench.py --mode train --steps 5 --warmup 2 > gpurun_out/v_bench_train_aten.json 2> gpurun_out/v_bench_train_aten.err; echo "bench aten rc=$?"; cat gpurun_out/v_bench_train_aten.json
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/v_gpu_tests.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/v_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/v_smoke.log
