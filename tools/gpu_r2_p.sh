#!/bin/bash
# Round-2 visit P (1 GPU): where the folded-LayerNorm consumers lose their time — graph-timed A/B of the four GEMMs, then one ncu --set full
# capture of each (source view), then the kernel tests and the graph-timed "+st" / "+ln" shapes of the step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tools/ln_fold_probe.py --time > gpurun_out/ln_probe_time.txt 2>&1; cat gpurun_out/ln_probe_time.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 4 -f -o gpurun_out/prof_ln_probe python tools/ln_fold_probe.py > gpurun_out/ncu_ln_probe.log 2>&1; echo "ncu rc=$?"
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "layernorm or qkv_and_attention or geglu or fp16_copy" > gpurun_out/t_ln.log 2>&1
echo "ln-tests rc=$?"; tail -3 gpurun_out/t_ln.log
cp mvdfusion_b200/gemm_tuning.json gpurun_out/t3.json
timeout 500 python tools/tune_gemm.py --only "+ln" --merge gpurun_out/t3.json --out gpurun_out/t6.json > gpurun_out/tune_v8.log 2>&1; echo "tune rc=$?"; grep -v "^----" gpurun_out/tune_v8.log
