#!/bin/bash
# Round-2 visit I (1 GPU): new kernel tests (strided implicit conv, masked attention, fp32 LayerNorm), CLIP encoder parity, the whole suite, bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_clip.py -x -q -m gpu > gpurun_out/t_new.log 2>&1; echo "new rc=$?"; tail -6 gpurun_out/t_new.log
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "gpu-tests rc=$?"; tail -3 gpurun_out/gpu_tests.log
timeout 600 python bench.py --kernel-table gpurun_out/kernels.json > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["kernels_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["cpu_baseline"], d["e2e"]["value"])
PY
timeout 300 python tools/step_profile.py --out gpurun_out/step_profile_r2b.json > gpurun_out/step_profile_r2b.txt 2>&1; head -3 gpurun_out/step_profile_r2b.txt
