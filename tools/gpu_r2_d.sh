#!/bin/bash
# Round-2 visit D: re-measure the GEMM configurations with the division-free mainloop (all shard shapes + 64x64 latents), step profile.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp mvdfusion_b200/gemm_tuning.json gpurun_out/gemm_tuning_prev.json
timeout 1500 python tools/tune_gemm.py --shards 1,2,4,8 --out gpurun_out/gemm_tuning_r2a.json > gpurun_out/tune_r2a.log 2>&1; echo "tune rc=$?"; tail -2 gpurun_out/tune_r2a.log
timeout 900 python tools/tune_gemm.py --latent 64 --merge gpurun_out/gemm_tuning_r2a.json --out gpurun_out/gemm_tuning_r2.json > gpurun_out/tune_r2b.log 2>&1; echo "tune64 rc=$?"; tail -2 gpurun_out/tune_r2b.log
cp gpurun_out/gemm_tuning_r2.json mvdfusion_b200/gemm_tuning.json
timeout 300 python tools/step_profile.py --out gpurun_out/step_profile_r2.json > gpurun_out/step_profile_r2.txt 2>&1; echo "profile rc=$?"; head -40 gpurun_out/step_profile_r2.txt
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_tuned.json 2> gpurun_out/bench_tuned.err
echo "bench rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/bench_tuned.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['kernels_per_step'], d['roofline']['achieved'])")"
