#!/bin/bash
# Round-2 visit U (1 GPU): nearest x2 upsample folded into the convolution (ABI 14, conv_up2) — kernel tests, tile configurations of the
# three "+up" shapes, A/B bench against the upsample-pass program (MVD_NO_FOLD_UP=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "upsample or conv3x3 or layernorm or plain_bias or splitk" > gpurun_out/t_up.log 2>&1
echo "up-tests rc=$?"; tail -4 gpurun_out/t_up.log
cp mvdfusion_b200/gemm_tuning.json gpurun_out/t7.json
timeout 500 python tools/tune_gemm.py --only "+up" --merge gpurun_out/t7.json --out gpurun_out/t8.json > gpurun_out/tune_v10.log 2>&1; echo "tune rc=$?"; grep -v "^----" gpurun_out/tune_v10.log
cp gpurun_out/t8.json mvdfusion_b200/gemm_tuning.json
timeout 200 python bench.py --no-cpu-baseline --reps 3 > gpurun_out/bench_upfold.json 2> gpurun_out/bench_upfold.err; echo "bench fold rc=$?"
MVD_NO_FOLD_UP=1 timeout 200 python bench.py --no-cpu-baseline --reps 3 > gpurun_out/bench_uppass.json 2> gpurun_out/bench_uppass.err; echo "bench pass rc=$?"
python - <<'PY'
import json
for n in ("upfold", "uppass"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"], 2), round(d["ms_per_step"], 4), d["kernels_per_step"], round(d["e2e"]["value"], 2), d["roofline"]["achieved"], d["roofline"]["frac"], d["step_roofline"])
    except Exception as e:
        print(n, "failed", e)
PY
timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "full_size_N2_vs_oracle or unet_small or ddim4 or S64_vs_oracle" > gpurun_out/t_up_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/t_up_parity.log
