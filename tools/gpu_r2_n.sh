#!/bin/bash
# Round-2 visit N (1 GPU): LayerNorm fold v2 (statistics reduced by a helper warp one unit ahead) — kernel tests, tile configurations of
# the "+st" producers / "+ln" consumers measured, A/B bench against the ln_kernel program, per-shape table.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "layernorm or qkv_and_attention or geglu or fp16_copy" > gpurun_out/t_ln.log 2>&1
echo "ln-tests rc=$?"; tail -3 gpurun_out/t_ln.log
timeout 200 python bench.py --no-cpu-baseline --reps 3 > gpurun_out/bench_lnfold2.json 2> gpurun_out/bench_lnfold2.err; echo "bench fold (untuned) rc=$?"
cp mvdfusion_b200/gemm_tuning.json gpurun_out/t3.json
timeout 500 python tools/tune_gemm.py --only "+" --merge gpurun_out/t3.json --out gpurun_out/t4.json > gpurun_out/tune_v6.log 2>&1; echo "tune rc=$?"; tail -1 gpurun_out/tune_v6.log
cp gpurun_out/t4.json mvdfusion_b200/gemm_tuning.json
timeout 200 python bench.py --no-cpu-baseline --reps 3 > gpurun_out/bench_lnfold3.json 2> gpurun_out/bench_lnfold3.err; echo "bench fold (tuned) rc=$?"
MVD_NO_LN_FOLD=1 timeout 200 python bench.py --no-cpu-baseline --reps 3 > gpurun_out/bench_lnpass2.json 2> gpurun_out/bench_lnpass2.err; echo "bench pass rc=$?"
python - <<'PY'
import json
for n in ("lnfold2", "lnfold3", "lnpass2"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"], 2), round(d["ms_per_step"], 4), d["kernels_per_step"], round(d["e2e"]["value"], 2), d["roofline"]["achieved"], d["roofline"]["frac"],
              [(k["kernel"], k["calls"], round(k["ms"], 3)) for k in d["kernels"][:5]])
    except Exception as e:
        print(n, "failed", e)
PY
timeout 200 python tools/step_profile.py --reps 7 > gpurun_out/step_profile_lnfold2.txt 2>&1; echo "profile rc=$?"; head -2 gpurun_out/step_profile_lnfold2.txt
