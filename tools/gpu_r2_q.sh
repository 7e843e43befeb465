#!/bin/bash
# Round-2 visit Q (1 GPU): folded-LayerNorm consumers with the helper warp staging the unit's column sums / bias — tests, probe timing, shapes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "layernorm or qkv_and_attention or geglu or fp16_copy" > gpurun_out/t_ln.log 2>&1
echo "ln-tests rc=$?"; tail -3 gpurun_out/t_ln.log
timeout 120 python tools/ln_fold_probe.py --time > gpurun_out/ln_probe_time2.txt 2>&1; cat gpurun_out/ln_probe_time2.txt
cp mvdfusion_b200/gemm_tuning.json gpurun_out/t3.json
timeout 500 python tools/tune_gemm.py --only "+" --merge gpurun_out/t3.json --out gpurun_out/t7.json > gpurun_out/tune_v9.log 2>&1; echo "tune rc=$?"; grep -v "^----" gpurun_out/tune_v9.log
cp gpurun_out/t7.json mvdfusion_b200/gemm_tuning.json
timeout 200 python bench.py --no-cpu-baseline --reps 3 > gpurun_out/bench_lnfold5.json 2> gpurun_out/bench_lnfold5.err; echo "bench fold (tuned) rc=$?"
MVD_NO_LN_FOLD=1 timeout 200 python bench.py --no-cpu-baseline --reps 3 > gpurun_out/bench_lnpass4.json 2> gpurun_out/bench_lnpass4.err; echo "bench pass rc=$?"
python - <<'PY'
import json
for n in ("lnfold5", "lnpass4"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"], 2), round(d["ms_per_step"], 4), d["kernels_per_step"], round(d["e2e"]["value"], 2), d["roofline"]["achieved"], d["roofline"]["frac"],
              [(k["kernel"], k["calls"], round(k["ms"], 3)) for k in d["kernels"][:5]])
    except Exception as e:
        print(n, "failed", e)
PY
