#!/bin/bash
# Round-2 visit L (1 GPU): LayerNorm carried by its neighbours' epilogues (ABI 13) — kernel tests, A/B bench against the ln_kernel
# program (MVD_NO_LN_FOLD=1), then the whole GPU suite and the smoke step on the new default.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "layernorm or qkv_and_attention or geglu or fp16_copy" > gpurun_out/t_ln.log 2>&1
echo "ln-tests rc=$?"; tail -4 gpurun_out/t_ln.log
timeout 300 python bench.py --no-cpu-baseline --reps 3 > gpurun_out/bench_lnfold.json 2> gpurun_out/bench_lnfold.err; echo "bench fold rc=$?"
MVD_NO_LN_FOLD=1 timeout 300 python bench.py --no-cpu-baseline --reps 3 > gpurun_out/bench_lnpass.json 2> gpurun_out/bench_lnpass.err; echo "bench pass rc=$?"
python - <<'PY'
import json
for n in ("lnfold", "lnpass"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"], 2), round(d["ms_per_step"], 4), d["kernels_per_step"], round(d["e2e"]["value"], 2), d["roofline"]["achieved"], d["roofline"]["frac"],
              [(k["kernel"], k["calls"], round(k["ms"], 3)) for k in d["kernels"][:6]])
    except Exception as e:
        print(n, "failed", e)
PY
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests_v6.log 2>&1; echo "all-tests rc=$?"; tail -6 gpurun_out/gpu_tests_v6.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_v6.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_v6.log
