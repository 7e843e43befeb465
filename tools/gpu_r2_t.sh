#!/bin/bash
# Round-2 visit T (1 GPU): final validation of build v6 with its defaults (LayerNorm fold up to 4096 rows): whole GPU suite, smoke, bench of
# BASELINE configs[1] (with the CPU reference arm) and configs[4] on one GPU, one rank of the 8-way sharded step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests_v6b.log 2>&1; echo "all-tests rc=$?"; tail -2 gpurun_out/gpu_tests_v6b.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_v6b.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_v6b.log
timeout 400 python bench.py > gpurun_out/bench_v6b.json 2> gpurun_out/bench_v6b.err; echo "bench rc=$?"
timeout 300 python bench.py --latent 64 --no-cpu-baseline --reps 3 > gpurun_out/bench_s64_v6b.json 2> gpurun_out/bench_s64_v6b.err; echo "bench s64 rc=$?"
MVD_NO_LN_FOLD=1 timeout 200 python bench.py --no-cpu-baseline --reps 3 > gpurun_out/bench_v6b_lnpass.json 2> gpurun_out/bench_v6b_lnpass.err; echo "bench pass rc=$?"
python - <<'PY'
import json
for n in ("v6b", "s64_v6b", "v6b_lnpass"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"], 2), round(d["ms_per_step"], 4), d["kernels_per_step"], round(d["e2e"]["value"], 2), d["roofline"]["achieved"], d["roofline"]["frac"], d["step_roofline"])
    except Exception as e:
        print(n, "failed", e)
PY
timeout 200 python tools/step_profile.py --world 8 --reps 7 > gpurun_out/step_profile_w8_v6b.txt 2>&1; head -1 gpurun_out/step_profile_w8_v6b.txt
timeout 200 python tools/step_profile.py --reps 7 > gpurun_out/step_profile_v6b.txt 2>&1; head -2 gpurun_out/step_profile_v6b.txt
