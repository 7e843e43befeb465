#!/bin/bash
# Round-2 visit K (1 GPU): GEMM configurations re-measured for the view-sharded shapes (4 / 2 / 1 views per GPU) and for 64^2 latents with
# the TMA epilogue / 320-column tiles in the candidate set; bench of BASELINE configs[1] and configs[4] (one GPU).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp mvdfusion_b200/gemm_tuning.json gpurun_out/t0.json
timeout 600 python tools/tune_gemm.py --shards 2,4,8 --merge gpurun_out/t0.json --out gpurun_out/t1.json > gpurun_out/tune_shards_v4.log 2>&1; echo "tune shards rc=$?"; tail -1 gpurun_out/tune_shards_v4.log
timeout 900 python tools/tune_gemm.py --latent 64 --shards 1,8 --merge gpurun_out/t1.json --out gpurun_out/t2.json > gpurun_out/tune_s64_v4.log 2>&1; echo "tune s64 rc=$?"; tail -1 gpurun_out/tune_s64_v4.log
cp gpurun_out/t2.json mvdfusion_b200/gemm_tuning.json
timeout 300 python bench.py > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err
timeout 400 python bench.py --latent 64 > gpurun_out/bench_s64_v5.json 2> gpurun_out/bench_s64_v5.err
python - <<PY
import json
for f in ("gpurun_out/bench_v5.json","gpurun_out/bench_s64_v5.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["e2e"]["value"], d["step_roofline"])
PY
