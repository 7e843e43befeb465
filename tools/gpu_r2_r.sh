#!/bin/bash
# Round-2 visit R (1 GPU): validation of build v6 (LayerNorm fold on by default) — whole GPU suite, sanitizer over the new epilogue paths,
# bench of BASELINE configs[1] and configs[4] (one GPU) with the fold on / off, one rank of the 8-way sharded step, warm ncu launch list, smoke.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/san
rm -f gpurun_out/parity.jsonl
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests_v6.log 2>&1; echo "all-tests rc=$?"; tail -3 gpurun_out/gpu_tests_v6.log
K="test_gemm_layernorm_statistics_from_the_epilogue and (130-64 or 2048-320) or test_gemm_layernorm_folded_into_qkv_and_geglu and (1-200-64 or 2-16-1280)"
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -q -m gpu -k "$K" > gpurun_out/san/san4_memcheck_lnfold.log 2>&1; echo "memcheck rc=$? $(grep -c 'ERROR SUMMARY: 0 errors' gpurun_out/san/san4_memcheck_lnfold.log)"
timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_ops.py -q -m gpu -k "test_gemm_layernorm_folded_into_qkv_and_geglu and 1-200-64 or test_gemm_layernorm_statistics_from_the_epilogue and 130-64" > gpurun_out/san/san4_racecheck_lnfold.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san/san4_racecheck_lnfold.log | tail -3
timeout 400 python bench.py > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; echo "bench rc=$?"
timeout 300 python bench.py --latent 64 --no-cpu-baseline --reps 3 > gpurun_out/bench_s64_v6.json 2> gpurun_out/bench_s64_v6.err; echo "bench s64 rc=$?"
MVD_NO_LN_FOLD=1 timeout 300 python bench.py --latent 64 --no-cpu-baseline --reps 3 > gpurun_out/bench_s64_v6_lnpass.json 2> gpurun_out/bench_s64_v6_lnpass.err; echo "bench s64 pass rc=$?"
python - <<'PY'
import json
for n in ("v6", "s64_v6", "s64_v6_lnpass"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"], 2), round(d["ms_per_step"], 4), d["kernels_per_step"], round(d["e2e"]["value"], 2), d["roofline"]["achieved"], d["roofline"]["frac"], d["step_roofline"])
    except Exception as e:
        print(n, "failed", e)
PY
timeout 200 python tools/step_profile.py --world 8 --reps 7 > gpurun_out/step_profile_w8_v6.txt 2>&1; head -1 gpurun_out/step_profile_w8_v6.txt
MVD_NO_LN_FOLD=1 timeout 200 python tools/step_profile.py --world 8 --reps 7 > gpurun_out/step_profile_w8_v6_lnpass.txt 2>&1; head -1 gpurun_out/step_profile_w8_v6_lnpass.txt
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
timeout 900 ncu --metrics $M --clock-control none --cache-control none -c 1400 --csv --log-file gpurun_out/launches_v6.csv python bench.py --steps 2 --warmup 1 --reps 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench_v6.log 2>&1; echo "ncu-list rc=$?"
python tools/ncu_summary.py gpurun_out/launches_v6.csv --out gpurun_out/launches_summary_v6.json --traffic gpurun_out/gemm_traffic_v6.json --how "ncu --cache-control none --clock-control none, eager launches (--no-graph), third step of the run" | head -8
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_v6.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_v6.log
