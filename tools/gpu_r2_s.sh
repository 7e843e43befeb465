#!/bin/bash
# Round-2 visit S (1 GPU): profiles of build v6 — warm ncu launch list of the bench command (time, DRAM traffic, tcgen05-aware tensor-pipe
# counter), ncu --set full of the QKV / GEGLU GEMMs with and without the LayerNorm fold (tools/ln_fold_probe.py)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
timeout 1200 ncu --metrics $M --clock-control none --cache-control none -c 2600 --csv --log-file gpurun_out/launches_v6.csv python bench.py --steps 2 --warmup 1 --reps 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench_v6.log 2>&1; echo "ncu-list rc=$?"
python tools/ncu_summary.py gpurun_out/launches_v6.csv --out gpurun_out/launches_summary_v6.json --traffic gpurun_out/gemm_traffic_v6.json --how "ncu --cache-control none --clock-control none, eager launches (--no-graph), third step of the run" > gpurun_out/ncu_summary_v6.txt 2>&1; head -8 gpurun_out/ncu_summary_v6.txt
timeout 120 python tools/ln_fold_probe.py --time > gpurun_out/ln_probe_time3.txt 2>&1; cat gpurun_out/ln_probe_time3.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 4 -f -o gpurun_out/prof_ln_probe_v6 python tools/ln_fold_probe.py > gpurun_out/ncu_ln_probe_v6.log 2>&1; echo "ncu rc=$?"
for X in 4096 1024; do
  MVD_LN_FOLD_MAX_ROWS=$X timeout 200 python bench.py --no-cpu-baseline --reps 3 > gpurun_out/bench_lnfold_max$X.json 2> gpurun_out/bench_lnfold_max$X.err; echo "bench max-rows $X rc=$?"
done
timeout 200 python bench.py --no-cpu-baseline --reps 3 > gpurun_out/bench_lnfold_all.json 2> gpurun_out/bench_lnfold_all.err; echo "bench all rc=$?"
python - <<'PY'
import json
for n in ("lnfold_max4096", "lnfold_max1024", "lnfold_all"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"], 2), round(d["ms_per_step"], 4), d["kernels_per_step"], round(d["e2e"]["value"], 2), d["roofline"]["achieved"], d["roofline"]["frac"])
    except Exception as e:
        print(n, "failed", e)
PY
