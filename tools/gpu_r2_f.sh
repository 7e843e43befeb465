#!/bin/bash
# Round-2 visit F (1 GPU): whole GPU suite (incl. training gradients), smoke, bench (native + train mode), warm ncu launch list with the
# tcgen05-aware tensor-pipe counter and DRAM traffic.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "gpu-tests rc=$?"; tail -4 gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --kernel-table gpurun_out/kernels.json > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench.json
timeout 600 python bench.py --mode train --steps 5 --warmup 2 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train rc=$?"; cut -c1-700 gpurun_out/bench_train.json; tail -3 gpurun_out/bench_train.err
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
timeout 900 ncu --metrics $M --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --reps 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu-list rc=$?"
python tools/ncu_summary.py gpurun_out/launches.csv --out gpurun_out/launches_summary.json --traffic gpurun_out/gemm_traffic.json --how "ncu --cache-control none --clock-control none, eager launches (--no-graph), third step of the run" | head -8
