#!/bin/bash
# Round-end GPU visit: full GPU test-suite, smoke, bench (both arms), ncu launch list of one step and full captures of the
# dominant kernels.  Everything lands in gpurun_out/; copy what should be judged into profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "gpu-tests rc=$?"; tail -3 gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --kernel-table gpurun_out/kernels.json > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench-ref rc=$?"
cut -c1-400 gpurun_out/bench.json; cut -c1-400 gpurun_out/bench_ref.json
# launch list (eager path: this driver's ncu dies inside replayed graphs of this size), 3 steps; the digest takes the last one
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu-list rc=$?"
python tools/ncu_summary.py gpurun_out/launches.csv --out gpurun_out/launches_summary.json --traffic gpurun_out/gemm_traffic.json | head -8
# full captures: QKV GEMM (epilogue-bound, 3 warpgroups), GEGLU GEMM, conv (pair mode), GroupNorm
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 20 -c 1 -f -o gpurun_out/prof_gemm_qkv tests/native/gemm_check bench 13 > /dev/null 2>&1; echo "ncu-qkv rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 20 -c 1 -f -o gpurun_out/prof_gemm_geglu tests/native/gemm_check bench 9 > /dev/null 2>&1; echo "ncu-geglu rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 20 -c 1 -f -o gpurun_out/prof_gemm_linres tests/native/gemm_check bench 0 > /dev/null 2>&1; echo "ncu-linres rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_cluster -s 10 -c 1 -f -o gpurun_out/prof_groupnorm tests/native/norm_bench > /dev/null 2>&1; echo "ncu-gn rc=$?"
ls -la gpurun_out/*.ncu-rep
