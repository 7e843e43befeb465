#!/bin/bash
# Round-2 visit J (1 GPU): compute-sanitizer over the kernels added this round (TMA-epilogue GEMM incl. 320-column pair tiles, fused DiT
# kernel), warm ncu launch list of the bench command with the tcgen05-aware tensor-pipe counter and DRAM traffic, DiT phase timeline.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/san
for c in 2 5 7 33 34 35 36; do
  timeout 120 compute-sanitizer --tool memcheck tests/native/gemm_check $c > gpurun_out/san/san3_memcheck_gemm_$c.log 2>&1; echo "memcheck gemm $c rc=$? $(grep -c 'ERROR SUMMARY: 0 errors' gpurun_out/san/san3_memcheck_gemm_$c.log)"
done
for c in 2 33 35; do
  timeout 200 compute-sanitizer --tool racecheck tests/native/gemm_check $c > gpurun_out/san/san3_racecheck_gemm_$c.log 2>&1; echo "racecheck gemm $c rc=$? $(grep -c 'RACECHECK SUMMARY: 0 hazards' gpurun_out/san/san3_racecheck_gemm_$c.log)"
done
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -q -m gpu -k "test_gridattn_dit_kernel and (384-8-3 or 160-8-3 or 512-4-2) or test_dit_fold" > gpurun_out/san/san3_memcheck_dit.log 2>&1; echo "memcheck dit rc=$?"; tail -3 gpurun_out/san/san3_memcheck_dit.log
timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_ops.py -q -m gpu -k "test_gridattn_dit_kernel and 384-8-1" > gpurun_out/san/san3_racecheck_dit.log 2>&1; echo "racecheck dit rc=$?"; tail -3 gpurun_out/san/san3_racecheck_dit.log
MVD_B200_LIB=$PWD/mvdfusion_b200/libmvd_b200_trace.so timeout 200 python tools/dit_trace.py > gpurun_out/dit_trace.txt 2>&1; head -3 gpurun_out/dit_trace.txt
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
timeout 900 ncu --metrics $M --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --reps 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu-list rc=$?"
python tools/ncu_summary.py gpurun_out/launches.csv --out gpurun_out/launches_summary.json --traffic gpurun_out/gemm_traffic.json --how "ncu --cache-control none --clock-control none, eager launches (--no-graph), third step of the run" | head -12
