#!/bin/bash
# Round-2 visit B: split-precision (hi/lo) operands — kernel tests, the hardened parity suite at MVD_HILO=2, step cost of each level,
# racecheck of the pair-mode prologue fix, calibration of the tcgen05 tensor-op counters on a GEMM of known FLOPs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > gpurun_out/t_ops.log 2>&1; echo "ops rc=$?"; tail -5 gpurun_out/t_ops.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/t_parity_hilo2.log 2>&1; echo "parity rc=$?"; tail -8 gpurun_out/t_parity_hilo2.log
cp gpurun_out/parity.jsonl gpurun_out/parity_hilo2.jsonl
for l in 0 1 2; do
  MVD_HILO=$l timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_hilo$l.json 2> gpurun_out/bench_hilo$l.err
  echo "bench hilo=$l rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/bench_hilo$l.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['kernels_per_step'])")"
done
for c in 22 24; do
  timeout 240 compute-sanitizer --tool racecheck tests/native/gemm_check $c > gpurun_out/san2_racecheck_gemm_$c.log 2>&1
  echo "san racecheck gemm case $c rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/san2_racecheck_gemm_$c.log | tail -1)"
done
M="sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32.sum,sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32.sum.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,sm__cycles_elapsed.avg,gpu__time_duration.sum,sm__inst_executed_pipe_tc.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tmem.sum"
for c in 27 0 9; do
  timeout 200 ncu --metrics $M --clock-control none -k regex:gemm_tc -s 5 -c 2 --csv --log-file gpurun_out/ncu_tc_calib_$c.csv tests/native/gemm_check bench $c > gpurun_out/ncu_tc_calib_$c.log 2>&1
  echo "ncu calib $c rc=$?"
done
tail -25 gpurun_out/ncu_tc_calib_27.csv | cut -c1-260
