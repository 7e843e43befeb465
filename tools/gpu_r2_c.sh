#!/bin/bash
# Round-2 visit C: A/B of GEMM kernel variants on the SAME box (per-shape graph-timed bench), ops tests, step bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu > gpurun_out/t_ops.log 2>&1; echo "ops rc=$?"; tail -3 gpurun_out/t_ops.log
tests/native/gemm_check > gpurun_out/gc.log 2>&1; echo "gemm_check rc=$?"; tail -1 gpurun_out/gc.log
for r in 1 2; do
  _ab_old/tests/native/gemm_check bench > gpurun_out/gb_old_$r.log 2>&1
  tests/native/gemm_check bench > gpurun_out/gb_new_$r.log 2>&1
done
paste <(cut -c1-58 gpurun_out/gb_old_2.log) <(cut -c49-58 gpurun_out/gb_new_2.log) <(cut -c49-58 gpurun_out/gb_old_1.log) <(cut -c49-58 gpurun_out/gb_new_1.log)
for l in 0 2; do
  MVD_HILO=$l timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_hilo$l.json 2> gpurun_out/bench_hilo$l.err
  echo "bench hilo=$l rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/bench_hilo$l.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['kernels_per_step'], d['roofline']['achieved'])")"
done
