#!/bin/bash
# Round-2 visit A: baseline of the hardened parity cases on the round-1 kernels, tcgen05-aware ncu counter names,
# compute-sanitizer passes (memcheck / racecheck) over the native GEMM / GroupNorm checks and one smoke step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "sweep or ten_steps or full_size" > gpurun_out/t_parity_base.log 2>&1
echo "parity-base rc=$?"; tail -15 gpurun_out/t_parity_base.log
cp gpurun_out/parity.jsonl gpurun_out/parity_base.jsonl 2>/dev/null
ncu --query-metrics > gpurun_out/ncu_metrics_all.txt 2>&1
grep -iE 'umma|utc|pipe_tc|tmem|tensor' gpurun_out/ncu_metrics_all.txt > gpurun_out/ncu_metrics_tensor.txt
wc -l gpurun_out/ncu_metrics_all.txt gpurun_out/ncu_metrics_tensor.txt
for tool in memcheck racecheck; do
  for c in 9 10 22 23 24 25; do
    timeout 240 compute-sanitizer --tool $tool tests/native/gemm_check $c > gpurun_out/san_${tool}_gemm_$c.log 2>&1
    echo "san $tool gemm case $c rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_${tool}_gemm_$c.log | tail -1)"
  done
  timeout 300 compute-sanitizer --tool $tool tests/native/norm_bench 1 > gpurun_out/san_${tool}_norm.log 2>&1
  echo "san $tool norm rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_${tool}_norm.log | tail -1)"
done
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_memcheck_smoke.log 2>&1
echo "san memcheck smoke rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/san_memcheck_smoke.log | tail -1)"
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_r2_base.json 2> gpurun_out/bench_r2_base.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_r2_base.json
