#!/bin/bash
# One GPU-box visit: new-kernel tests first (bounded), then the whole GPU suite, then A/B benches.  Logs -> gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "fp16_copy or groupnorm or layernorm" > gpurun_out/t_new.log 2>&1
echo "new-tests rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/t_new.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1
echo "all-tests rc=$?" | tee -a gpurun_out/summary.txt
tail -8 gpurun_out/t_all.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err
echo "bench-new rc=$?" | tee -a gpurun_out/summary.txt
# A/B switches that exist in the build: MVD_NO_FUSE_CAT (cast / concat passes instead of producer-written operands),
# MVD_GEMM_WG2 (two epilogue warpgroups everywhere), MVD_GEMM_NO_PAIR, MVD_NO_PDL, MVD_GEMM_NO_TUNING
MVD_NO_FUSE_CAT=1 MVD_GEMM_WG2=1 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_old.json 2> gpurun_out/bench_old.err
echo "bench-old rc=$?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
for n in ("new", "old"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["kernels_per_step"], d["e2e"]["value"], [(k["kernel"], k["calls"], round(k["ms"], 3)) for k in d["kernels"][:6]])
    except Exception as e:
        print(n, "failed", e)
PY
