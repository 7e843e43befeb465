#!/bin/bash
# Round-2 visit G (8 GPUs): the driver's N=8 command — replicas headline, view-sharded record with the in-graph all-gather and the
# sharded-vs-unsharded check, BASELINE configs[2] (N=16, 2 views/GPU) and configs[4] (64x64 latents) keys.  Bounded by `timeout`.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench n8 rc=$?"; tail -3 gpurun_out/bench_n8.err | cut -c1-300; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n8.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","timed_repetitions","n_gpus")}); print(d.get("sharded")); print(d.get("configs2_n16_sharded")); print(d.get("configs4_s64_sharded")); print(d["e2e"]["value"])
PY
