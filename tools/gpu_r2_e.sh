#!/bin/bash
# Round-2 visit E (2 GPUs): view-sharded mode with the all-gather captured inside the step graph; sharded-vs-unsharded check on hardware.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?"; tail -5 gpurun_out/bench_n2.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","timed_repetitions","n_gpus")}, d.get("sharded"), d["e2e"]["value"])
PY
