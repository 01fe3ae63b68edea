#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --config sd3_fullft --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_sd3_fullft_n8.json 2> gpurun_out/err34.txt
tail -n 3 gpurun_out/err34.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_sd3_fullft_n8.json").read().strip().splitlines()[-1])
print(round(d["value"], 2), round(d["ms_per_step"], 2), d.get("per_rank_ms_per_step"), d["config"].get("grad_exchange"), d["clocks"])
PY
