#!/bin/bash
# 2-GPU run: the non-headline BASELINE configs through the multi-rank path (NCCL gradient exchange), max-over-ranks timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --config sd3_fullft --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_sd3_fullft_n2.json 2> gpurun_out/err22a.txt
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --config flux_lokr --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_bench_flux_lokr_n2.json 2> gpurun_out/err22b.txt
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --config pixart_lora --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_pixart_n2.json 2> gpurun_out/err22c.txt
for f in sd3_fullft flux_lokr pixart; do tail -2 gpurun_out/err22*.txt | tail -2; done
python - <<'PY'
import json
for f in ("r02_bench_sd3_fullft_n2", "r02_bench_flux_lokr_n2", "r02_bench_pixart_n2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("per_rank_ms_per_step"), d["config"].get("grad_exchange"))
    except Exception as e:
        print(f, "ERR", e)
PY
