#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 2 --config sd3_fullft --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_sd3_fullft_n2_pipelined.json 2> gpurun_out/err30a.txt
STB_GRAD_CHUNKS=0 timeout 600 $TR --master-port 29522 bench.py --gpus 2 --config sd3_fullft --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_sd3_fullft_n2_flat.json 2> gpurun_out/err30b.txt
timeout 600 $TR --master-port 29523 bench.py --gpus 2 --config pixart_lora --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_pixart_n2.json 2> gpurun_out/err30c.txt
tail -n 3 gpurun_out/err30a.txt gpurun_out/err30b.txt gpurun_out/err30c.txt
python - <<'PY'
import json
for f in ("r02_bench_sd3_fullft_n2_pipelined", "r02_bench_sd3_fullft_n2_flat", "r02_bench_pixart_n2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), round(d["ms_per_step"], 2), d.get("per_rank_ms_per_step"), d["config"].get("grad_exchange"), d["steps"])
    except Exception as e:
        print(f, "ERR", e)
PY
