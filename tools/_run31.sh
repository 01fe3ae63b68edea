#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OMP_NUM_THREADS=1 timeout 600 python bench.py --config pixart_lora --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_pixart_omp1.json 2> gpurun_out/err31a.txt
timeout 600 python bench.py --config pixart_lora --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_pixart_default.json 2> gpurun_out/err31b.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 1 --config pixart_lora --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_pixart_torchrun1.json 2> gpurun_out/err31c.txt
python - <<'PY'
import json
for f in ("r02_pixart_omp1", "r02_pixart_default", "r02_pixart_torchrun1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), round(d["ms_per_step"], 2), d["clocks"])
    except Exception as e:
        print(f, "ERR", e)
PY
