#!/bin/bash
# GPU run 20: full GPU suite (log kept as the round's evidence) + LoKr bench (BASELINE configs[3], N=1 leg)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu_full.log
timeout 900 python bench.py --config flux_lokr --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_bench_flux_lokr_n1.json 2> gpurun_out/err20a.txt
tail -3 gpurun_out/r02_pytest_gpu_full.log
tail -5 gpurun_out/err20a.txt
python - <<'PY'
import json
for f in ("r02_bench_flux_lokr_n1",):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("peak_mem_gb"), d.get("kernels"))
    except Exception as e:
        print(f, "ERR", e)
PY
