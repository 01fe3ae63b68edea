#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_full.log
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/r02_bench_n1_v3.json 2> gpurun_out/err29a.txt
timeout 600 python bench.py --config sd3_fullft --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_sd3_fullft_n1_v2.json 2> gpurun_out/err29b.txt
timeout 600 python bench.py --config flux_lokr --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_bench_flux_lokr_n1_v2.json 2> gpurun_out/err29c.txt
timeout 600 python bench.py --config vae_encode --steps 5 --warmup 3 > gpurun_out/r02_bench_vae_n1_v2.json 2> gpurun_out/err29d.txt
tail -4 gpurun_out/r02_pytest_gpu_full.log
python - <<'PY'
import json
for f in ("r02_bench_n1_v3", "r02_bench_sd3_fullft_n1_v2", "r02_bench_flux_lokr_n1_v2", "r02_bench_vae_n1_v2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 3), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 3), d.get("peak_mem_gb"), (d.get("gpu_eager_baseline") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
