#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --config pixart_lora --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_pixart_graph.json 2> gpurun_out/err32a.txt
timeout 600 python bench.py --config pixart_lora --steps 2 --warmup 3 --no-cpu-baseline --graph off > gpurun_out/r02_pixart_eager.json 2> gpurun_out/err32b.txt
OMP_NUM_THREADS=1 timeout 600 python bench.py --config pixart_lora --steps 2 --warmup 3 --no-cpu-baseline --graph off > gpurun_out/r02_pixart_eager_omp1.json 2> gpurun_out/err32c.txt
python -m pytest tests/test_pixart_parity_gpu.py tests/test_flux_parity_gpu.py tests/test_sd3_fullft_gpu.py tests/test_step_graph_gpu.py -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_run32_pytest.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_bench_flux_savenh.json 2> gpurun_out/err32d.txt
tail -n 4 gpurun_out/err32a.txt
cat gpurun_out/r02_run32_pytest.log
python - <<'PY'
import json
for f in ("r02_pixart_graph", "r02_pixart_eager", "r02_pixart_eager_omp1", "r02_bench_flux_savenh"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), d["config"].get("cuda_graph"), d.get("peak_mem_gb"))
    except Exception as e:
        print(f, "ERR", e)
PY
