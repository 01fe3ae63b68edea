#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_flux_lokr_gpu.py tests/test_loss_curve_gpu.py "tests/test_kernels_gpu.py" -m gpu -q -k "lokr or gelu or loss_curve" 2>&1 | tail -25 > gpurun_out/r02_run21_pytest.log
timeout 900 python bench.py --config flux_lokr --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_bench_flux_lokr_n1.json 2> gpurun_out/err21a.txt
tail -4 gpurun_out/r02_run21_pytest.log
tail -5 gpurun_out/err21a.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_flux_lokr_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d.get("peak_mem_gb"), d.get("kernels"))
PY
