#!/bin/bash
# GPU run 17: LoRA-target parity + dropout tests, then Flux benches with all+ffs (and default dropout 0.1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_flux_parity_gpu.py tests/test_lora_dropout_gpu.py tests/test_kernels_gpu.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_run17_pytest.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline --lora-target all+ffs > gpurun_out/r02_bench_flux_allffs.json 2> gpurun_out/err17a.txt
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline --lora-dropout 0.1 > gpurun_out/r02_bench_flux_dropout01.json 2> gpurun_out/err17b.txt
tail -3 gpurun_out/r02_run17_pytest.log
python - <<'PY'
import json
for f in ("r02_bench_flux_allffs", "r02_bench_flux_dropout01"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("peak_mem_gb"))
    except Exception as e:
        print(f, "ERR", e)
PY
