#!/bin/bash
# w_kn (dgrad on the forward weights, no transposed copies): kernel checks, parity suites and benches with STB_DGRAD_WKN=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r02_run27.log
python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "wkn or gemm" 2>&1 | tail -6 > $L
STB_DGRAD_WKN=1 python -m pytest tests/test_flux_parity_gpu.py tests/test_sd3_parity_gpu.py tests/test_sd3_fullft_gpu.py tests/test_flux_lokr_gpu.py tests/test_lora_dropout_gpu.py tests/test_fullwidth_parity_gpu.py -m gpu -q 2>&1 | tail -6 >> $L
STB_DGRAD_WKN=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_bench_flux_wkn.json 2> gpurun_out/err27a.txt
STB_DGRAD_WKN=1 timeout 600 python bench.py --config sd3_fullft --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/r02_bench_sd3_wkn.json 2> gpurun_out/err27b.txt
timeout 600 python bench.py --config sd3_fullft --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/r02_bench_sd3_base.json 2> gpurun_out/err27c.txt
cat $L
python - <<'PY'
import json
for f in ("r02_bench_flux_wkn", "r02_bench_sd3_wkn", "r02_bench_sd3_base"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("peak_mem_gb"), (d.get("kernels") or {}).get("gemm"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/err27a.txt gpurun_out/err27b.txt
