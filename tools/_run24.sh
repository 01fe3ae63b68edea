#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_text_encoders.py tests/test_text_oracle.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_run24_pytest.log
python tools/one_kernel.py time_tail > gpurun_out/r02_tail_kernels.json 2> gpurun_out/err24a.txt
timeout 600 python bench.py --config text_encode --steps 5 --warmup 3 > gpurun_out/r02_bench_text_n1.json 2> gpurun_out/err24b.txt
python -m pytest tests/test_adamw_bf16.py tests/test_flux_parity_gpu.py tests/test_sd3_fullft_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -5 >> gpurun_out/r02_run24_pytest.log
tail -12 gpurun_out/r02_run24_pytest.log
cat gpurun_out/r02_tail_kernels.json; tail -3 gpurun_out/err24a.txt gpurun_out/err24b.txt
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_text_n1.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["achieved"], d.get("peak_mem_gb"))
except Exception as e:
    print("ERR", e)
PY
