#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_text_encoders.py -m gpu -q 2>&1 | grep -E "^E |passed|failed|cos|assert" | head -30 > gpurun_out/r02_run26.log
for v in 0 1 2 3 4 5; do STB_ROPE_BWD_VARIANT=$v STB_ROPE_FWD_VARIANT=$((v==5)) python tools/one_kernel.py time_rope_bwd >> gpurun_out/r02_run26.log 2>&1; done
python tools/one_kernel.py time_tail >> gpurun_out/r02_run26.log 2>&1
python tools/vae_bench.py >> gpurun_out/r02_run26.log 2>&1
STB_ROPE_BWD_VARIANT=5 STB_ROPE_FWD_VARIANT=1 python -m pytest tests/test_kernels_gpu.py tests/test_flux_parity_gpu.py -m gpu -q -k "rope or parity_small or ragged" 2>&1 | tail -3 >> gpurun_out/r02_run26.log
cat gpurun_out/r02_run26.log | cut -c1-600
