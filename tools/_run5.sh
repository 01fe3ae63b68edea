mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_shim_gpu.py -q -x -k "attn or shim or pair" 2>&1 | tail -5
timeout 200 python tools/attn_vs_sdpa.py r02_v2 > gpurun_out/r02_attn_v2.log 2>&1; head -1 gpurun_out/r02_attn_v2.log | cut -c1-400
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r02_bench_n1_v1.json 2> gpurun_out/r02_bench_n1_v1.err; tail -c 2500 gpurun_out/r02_bench_n1_v1.json; tail -3 gpurun_out/r02_bench_n1_v1.err
