mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_run15.log; tail -15 gpurun_out/r02_pytest_gpu_run15.log
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_bench_n1_v2.json 2> gpurun_out/err15.txt; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n1_v2.json').read().strip().splitlines()[-1]); print('flux', d['value'], d['ms_per_step'], d['e2e']['value'], d['peak_mem_gb'], d['kernels'])"; tail -2 gpurun_out/err15.txt
timeout 400 python bench.py --config pixart_lora --steps 4 --warmup 3 > gpurun_out/r02_bench_pixart_n1.json 2> gpurun_out/err15b.txt; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_pixart_n1.json').read().strip().splitlines()[-1]); print('pixart', d['value'], d['ms_per_step'], d['e2e']['value'], d['model_tflops'])"; tail -2 gpurun_out/err15b.txt
timeout 300 python bench.py --config vae_encode --steps 5 --warmup 3 > gpurun_out/r02_bench_vae_n1.json 2> gpurun_out/err15c.txt; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_vae_n1.json').read().strip().splitlines()[-1]); print('vae', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'])"; tail -2 gpurun_out/err15c.txt
