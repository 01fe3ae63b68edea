mkdir -p gpurun_out
for dp in flat ddp; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 8 --warmup 3 --dp $dp --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_bench_n8_$dp.json 2> gpurun_out/r02_bench_n8_$dp.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r02_bench_n8_$dp.json')); print('$dp', d['value'], d['ms_per_step'], d['e2e']['value'])
except Exception as e: print('$dp failed', e)"
  tail -2 gpurun_out/r02_bench_n8_$dp.err
done
