#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for rep in 1 2; do
for nh in 1 0; do
STB_SAVE_NH=$nh timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_ab_nh${nh}_$rep.json 2> gpurun_out/err33.txt
done
done
python - <<'PY'
import json
for rep in (1, 2):
    for nh in (1, 0):
        d = json.loads(open(f"gpurun_out/r02_ab_nh{nh}_{rep}.json").read().strip().splitlines()[-1])
        print("SAVE_NH", nh, "rep", rep, round(d["ms_per_step"], 2), d["clocks"]["sm_mhz"], d.get("peak_mem_gb"))
PY
