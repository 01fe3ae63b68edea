mkdir -p gpurun_out
timeout 300 python - <<'PY' > gpurun_out/r02_bwd128_checks.log 2>&1
import json
from tests.kernel_checks import CHECKS
for n in [k for k in CHECKS if k.startswith("attn_bwd") and "fused" not in k]:
    try:
        r = CHECKS[n]()
        print(n, r["ok"], r.get("max_err"), flush=True)
        if not r["ok"]: print(json.dumps(r)[:1500], flush=True)
    except Exception as e:
        print(n, "EXC", str(e)[:300], flush=True)
PY
tail -30 gpurun_out/r02_bwd128_checks.log
timeout 200 python tools/attn_vs_sdpa.py r02_bwd128 > gpurun_out/r02_attn_bwd128.log 2>&1; tail -4 gpurun_out/r02_attn_bwd128.log
STB_ATTN_BWD_DKDV=64 timeout 200 python tools/attn_vs_sdpa.py r02_dkdv64_dq128 > gpurun_out/r02_attn_mix1.log 2>&1; tail -4 gpurun_out/r02_attn_mix1.log
STB_ATTN_BWD_DQ=64 timeout 200 python tools/attn_vs_sdpa.py r02_dkdv128_dq64 > gpurun_out/r02_attn_mix2.log 2>&1; tail -4 gpurun_out/r02_attn_mix2.log
timeout 120 tools/probe/sm_probe > gpurun_out/r02_sm_probe_v2.log 2>&1; cat gpurun_out/r02_sm_probe_v2.log
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_pytest_gpu_run3.log; tail -40 gpurun_out/r02_pytest_gpu_run3.log
